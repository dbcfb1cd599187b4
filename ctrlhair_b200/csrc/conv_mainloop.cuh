// Main loop of the implicit-GEMM conv kernel: the TMA producer role and the tcgen05 MMA issuer role.
//
// Three ways an A operand (activations) reaches the tensor core:
//   * per-tap tiles   — one 128-row TMA box per (tap, 64/32-channel chunk), tap offset folded into the coordinates
//   * halo tiles      — one (TH+2)x(TW+2) box per channel chunk; the taps are shifted UMMA views of it
// and two ways for the B operand (weights):
//   * streamed        — 64-K chunks through the mbarrier stage ring (1 or 3 taps per stage)
//   * weight-stationary (p.wstat) — the whole [BN x K] weight slab of this CTA's n_tile is loaded once and stays
//     in shared memory; only halo tiles stream.  Used by the layers whose weights are small (mlp_shared, conv_img,
//     the 64-channel convs of up_3): those are bound by per-chunk barrier/TMA overhead, not by math.
#pragma once
#include "conv_epilogue.cuh"

namespace chb {

struct TileOrigin {
  int x0, y0, b0, n0;
};
__device__ __forceinline__ TileOrigin tile_origin(const ConvKParams& p, int tile) {
  const int n_tile = tile / p.m_tiles;
  int m = tile - n_tile * p.m_tiles;
  const int xt = m % p.tiles_x;
  m /= p.tiles_x;
  const int yt = m % p.tiles_y;
  const int bt = m / p.tiles_y;
  TileOrigin o;
  o.x0 = xt * p.TW; o.y0 = yt * p.TH; o.b0 = bt * p.TB; o.n0 = n_tile * p.BN;
  return o;
}

template <bool WSTAT>
__device__ __forceinline__ void producer_role(const ConvKParams& p, const Smem& sm) {
  const uint32_t nst = (uint32_t)p.nstages, nh = (uint32_t)p.nhalo;
  uint32_t stage = 0, phase = 0, hs = 0, hphase = 0;
  // The whole warp executes these loops in lock step and one elected lane issues: loop state stays in uniform
  // registers, which keeps the compiler from wrapping every TMA / mbarrier instruction in a per-thread loop.
  if (WSTAT) {
    const int n0 = ((int)blockIdx.x % p.n_tiles) * p.BN;
    if (elect_one()) mbar_arrive_expect_tx(sm.wbar, (uint32_t)p.wstat_bytes);
    __syncwarp();
    for (int s = 0; s < p.nseg; ++s) {
      const SegK sg = p.seg[s];
      const uint32_t wbytes = (uint32_t)p.BN * (uint32_t)sg.kc * 2u;
      for (int i = 0; i < sg.taps * sg.nchunk_w; ++i) {
        if (elect_one())
          tma_load_3d(&p.tmW[s], sm.wstat_base + sg.wofs + (size_t)i * wbytes, sm.wbar, i * sg.kc, n0, 0);
        __syncwarp();
      }
    }
  }
  for (uint32_t t = 0;; ++t) {
    int tile = sched_tile<WSTAT>(p, t);
    if (tile < 0) break;
    if (t == 0 && (threadIdx.x & 31) == 0) CHB_TRACE_AT(2);
    const int ks = split_of_tile(p, tile);
    const TileOrigin o = tile_origin(p, tile);
    for (int s = 0; s < p.nseg; ++s) {
      const SegK sg = p.seg[s];
      const uint32_t wbytes = (uint32_t)p.BN * (uint32_t)sg.kc * 2u;
      int cb, ce;   // this split's channel chunks (all of them without split-K)
      split_chunks(p, ks, sg.nchunk, cb, ce);
      if (WSTAT || sg.halo) {
        // a 1x1 segment needs no halo: its exact 128-row tile goes into the halo ring buffer instead (29 % fewer bytes)
        const bool exact = sg.taps == 1;
        const uint32_t hbytes = (exact ? (uint32_t)p.rows : (uint32_t)((p.TW + 2) * (p.TH + 2))) * (uint32_t)sg.kc * 2u;
        const int tg = sg.taps == 9 ? p.hg : 1;  // taps per weight stage
        for (int c = cb; c < ce; ++c) {
          const int cw = c >= sg.nchunk_w ? c - sg.nchunk_w : c;  // weight chunk (hi+lo split: both halves share it)
          mbar_wait(&sm.hempty[hs], hphase ^ 1u);
          if (elect_one()) {
            mbar_arrive_expect_tx(&sm.hfull[hs], hbytes);
            if (exact)
              tma_load_4d(&p.tmA[s], sm.halo_base + (size_t)hs * p.halo_buf_bytes, &sm.hfull[hs], sg.ch_off + c * sg.kc,
                          o.x0 + sg.xy_off, o.y0 + sg.xy_off, o.b0);
            else
              tma_load_4d(&p.tmH[s], sm.halo_base + (size_t)hs * p.halo_buf_bytes, &sm.hfull[hs], sg.ch_off + c * sg.kc,
                          o.x0 - 1 + sg.xy_off, o.y0 - 1 + sg.xy_off, o.b0);
          }
          __syncwarp();
          if (++hs == nh) {
            hs = 0;
            hphase ^= 1u;
          }
          if (WSTAT) continue;
          for (int t0 = 0; t0 < sg.taps; t0 += tg) {
            mbar_wait(&sm.empty[stage], phase ^ 1u);
            uint8_t* sb = sm.stage_base + (size_t)stage * p.stage_bytes + p.a_region;
            if (elect_one()) {
              mbar_arrive_expect_tx(&sm.full[stage], wbytes * (uint32_t)tg);
              for (int g = 0; g < tg; ++g)
                tma_load_3d(&p.tmW[s], sb + (size_t)g * wbytes, &sm.full[stage], ((t0 + g) * sg.nchunk_w + cw) * sg.kc,
                            o.n0, sg.per_image ? o.b0 : 0);
            }
            __syncwarp();
            if (++stage == nst) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      } else {
        const uint32_t bytes = (uint32_t)p.rows * (uint32_t)sg.kc * 2u + wbytes;
        // channel chunk outer, tap inner: the same accumulation order as the halo paths, so a layer gives bitwise the
        // same result whether its tiles take this path (several small images per tile) or a halo path (one image)
        for (int c = cb; c < ce; ++c) {
          const int cw = c >= sg.nchunk_w ? c - sg.nchunk_w : c;
          for (int tap = 0; tap < sg.taps; ++tap) {
            const int dy = sg.taps == 9 ? tap / 3 - 1 : 0;
            const int dx = sg.taps == 9 ? tap % 3 - 1 : 0;
            mbar_wait(&sm.empty[stage], phase ^ 1u);
            uint8_t* sa = sm.stage_base + (size_t)stage * p.stage_bytes;
            if (elect_one()) {
              mbar_arrive_expect_tx(&sm.full[stage], bytes);
              tma_load_4d(&p.tmA[s], sa, &sm.full[stage], sg.ch_off + c * sg.kc, o.x0 + dx + sg.xy_off,
                          o.y0 + dy + sg.xy_off, o.b0);
              tma_load_3d(&p.tmW[s], sa + p.a_region, &sm.full[stage], (tap * sg.nchunk_w + cw) * sg.kc, o.n0,
                          sg.per_image ? o.b0 : 0);
            }
            __syncwarp();
            if (++stage == nst) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  }
}

template <bool WSTAT>
__device__ __forceinline__ void mma_role(const ConvKParams& p, const Smem& sm, uint32_t tmem_base) {
  // One elected lane feeds the tensor core, but the whole warp walks the loops in lock step so that descriptors and
  // loop state live in uniform registers (code issued from a divergent `if (lane == 0)` region gets every tcgen05
  // instruction wrapped in a per-thread "waterfall" loop, ~100 cycles per MMA).  Descriptor high words are hoisted
  // per segment, the K steps of a chunk go out in one asm block, tap offsets advance incrementally.
  const uint32_t nst = (uint32_t)p.nstages, nh = (uint32_t)p.nhalo;
  const uint32_t idesc = umma_idesc_f16((uint32_t)p.BN);
  const uint32_t hw = (uint32_t)(p.TW + 2);
  const uint32_t stage_base = smem_u32(sm.stage_base), stage_bytes = (uint32_t)p.stage_bytes;
  const uint32_t halo_base = smem_u32(sm.halo_base), halo_bytes = (uint32_t)p.halo_buf_bytes;
  const uint32_t wstat_base = smem_u32(sm.wstat_base);
  const uint32_t a_region = (uint32_t)p.a_region;
  uint32_t stage = 0, phase = 0, hs = 0, hphase = 0;
  if (WSTAT) {
    mbar_wait(sm.wbar, 0u);
    tc_fence_after();
    if ((threadIdx.x & 31) == 0) CHB_TRACE_AT(9);
  }
  for (uint32_t t = 0;; ++t) {
    int tile = sched_tile<WSTAT>(p, t);
    if (tile < 0) break;
    const int ks = split_of_tile(p, tile);
    const uint32_t acc = t & 1u, acc_phase = (t >> 1) & 1u;
    mbar_wait(&sm.tempty[acc], acc_phase ^ 1u);
    tc_fence_after();
    if (t == 0 && (threadIdx.x & 31) == 0) CHB_TRACE_AT(3);
    const uint32_t d_tmem = tmem_base + acc * 256u;
    uint32_t accumulate = 0;
    for (int s = 0; s < p.nseg; ++s) {
      const SegK sg = p.seg[s];
      const uint32_t row_bytes = (uint32_t)sg.kc * 2u;
      const uint32_t wbytes = (uint32_t)p.BN * row_bytes;
      const uint32_t hiB = umma_desc_hi(row_bytes, row_bytes * 8u);
      const bool k64 = sg.kc == 64;
      int cb, ce;
      split_chunks(p, ks, sg.nchunk, cb, ce);
      if (WSTAT || sg.halo) {
        // tap (ky,kx): tile pixel (y,x) reads halo row (y+ky)*(TW+2) + (x+kx).  With TW == 8 every 8-row core group
        // of the UMMA operand is one tile row, (TW+2)*row_bytes apart.  The swizzle is a function of the absolute
        // shared-memory address bits (verified on B200: base_offset must stay 0 for views that start off the
        // swizzle-atom boundary), so a shifted start address is all a tap needs.
        const int ntaps = sg.taps;
        // 3x3: 8-row core groups are tile rows of the halo tile, (TW+2) pixels apart; 1x1: the exact tile, 8 rows apart
        const uint32_t hiA = umma_desc_hi(row_bytes, ntaps == 9 ? hw * row_bytes : 8u * row_bytes);
        const int tg = ntaps == 9 ? p.hg : 1;
        const uint32_t bstep = WSTAT ? (uint32_t)sg.nchunk_w * wbytes : wbytes;  // weight slab of the next tap
        const uint32_t row_step = hw * row_bytes;
        for (int c = cb; c < ce; ++c) {
          const int cw = c >= sg.nchunk_w ? c - sg.nchunk_w : c;  // resident weight chunk of this activation chunk
          mbar_wait(&sm.hfull[hs], hphase);
          tc_fence_after();
          const uint32_t hb = halo_base + hs * halo_bytes;
          uint32_t voff = 0u;  // view offset of the current tap (a 1x1 segment holds its exact tile)
          uint32_t kx = 0;
          if (WSTAT && ntaps == 9) {
            // resident weights: nothing to wait for between taps, all 9 x KS MMAs of the chunk go out in one asm block
            const uint64_t adesc = umma_desc_make(hiA, hb);
            const uint64_t bdesc = umma_desc_make(hiB, wstat_base + (uint32_t)sg.wofs + (uint32_t)cw * wbytes);
            if (elect_one()) {
              if (k64) umma_f16_ss_tile9<4>(d_tmem, adesc, bdesc, row_bytes >> 4, row_step >> 4, bstep >> 4, idesc, accumulate);
              else umma_f16_ss_tile9<2>(d_tmem, adesc, bdesc, row_bytes >> 4, row_step >> 4, bstep >> 4, idesc, accumulate);
            }
            __syncwarp();
            accumulate = 1;
          } else
          for (int t0 = 0; t0 < ntaps; t0 += tg) {
            uint32_t sb;
            if (WSTAT) {
              sb = wstat_base + (uint32_t)sg.wofs + (uint32_t)(t0 * sg.nchunk_w + cw) * wbytes;
            } else {
              mbar_wait(&sm.full[stage], phase);
              tc_fence_after();
              sb = stage_base + stage * stage_bytes + a_region;
            }
            if (tg == 3) {
              // a whole kernel row (kx = 0, 1, 2) in one asm block
              const uint64_t adesc = umma_desc_make(hiA, hb + voff);
              const uint64_t bdesc = umma_desc_make(hiB, sb);
              if (elect_one()) {
                if (k64) umma_f16_ss_row3<4>(d_tmem, adesc, bdesc, row_bytes >> 4, bstep >> 4, idesc, accumulate);
                else umma_f16_ss_row3<2>(d_tmem, adesc, bdesc, row_bytes >> 4, bstep >> 4, idesc, accumulate);
              }
              __syncwarp();
              accumulate = 1;
              voff += row_step;
            } else
            for (int g = 0; g < tg; ++g) {
              const uint64_t adesc = umma_desc_make(hiA, hb + voff);
              const uint64_t bdesc = umma_desc_make(hiB, sb);
              if (elect_one()) {
                if (k64) umma_f16_ss_k<4>(d_tmem, adesc, bdesc, idesc, accumulate);
                else umma_f16_ss_k<2>(d_tmem, adesc, bdesc, idesc, accumulate);
              }
              __syncwarp();
              accumulate = 1;
              sb += bstep;
              // next tap: one pixel right, or to the start of the next halo row
              if (++kx == 3u) {
                kx = 0;
                voff += row_step - 2u * row_bytes;
              } else {
                voff += row_bytes;
              }
            }
            if (!WSTAT) {
              if (elect_one()) umma_commit(&sm.empty[stage]);
              __syncwarp();
              if (++stage == nst) {
                stage = 0;
                phase ^= 1u;
              }
            }
          }
          if (elect_one()) umma_commit(&sm.hempty[hs]);
          __syncwarp();
          if (++hs == nh) {
            hs = 0;
            hphase ^= 1u;
          }
        }
      } else {
        const int chunks = sg.taps * (ce - cb);
        for (int c = 0; c < chunks; ++c) {
          mbar_wait(&sm.full[stage], phase);
          tc_fence_after();
          const uint32_t sa = stage_base + stage * stage_bytes;
          const uint64_t adesc = umma_desc_make(hiB, sa);
          const uint64_t bdesc = umma_desc_make(hiB, sa + a_region);
          if (elect_one()) {
            if (k64) umma_f16_ss_k<4>(d_tmem, adesc, bdesc, idesc, accumulate);
            else umma_f16_ss_k<2>(d_tmem, adesc, bdesc, idesc, accumulate);
            umma_commit(&sm.empty[stage]);
          }
          __syncwarp();
          accumulate = 1;
          if (++stage == nst) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
    if (elect_one()) umma_commit(&sm.tfull[acc]);
    __syncwarp();
    if (t == 0 && (threadIdx.x & 31) == 0) CHB_TRACE_AT(4);
  }
}

}  // namespace chb
