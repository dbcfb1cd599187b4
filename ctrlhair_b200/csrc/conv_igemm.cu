// Implicit-GEMM 3x3 / 1x1 convolution for sm_100a.
//
//   M = 128 output pixels of a TB x TH x TW tile, N = BN output channels, K = sum over segments of taps * C.
//   A tiles are fetched by TMA straight from the NHWC activation tensor with the tap offset folded into the
//   box coordinates (out-of-bounds -> zero fill = the conv's zero padding), B tiles by TMA from the K-major
//   weight matrix (optionally a per-image one: the SEAN region-factored style weights).  One thread issues
//   tcgen05.mma into a double-buffered fp32 TMEM accumulator, eight epilogue warps drain it with tcgen05.ld
//   and apply the fused epilogue (bias / residual / activation, or the ACE normalise-modulate of
//   sean_codes/models/networks/normalization.py:111-112,177-187) while the next tile's main loop runs.
//
// Reference call sites this operator stands in for: normalization.py:172-173 (conv_gamma/conv_beta),
// :241-256 (SPADE mlp_shared / mlp_gamma / mlp_beta), architecture.py:75,79,90 (conv_0, conv_1, conv_s),
// generator.py:76,107 (fc, conv_img).
#include "conv_igemm.cuh"
#include "ptx_sm100.cuh"

#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace chb {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
const char* last_error_cstr() { return g_err.c_str(); }

int device_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

// ------------------------------------------------------------------------------------------------
// Epilogue math shared by the tcgen05 kernel and the SIMT checker kernel.
// ------------------------------------------------------------------------------------------------
template <int ACT>
__device__ __forceinline__ float act_t(float v) {
  if (ACT == CHB_ACT_RELU) return fmaxf(v, 0.f);
  if (ACT == CHB_ACT_LRELU) return fmaxf(v, 0.2f * v);
  if (ACT == CHB_ACT_TANH) return tanhf(v);
  return v;
}

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case CHB_ACT_RELU: return act_t<CHB_ACT_RELU>(v);
    case CHB_ACT_LRELU: return act_t<CHB_ACT_LRELU>(v);
    case CHB_ACT_TANH: return act_t<CHB_ACT_TANH>(v);
    default: return v;
  }
}

__device__ __forceinline__ long long out_offset(const EpiK& e, int b, int y, int x, int n) {
  long long off = (long long)b * e.o_sb + (long long)y * e.o_sy + (long long)x * e.o_sx;
  if (e.o_ngroup > 0) {
    off += (long long)(n / e.o_ngroup) * e.o_sgroup + (long long)(n % e.o_ngroup) * e.o_sn;
  } else {
    off += (long long)n * e.o_sn;
  }
  return off;
}

template <int ACT>
__device__ __forceinline__ void plain_store_elem(const EpiK& e, int b, int y, int x, int n, float acc) {
  float v = acc;
  if (e.bias) v += __ldg(e.bias + (e.bias_per_image ? (long long)b * e.nrows : 0) + n);
  if (e.res) {
    v += __ldg(e.res + (long long)b * e.r_sb + (long long)(y >> e.r_shift) * e.r_sy +
               (long long)(x >> e.r_shift) * e.r_sx + n);
  }
  v = act_t<ACT>(v);
  const long long off = out_offset(e, b, y, x, n);
  if (e.out_dtype == CHB_F16) {
    reinterpret_cast<__half*>(e.out)[off] = __float2half_rn(v);
  } else {
    reinterpret_cast<float*>(e.out)[off] = v;
  }
}

// xn = (x + noise*noise_var - mean) * rstd  (folded: a = rstd, c = -mean*rstd, nv = noise_var*rstd)
// out = act(xn * (1 + gamma) + beta)
template <int ACT>
__device__ __forceinline__ float modulate_elem(float xv, float nz, float a, float c, float nv, float gamma,
                                               float beta) {
  const float xn = fmaf(xv, a, fmaf(nz, nv, c));
  return act_t<ACT>(fmaf(xn, 1.f + gamma, beta));
}

// ------------------------------------------------------------------------------------------------
// Per-warp staging block (4 KB of shared memory per epilogue warp): 32 tile rows x (CH16 * 16) bytes.
// The TMEM accumulator layout gives every lane one tile row (= one pixel), but pixels are C*elem bytes apart in
// the NHWC tensors, so a row-per-lane global access touches 32 cache lines per instruction.  Going through this
// block turns the global side into 64/128-byte contiguous runs per row (CH16 lanes per row).  The XOR swizzle
// makes both access patterns (row per lane / CH16 lanes per row) free of bank conflicts.
// ------------------------------------------------------------------------------------------------
template <int CH16>
__device__ __forceinline__ uint32_t stg_off(int row, int c) {
  const int swz = CH16 == 8 ? (row & 7) : ((row >> 1) & 3);
  return (uint32_t)(row * (CH16 * 16) + ((c ^ swz) << 4));
}
__device__ __forceinline__ void sts128(uint32_t saddr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr)
               : "memory");
  return v;
}

// Pixel of tile row m (0..127) for the tile at (b0, y0, x0); returns false when the row is padding / out of range.
struct TileGeo {
  int TW, TH, TB, tpix, rows, B, H, W, b0, y0, x0;
  int tw_sh, tpix_sh;  // log2 when TW / TW*TH are powers of two (the generator's tiles), else -1
  __device__ __forceinline__ void init(const ConvKParams& p) {
    TW = p.TW; TH = p.TH; TB = p.TB; tpix = p.TW * p.TH; rows = p.rows; B = p.B; H = p.H; W = p.W;
    tw_sh = (TW & (TW - 1)) == 0 ? 31 - __clz(TW) : -1;
    tpix_sh = (tpix & (tpix - 1)) == 0 ? 31 - __clz(tpix) : -1;
    b0 = y0 = x0 = 0;
  }
  // positions the geometry on `tile`, returns its n-tile index
  __device__ __forceinline__ int set_tile(const ConvKParams& p, int tile) {
    const int n_tile = tile / p.m_tiles;
    int m = tile - n_tile * p.m_tiles;
    const int xt = m % p.tiles_x;
    m /= p.tiles_x;
    const int yt = m % p.tiles_y;
    const int bt = m / p.tiles_y;
    b0 = bt * TB; y0 = yt * TH; x0 = xt * TW;
    return n_tile;
  }
  __device__ __forceinline__ bool pixel(int m, int& b, int& y, int& x) const {
    int tb, rem, ty;
    if (tpix_sh >= 0) { tb = m >> tpix_sh; rem = m & (tpix - 1); } else { tb = m / tpix; rem = m - tb * tpix; }
    if (tw_sh >= 0) { ty = rem >> tw_sh; x = x0 + (rem & (TW - 1)); } else { ty = rem / TW; x = x0 + rem - ty * TW; }
    b = b0 + tb;
    y = y0 + ty;
    return m < rows && b < B && y < H && x < W;
  }
};

// Issue half of a gather: global (CH16*16 contiguous bytes per tile row, CH16 lanes per row) -> registers, still
// in the global (coalesced) arrangement.  rowptr(b,y,x) -> const char*.
template <int CH16, class RowPtr>
__device__ __forceinline__ void gather_issue(int lane, int row_base, const TileGeo& tg, RowPtr rowptr,
                                             uint4 (&g)[CH16]) {
  constexpr int RPI = 32 / CH16;
  const int cl = lane % CH16, r0 = lane / CH16;
#pragma unroll
  for (int k = 0; k < CH16; ++k) {
    int b, y, x;
    g[k] = make_uint4(0u, 0u, 0u, 0u);
    if (tg.pixel(row_base + r0 + RPI * k, b, y, x)) g[k] = __ldg(reinterpret_cast<const uint4*>(rowptr(b, y, x) + cl * 16));
  }
}
// Second half: through the staging block into the one-row-per-lane arrangement.
template <int CH16>
__device__ __forceinline__ void gather_commit(uint32_t stg, int lane, const uint4 (&g)[CH16], uint4 (&regs)[CH16]) {
  constexpr int RPI = 32 / CH16;
  const int cl = lane % CH16, r0 = lane / CH16;
#pragma unroll
  for (int k = 0; k < CH16; ++k) sts128(stg + stg_off<CH16>(r0 + RPI * k, cl), g[k]);
  __syncwarp();
#pragma unroll
  for (int c = 0; c < CH16; ++c) regs[c] = lds128(stg + stg_off<CH16>(lane, c));
  __syncwarp();
}
template <int CH16, class RowPtr>
__device__ __forceinline__ void stage_gather(uint32_t stg, int lane, int row_base, const TileGeo& tg, RowPtr rowptr,
                                             uint4 (&regs)[CH16]) {
  uint4 g[CH16];
  gather_issue<CH16>(lane, row_base, tg, rowptr, g);
  gather_commit<CH16>(stg, lane, g, regs);
}

// registers (one row per lane) -> global, CH16*16 contiguous bytes per tile row.  rowptr(b,y,x) -> char*.
template <int CH16, class RowPtr>
__device__ __forceinline__ void stage_scatter(uint32_t stg, int lane, int row_base, const TileGeo& tg, RowPtr rowptr,
                                              const uint4 (&regs)[CH16]) {
  constexpr int RPI = 32 / CH16;
  const int cl = lane % CH16, r0 = lane / CH16;
#pragma unroll
  for (int c = 0; c < CH16; ++c) sts128(stg + stg_off<CH16>(lane, c), regs[c]);
  __syncwarp();
#pragma unroll
  for (int k = 0; k < CH16; ++k) {
    const int rl = r0 + RPI * k;
    const uint4 v = lds128(stg + stg_off<CH16>(rl, cl));
    int b, y, x;
    if (tg.pixel(row_base + rl, b, y, x)) *reinterpret_cast<uint4*>(rowptr(b, y, x) + cl * 16) = v;
  }
  __syncwarp();
}

// ---- precomputed row offsets: the pixel decode and 64-bit address math are done once per tile, not per access.
// o[k] = byte offset / 16 of the k-th row this lane touches in the CH16-lanes-per-row arrangement (~0u: no row).
template <int CH16, class Fn>
__device__ __forceinline__ void row_offsets(int lane, int row_base, const TileGeo& tg, int elem_bytes, Fn elem_off,
                                            uint32_t (&o)[CH16]) {
  constexpr int RPI = 32 / CH16;
  const int r0 = lane / CH16;
#pragma unroll
  for (int k = 0; k < CH16; ++k) {
    int b, y, x;
    o[k] = tg.pixel(row_base + r0 + RPI * k, b, y, x) ? (uint32_t)((elem_off(b, y, x) * elem_bytes) >> 4) : 0xFFFFFFFFu;
  }
}
template <int CH16>
__device__ __forceinline__ void gather_issue_o(const char* base, const uint32_t (&o)[CH16], uint4 (&g)[CH16]) {
#pragma unroll
  for (int k = 0; k < CH16; ++k) {
    g[k] = make_uint4(0u, 0u, 0u, 0u);
    if (o[k] != 0xFFFFFFFFu) g[k] = __ldg(reinterpret_cast<const uint4*>(base + ((size_t)o[k] << 4)));
  }
}
template <int CH16>
__device__ __forceinline__ void scatter_o(uint32_t stg, int lane, char* base, const uint32_t (&o)[CH16],
                                          const uint4 (&regs)[CH16]) {
  constexpr int RPI = 32 / CH16;
  const int cl = lane % CH16, r0 = lane / CH16;
#pragma unroll
  for (int c = 0; c < CH16; ++c) sts128(stg + stg_off<CH16>(lane, c), regs[c]);
  __syncwarp();
#pragma unroll
  for (int k = 0; k < CH16; ++k) {
    const uint4 v = lds128(stg + stg_off<CH16>(r0 + RPI * k, cl));
    if (o[k] != 0xFFFFFFFFu) *reinterpret_cast<uint4*>(base + ((size_t)o[k] << 4)) = v;
  }
  __syncwarp();
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// 32 accumulator columns of the PLAIN epilogue through the staging block (channels-last output, full block valid).
// ro / oo4 / oo8: per-tile row offsets (16-byte units) of the residual and of the output in the 8- and 4-lanes-per-row
// arrangements; bi = image of this lane's own row (for per-image bias).
template <int ACT>
__device__ __forceinline__ void plain_block32(const EpiK& e, uint32_t taddr, uint32_t stg, int lane, int n, int bi,
                                              const uint32_t (&ro)[8], const uint32_t (&oo4)[4],
                                              const uint32_t (&oo8)[8]) {
  float v[32];
  tmem_ld<32>(taddr, v);
  uint4 rr[8];
  if (e.res) {
    uint4 g[8];
    gather_issue_o<8>(reinterpret_cast<const char*>(e.res + n) + (lane & 7) * 16, ro, g);
    gather_commit<8>(stg, lane, g, rr);
  }
  tmem_ld_fence(v);
  if (e.bias) {
    const float4* bp = reinterpret_cast<const float4*>(e.bias + (e.bias_per_image ? (long long)bi * e.nrows : 0) + n);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 t = __ldg(bp + i);
      v[4 * i] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
    }
  }
  if (e.res) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[4 * i] += __uint_as_float(rr[i].x); v[4 * i + 1] += __uint_as_float(rr[i].y);
      v[4 * i + 2] += __uint_as_float(rr[i].z); v[4 * i + 3] += __uint_as_float(rr[i].w);
    }
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = act_t<ACT>(v[i]);
  const long long noff = e.o_ngroup > 0 ? (long long)(n / e.o_ngroup) * e.o_sgroup + (long long)(n % e.o_ngroup) : n;
  if (e.out_dtype == CHB_F16) {
    uint4 pk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      pk[i] = make_uint4(pack_h2(v[8 * i], v[8 * i + 1]), pack_h2(v[8 * i + 2], v[8 * i + 3]),
                         pack_h2(v[8 * i + 4], v[8 * i + 5]), pack_h2(v[8 * i + 6], v[8 * i + 7]));
    scatter_o<4>(stg, lane, reinterpret_cast<char*>(reinterpret_cast<__half*>(e.out) + noff) + (lane & 3) * 16, oo4, pk);
  } else {
    uint4 pk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
      pk[i] = make_uint4(__float_as_uint(v[4 * i]), __float_as_uint(v[4 * i + 1]), __float_as_uint(v[4 * i + 2]),
                         __float_as_uint(v[4 * i + 3]));
    scatter_o<8>(stg, lane, reinterpret_cast<char*>(reinterpret_cast<float*>(e.out) + noff) + (lane & 7) * 16, oo8, pk);
  }
}

// One chunk of NC accumulator columns of the PLAIN epilogue: TMEM load and the global loads it needs are all
// issued before the single wait, so the (few) epilogue warps have the latencies overlapped.
template <int NC, int ACT>
__device__ __forceinline__ void plain_chunk(const ConvKParams& p, const EpiK& e, uint32_t taddr, int n, bool valid,
                                            int b, int y, int x) {
  float v[NC];
  tmem_ld<NC>(taddr, v);
  const bool inb = valid && n < p.N;
  const bool vec = inb && (n + NC <= p.N) && e.o_sn == 1 && (e.o_ngroup <= 0 || (e.o_ngroup % NC) == 0);
  float4 bv[NC / 4], rv[NC / 4];
#pragma unroll
  for (int i = 0; i < NC / 4; ++i) bv[i] = rv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (vec) {
    if (e.bias) {
      const float4* bp = reinterpret_cast<const float4*>(e.bias + (e.bias_per_image ? (long long)b * e.nrows : 0) + n);
#pragma unroll
      for (int i = 0; i < NC / 4; ++i) bv[i] = __ldg(bp + i);
    }
    if (e.res) {
      const float4* rp = reinterpret_cast<const float4*>(e.res + (long long)b * e.r_sb +
                                                         (long long)(y >> e.r_shift) * e.r_sy +
                                                         (long long)(x >> e.r_shift) * e.r_sx + n);
#pragma unroll
      for (int i = 0; i < NC / 4; ++i) rv[i] = __ldg(rp + i);
    }
  }
  tmem_ld_fence(v);
  if (vec) {
#pragma unroll
    for (int i = 0; i < NC / 4; ++i) {
      v[4 * i] = act_t<ACT>(v[4 * i] + bv[i].x + rv[i].x);
      v[4 * i + 1] = act_t<ACT>(v[4 * i + 1] + bv[i].y + rv[i].y);
      v[4 * i + 2] = act_t<ACT>(v[4 * i + 2] + bv[i].z + rv[i].z);
      v[4 * i + 3] = act_t<ACT>(v[4 * i + 3] + bv[i].w + rv[i].w);
    }
    const long long off = out_offset(e, b, y, x, n);
    if (e.out_dtype == CHB_F16) {
      uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(e.out) + off);
#pragma unroll
      for (int i = 0; i < NC / 8; ++i) {
        uint32_t pk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          __half2 h = __floats2half2_rn(v[8 * i + 2 * k], v[8 * i + 2 * k + 1]);
          pk[k] = *reinterpret_cast<uint32_t*>(&h);
        }
        op[i] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    } else {
      float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + off);
#pragma unroll
      for (int i = 0; i < NC / 4; ++i) op[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    }
  } else if (inb) {
#pragma unroll 1
    for (int i = 0; i < NC; ++i)
      if (n + i < p.N) plain_store_elem<ACT>(e, b, y, x, n + i, v[i]);
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// tcgen05 kernel
// ------------------------------------------------------------------------------------------------
template <int EPI, int ACT>
__global__ void __launch_bounds__(kConvThreads, 1) conv_igemm_kernel(const __grid_constant__ ConvKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nst = p.nstages;
  uint8_t* stage_base = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)nst * p.stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kMaxStages;
  uint64_t* tfull = bars + 2 * kMaxStages;
  uint64_t* tempty = tfull + 2;
  uint64_t* hfull = tempty + 2;
  uint64_t* hempty = hfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(hempty + 2);
  // [stages][256 B barriers][8 x 4 KB epilogue staging][2 halo buffers (1 KB aligned)]
  uint8_t* halo_base = smem + (size_t)nst * p.stage_bytes + 1024 + kEpilogueWarps * 4096;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nseg; ++s) {
      tma_prefetch_desc(&p.tmA[s]);
      tma_prefetch_desc(&p.tmW[s]);
      if (p.seg[s].halo) tma_prefetch_desc(&p.tmH[s]);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&hfull[i], 1);
      mbar_init(&hempty[i], 1);
    }
    for (int i = 0; i < nst; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], kEpilogueWarps);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, hs = 0, hphase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_tile = tile / p.m_tiles;
        int m = tile - n_tile * p.m_tiles;
        const int xt = m % p.tiles_x;
        m /= p.tiles_x;
        const int yt = m % p.tiles_y;
        const int bt = m / p.tiles_y;
        const int x0 = xt * p.TW, y0 = yt * p.TH, b0 = bt * p.TB, n0 = n_tile * p.BN;
        for (int s = 0; s < p.nseg; ++s) {
          const SegK sg = p.seg[s];
          if (sg.halo) {
            // one halo tile per channel chunk feeds all nine taps; only the weights stream per tap
            const uint32_t hbytes = (uint32_t)((p.TW + 2) * (p.TH + 2)) * (uint32_t)sg.kc * 2u;
            const uint32_t wbytes = (uint32_t)p.BN * (uint32_t)sg.kc * 2u;
            for (int c = 0; c < sg.nchunk; ++c) {
              mbar_wait(&hempty[hs], hphase ^ 1u);
              mbar_arrive_expect_tx(&hfull[hs], hbytes);
              tma_load_4d(&p.tmH[s], halo_base + (size_t)hs * kHaloBufBytes, &hfull[hs], sg.ch_off + c * sg.kc, x0 - 1,
                          y0 - 1, b0);
              if (++hs == 2u) {
                hs = 0;
                hphase ^= 1u;
              }
              // weights: `hg` taps (one kernel row when hg == 3) share a pipeline stage and one barrier round
              for (int t0 = 0; t0 < 9; t0 += p.hg) {
                mbar_wait(&empty[stage], phase ^ 1u);
                uint8_t* sb = stage_base + (size_t)stage * p.stage_bytes + p.a_region;
                mbar_arrive_expect_tx(&full[stage], wbytes * (uint32_t)p.hg);
                for (int g = 0; g < p.hg; ++g)
                  tma_load_3d(&p.tmW[s], sb + (size_t)g * wbytes, &full[stage], ((t0 + g) * sg.nchunk + c) * sg.kc, n0,
                              sg.per_image ? b0 : 0);
                if (++stage == (uint32_t)nst) {
                  stage = 0;
                  phase ^= 1u;
                }
              }
            }
            continue;
          }
          const uint32_t bytes = (uint32_t)(p.rows + p.BN) * (uint32_t)sg.kc * 2u;
          for (int tap = 0; tap < sg.taps; ++tap) {
            const int dy = sg.taps == 9 ? tap / 3 - 1 : 0;
            const int dx = sg.taps == 9 ? tap % 3 - 1 : 0;
            for (int c = 0; c < sg.nchunk; ++c) {
              mbar_wait(&empty[stage], phase ^ 1u);
              uint8_t* sa = stage_base + (size_t)stage * p.stage_bytes;
              uint8_t* sb = sa + p.a_region;
              mbar_arrive_expect_tx(&full[stage], bytes);
              tma_load_4d(&p.tmA[s], sa, &full[stage], sg.ch_off + c * sg.kc, x0 + dx, y0 + dy, b0);
              tma_load_3d(&p.tmW[s], sb, &full[stage], (tap * sg.nchunk + c) * sg.kc, n0, sg.per_image ? b0 : 0);
              if (++stage == (uint32_t)nst) {
                stage = 0;
                phase ^= 1u;
              }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16((uint32_t)p.BN);
      uint32_t stage = 0, phase = 0, it = 0, hs = 0, hphase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
        mbar_wait(&tempty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256u;
        uint32_t accumulate = 0;
        for (int s = 0; s < p.nseg; ++s) {
          const SegK sg = p.seg[s];
          if (sg.halo) {
            const uint32_t hw = (uint32_t)(p.TW + 2);
            for (int c = 0; c < sg.nchunk; ++c) {
              mbar_wait(&hfull[hs], hphase);
              tc_fence_after();
              const uint32_t hb = smem_u32(halo_base + (size_t)hs * kHaloBufBytes);
              const uint32_t wbytes = (uint32_t)p.BN * 128u;
              for (int t0 = 0; t0 < 9; t0 += p.hg) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t sb = smem_u32(stage_base + (size_t)stage * p.stage_bytes) + (uint32_t)p.a_region;
                for (int g = 0; g < p.hg; ++g) {
                  const int tap = t0 + g;
                  // tap (ky,kx): tile pixel (y,x) reads halo row (y+ky)*(TW+2) + (x+kx); with TW == 8 every 8-row
                  // core group is one tile row, (TW+2)*128 bytes apart.  The 128B swizzle is a function of the
                  // absolute shared-memory address bits (measured: the descriptor's base_offset must stay 0 for
                  // views that start off a 1024-byte boundary).
                  const uint32_t aaddr = hb + ((uint32_t)(tap / 3) * hw + (uint32_t)(tap % 3)) * 128u;
                  const uint64_t adesc = umma_smem_desc_sw128(aaddr, hw * 128u, 0u);
                  const uint64_t bdesc = umma_smem_desc(sb + (uint32_t)g * wbytes, 128u);
                  for (int k = 0; k < 4; ++k) {
                    umma_f16_ss(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, accumulate);
                    accumulate = 1;
                  }
                }
                umma_commit(&empty[stage]);
                if (++stage == (uint32_t)nst) {
                  stage = 0;
                  phase ^= 1u;
                }
              }
              umma_commit(&hempty[hs]);
              if (++hs == 2u) {
                hs = 0;
                hphase ^= 1u;
              }
            }
            continue;
          }
          const int chunks = sg.taps * sg.nchunk;
          const uint32_t row_bytes = (uint32_t)sg.kc * 2u;
          const int ksteps = sg.kc / 16;
          for (int c = 0; c < chunks; ++c) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(stage_base + (size_t)stage * p.stage_bytes);
            const uint64_t adesc = umma_smem_desc(sa, row_bytes);
            const uint64_t bdesc = umma_smem_desc(sa + (uint32_t)p.a_region, row_bytes);
            for (int k = 0; k < ksteps; ++k) {
              umma_f16_ss(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, accumulate);
              accumulate = 1;
            }
            umma_commit(&empty[stage]);
            if (++stage == (uint32_t)nst) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
        umma_commit(&tfull[acc]);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue warps
    // 8 warps: TMEM lane quadrant = warp % 4 (hardware rule), column half = (warp - 2) / 4.
    const int q = warp & 3;
    const int chalf = (warp - 2) >> 2;
    const int row_base = q * 32;
    const EpiK& e = p.e;
    const uint32_t stg = smem_u32(smem + (size_t)nst * p.stage_bytes + 1024 + (size_t)(warp - 2) * 4096);
    TileGeo tg;
    tg.init(p);

    if (EPI == CHB_EPI_PLAIN) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
        const int n_tile = tg.set_tile(p, tile);
        int b, y, x;
        const bool valid = tg.pixel(row_base + lane, b, y, x);
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256u;
        const int ch = p.BN >> 1;  // columns per warp-half (multiple of 8)
        int j = chalf * ch;
        const int jend = j + ch;
        const int n0 = n_tile * p.BN;
        const bool staged = e.o_sn == 1 && (e.o_ngroup <= 0 || (e.o_ngroup % 32) == 0);
        uint32_t ro[8], oo4[4], oo8[8];
        if (staged) {
          const long long osb = e.o_sb, osy = e.o_sy, osx = e.o_sx;
          auto o_elem = [=](int bb_, int yy, int xx) { return (long long)bb_ * osb + (long long)yy * osy + (long long)xx * osx; };
          if (e.out_dtype == CHB_F16) row_offsets<4>(lane, row_base, tg, 2, o_elem, oo4);
          else row_offsets<8>(lane, row_base, tg, 4, o_elem, oo8);
          if (e.res) {
            const long long rsb = e.r_sb, rsy = e.r_sy, rsx = e.r_sx;
            const int rsh = e.r_shift;
            row_offsets<8>(lane, row_base, tg, 4, [=](int bb_, int yy, int xx) {
              return (long long)bb_ * rsb + (long long)(yy >> rsh) * rsy + (long long)(xx >> rsh) * rsx;
            }, ro);
          }
        }
        const int bi = b < p.B ? b : p.B - 1;
        for (; j + 32 <= jend; j += 32) {
          if (staged && n0 + j + 32 <= p.N) {
            plain_block32<ACT>(e, taddr + (uint32_t)j, stg, lane, n0 + j, bi, ro, oo4, oo8);
          } else {
            plain_chunk<32, ACT>(p, e, taddr + (uint32_t)j, n0 + j, valid, b, y, x);
          }
        }
        for (; j + 16 <= jend; j += 16) plain_chunk<16, ACT>(p, e, taddr + (uint32_t)j, n0 + j, valid, b, y, x);
        for (; j + 8 <= jend; j += 8) plain_chunk<8, ACT>(p, e, taddr + (uint32_t)j, n0 + j, valid, b, y, x);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
      }
    } else {
      // MODULATE: tile columns [0, BN/2) are gamma, [BN/2, BN) beta of channels c0 .. c0 + BN/2.
      // Each warp owns 32 rows x cw channels, processed in units of 32 channels.  The x block of the NEXT unit
      // (possibly of the next tile) is requested from global memory before the current unit is computed, so its
      // latency overlaps the math instead of stalling the (few) epilogue warps.
      const int half_n = p.BN >> 1;
      const int cw = half_n >> 1;  // channels per warp-half (multiple of 32)
      const int units = cw >> 5;
      const int cs = e.chan_stride;
      const long long xsb = e.x_sb, xsy = e.x_sy, xsx = e.x_sx;
      const int xsh = e.x_shift;
      const int cl8 = lane & 7, cl4 = lane & 3;
      auto x_elem = [=](int bb_, int yy, int xx) {
        return (long long)bb_ * xsb + (long long)(yy >> xsh) * xsy + (long long)(xx >> xsh) * xsx;
      };
      const long long osb = e.o_sb, osy = e.o_sy, osx = e.o_sx;
      auto o_elem = [=](int bb_, int yy, int xx) { return (long long)bb_ * osb + (long long)yy * osy + (long long)xx * osx; };
      uint4 pf[8];
      uint32_t xo[8], xon[8];
      if ((int)blockIdx.x < total_tiles) {
        const int nt0 = tg.set_tile(p, blockIdx.x);
        row_offsets<8>(lane, row_base, tg, 4, x_elem, xon);
        gather_issue_o<8>(reinterpret_cast<const char*>(e.x + nt0 * half_n + chalf * cw) + cl8 * 16, xon, pf);
      }
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
        const int n_tile = tg.set_tile(p, tile);
        const int c0 = n_tile * half_n;
        const int nrow0 = n_tile * p.BN;
        int b, y, x;
        const bool valid = tg.pixel(row_base + lane, b, y, x);
        float nz = 0.f;
        if (valid && e.noise) nz = __ldg(e.noise + ((long long)b * p.W + x) * p.H + y);
        const float* ca = e.chan + c0;
#pragma unroll
        for (int k = 0; k < 8; ++k) xo[k] = xon[k];
        uint32_t oo[4];
        row_offsets<4>(lane, row_base, tg, 2, o_elem, oo);
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256u;
        for (int u = 0; u < units; ++u) {
          const int j = chalf * cw + 32 * u;
          // current x block: registers (global layout) -> staging -> one row per lane
          uint4 xr[8];
          gather_commit<8>(stg, lane, pf, xr);
          // request the next x block
          if (u + 1 < units) {
            gather_issue_o<8>(reinterpret_cast<const char*>(e.x + c0 + j + 32) + cl8 * 16, xo, pf);
          } else if (tile + (int)gridDim.x < total_tiles) {
            TileGeo tn = tg;
            const int ntn = tn.set_tile(p, tile + gridDim.x);
            row_offsets<8>(lane, row_base, tn, 4, x_elem, xon);
            gather_issue_o<8>(reinterpret_cast<const char*>(e.x + ntn * half_n + chalf * cw) + cl8 * 16, xon, pf);
          }
          uint4 hk[4];
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            const int jj = j + 16 * sub;
            float g[16], be[16];
            tmem_ld<16>(taddr + (uint32_t)jj, g);
            tmem_ld<16>(taddr + (uint32_t)(half_n + jj), be);
            float4 bg[4], bb[4], av[4], cv[4], nv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              bg[i] = __ldg(reinterpret_cast<const float4*>(e.bias + nrow0 + jj) + i);
              bb[i] = __ldg(reinterpret_cast<const float4*>(e.bias + nrow0 + half_n + jj) + i);
              av[i] = __ldg(reinterpret_cast<const float4*>(ca + jj) + i);
              cv[i] = __ldg(reinterpret_cast<const float4*>(ca + cs + jj) + i);
              nv[i] = __ldg(reinterpret_cast<const float4*>(ca + 2 * cs + jj) + i);
            }
            tmem_ld_fence(g);
            tmem_ld_fence(be);
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint4 xq = xr[4 * sub + i];
              const float o0 = modulate_elem<ACT>(__uint_as_float(xq.x), nz, av[i].x, cv[i].x, nv[i].x, g[4 * i] + bg[i].x, be[4 * i] + bb[i].x);
              const float o1 = modulate_elem<ACT>(__uint_as_float(xq.y), nz, av[i].y, cv[i].y, nv[i].y, g[4 * i + 1] + bg[i].y, be[4 * i + 1] + bb[i].y);
              const float o2 = modulate_elem<ACT>(__uint_as_float(xq.z), nz, av[i].z, cv[i].z, nv[i].z, g[4 * i + 2] + bg[i].z, be[4 * i + 2] + bb[i].z);
              const float o3 = modulate_elem<ACT>(__uint_as_float(xq.w), nz, av[i].w, cv[i].w, nv[i].w, g[4 * i + 3] + bg[i].w, be[4 * i + 3] + bb[i].w);
              pk[2 * i] = pack_h2(o0, o1);
              pk[2 * i + 1] = pack_h2(o2, o3);
            }
            hk[2 * sub] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            hk[2 * sub + 1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
          if (u + 1 == units) {
            // the accumulator has been fully read: hand the TMEM buffer back before the stores
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
          }
          scatter_o<4>(stg, lane, reinterpret_cast<char*>(reinterpret_cast<__half*>(e.out) + c0 + j) + cl4 * 16, oo, hk);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// SIMT checker kernel: same contract, one thread per output element, no tensor cores / TMA.
// Used only by tests (impl = CHB_IMPL_SIMT_DEBUG) to bisect the tcgen05 main loop from the epilogue.
// ------------------------------------------------------------------------------------------------
struct SimtSeg {
  const __half* a;
  long long a_sb, a_sy, a_sx;
  int ch_off, C, taps, per_image;
  const __half* w;
  long long w_sb;
};
struct SimtParams {
  SimtSeg seg[kMaxSeg];
  int nseg, B, H, W, N, BN, epi;
  EpiK e;
};

__device__ void plain_store_elem_rt(const EpiK& e, int b, int y, int x, int n, float acc) {
  switch (e.act) {
    case CHB_ACT_RELU: plain_store_elem<CHB_ACT_RELU>(e, b, y, x, n, acc); break;
    case CHB_ACT_LRELU: plain_store_elem<CHB_ACT_LRELU>(e, b, y, x, n, acc); break;
    case CHB_ACT_TANH: plain_store_elem<CHB_ACT_TANH>(e, b, y, x, n, acc); break;
    default: plain_store_elem<CHB_ACT_NONE>(e, b, y, x, n, acc); break;
  }
}

__device__ float simt_dot(const SimtParams& p, int b, int y, int x, int nrow) {
  float acc = 0.f;
  for (int s = 0; s < p.nseg; ++s) {
    const SimtSeg& sg = p.seg[s];
    const int K = sg.taps * sg.C;
    const __half* wrow = sg.w + (sg.per_image ? (long long)b * (sg.w_sb > 0 ? sg.w_sb : (long long)p.e.nrows * K) : 0) + (long long)nrow * K;
    for (int tap = 0; tap < sg.taps; ++tap) {
      const int yy = y + (sg.taps == 9 ? tap / 3 - 1 : 0);
      const int xx = x + (sg.taps == 9 ? tap % 3 - 1 : 0);
      if (yy < 0 || yy >= p.H || xx < 0 || xx >= p.W) continue;
      const __half* ap = sg.a + (long long)b * sg.a_sb + (long long)yy * sg.a_sy + (long long)xx * sg.a_sx + sg.ch_off;
      const __half* wp = wrow + tap * sg.C;
      for (int c = 0; c < sg.C; ++c) acc = fmaf(__half2float(ap[c]), __half2float(wp[c]), acc);
    }
  }
  return acc;
}

__global__ void conv_simt_kernel(const SimtParams p) {
  const long long cols = p.epi == CHB_EPI_PLAIN ? p.N : p.N / 2;
  const long long total = (long long)p.B * p.H * p.W * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i % cols);
    long long pix = i / cols;
    const int x = (int)(pix % p.W);
    pix /= p.W;
    const int y = (int)(pix % p.H);
    const int b = (int)(pix / p.H);
    if (p.epi == CHB_EPI_PLAIN) {
      plain_store_elem_rt(p.e, b, y, x, n, simt_dot(p, b, y, x, n));
    } else {
      const int half_n = p.BN / 2;
      const int t = n / half_n, j = n % half_n;
      const int grow = t * p.BN + j, brow = grow + half_n;
      const float g = simt_dot(p, b, y, x, grow) + __ldg(p.e.bias + grow);
      const float be = simt_dot(p, b, y, x, brow) + __ldg(p.e.bias + brow);
      const float nz = p.e.noise ? __ldg(p.e.noise + ((long long)b * p.W + x) * p.H + y) : 0.f;
      const float xv = __ldg(p.e.x + (long long)b * p.e.x_sb + (long long)(y >> p.e.x_shift) * p.e.x_sy +
                             (long long)(x >> p.e.x_shift) * p.e.x_sx + n);
      const float* cp = reinterpret_cast<const float*>(p.e.chan);
      const float a = __ldg(cp + n), c = __ldg(cp + p.e.chan_stride + n), nvv = __ldg(cp + 2 * p.e.chan_stride + n);
      const float o = apply_act(fmaf(fmaf(xv, a, fmaf(nz, nvv, c)), 1.f + g, be), p.e.act);
      reinterpret_cast<__half*>(p.e.out)[(long long)b * p.e.o_sb + (long long)y * p.e.o_sy + (long long)x * p.e.o_sx + n] =
          __float2half_rn(o);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Host side: tensor maps, plans, launches
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}

static int encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_b,
                      const cuuint32_t* box, int kc) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return CHB_ERR_CUDA;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUtensorMapSwizzle sw = kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_b, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof buf,
             "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu,%llu,%llu,%llu box %u,%u,%u,%u stride0 %llu", (int)r,
             rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
             (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], box[1], box[2], rank > 3 ? box[3] : 0,
             (unsigned long long)strides_b[0]);
    set_error(buf);
    return CHB_ERR_CUDA;
  }
  return CHB_OK;
}

static int validate_desc(const chb_conv_desc& d) {
  char buf[256];
#define CHB_REQUIRE(cond, msg)                                      \
  if (!(cond)) {                                                    \
    snprintf(buf, sizeof buf, "chb_conv: invalid descriptor: %s", msg); \
    set_error(buf);                                                 \
    return CHB_ERR_ARG;                                             \
  }
  CHB_REQUIRE(d.B > 0 && d.H > 0 && d.W > 0, "B,H,W must be positive");
  CHB_REQUIRE(d.TW > 0 && d.TH > 0 && d.TB > 0 && d.TW * d.TH * d.TB <= 128, "tile must have 1..128 rows");
  CHB_REQUIRE(d.TW <= 256 && d.TH <= 256 && d.TB <= 256, "tile dims <= 256");
  CHB_REQUIRE(d.nseg >= 1 && d.nseg <= kMaxSeg, "1..3 segments");
  CHB_REQUIRE(d.BN >= 16 && d.BN <= 256 && d.BN % 16 == 0, "BN in 16..256, multiple of 16");
  CHB_REQUIRE(d.Nrows > 0 && d.Nrows % d.BN == 0, "Nrows must be a positive multiple of BN");
  CHB_REQUIRE(d.N > 0 && d.N <= d.Nrows, "0 < N <= Nrows");
  CHB_REQUIRE(d.epi == CHB_EPI_PLAIN || d.epi == CHB_EPI_MODULATE, "unknown epilogue");
  CHB_REQUIRE(d.out != nullptr, "out is NULL");
  for (int s = 0; s < d.nseg; ++s) {
    const chb_conv_seg& g = d.seg[s];
    CHB_REQUIRE(g.a && g.w, "segment pointers NULL");
    CHB_REQUIRE(g.taps == 9 || g.taps == 1, "taps must be 9 or 1");
    CHB_REQUIRE(g.C == 32 || (g.C > 0 && g.C % 64 == 0), "segment C must be 32 or a multiple of 64");
    CHB_REQUIRE(g.ch_off >= 0 && g.ch_off + g.C <= g.Ca, "channel window exceeds tensor");
    CHB_REQUIRE((g.a_sx % 8) == 0 && (g.a_sy % 8) == 0 && (g.a_sb % 8) == 0 && (g.ch_off % 8) == 0,
                "activation strides must be multiples of 8 elements (16 B)");
    CHB_REQUIRE(!g.per_image || d.TB == 1, "per-image weights need TB == 1");
    CHB_REQUIRE((reinterpret_cast<uintptr_t>(g.a) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.w) & 15) == 0,
                "operands must be 16-byte aligned");
  }
  if (d.epi == CHB_EPI_MODULATE) {
    CHB_REQUIRE(d.x && d.chan && d.bias, "modulate epilogue needs x, chan and bias");
    CHB_REQUIRE(d.BN % 128 == 0 && d.N == d.Nrows, "modulate epilogue needs BN % 128 == 0 and N == Nrows");
    CHB_REQUIRE(d.act != CHB_ACT_TANH, "modulate epilogue supports none / relu / lrelu");
    CHB_REQUIRE(d.o_sn == 1 || d.o_sn == 0, "modulate output is channels-last");
  }
#undef CHB_REQUIRE
  return CHB_OK;
}

int build_conv_plan(const chb_conv_desc& d, ConvPlan* plan) {
  int rc = validate_desc(d);
  if (rc != CHB_OK) return rc;
  memset(plan, 0, sizeof(*plan));
  plan->desc = d;
  ConvKParams& k = plan->kp;
  k.nseg = d.nseg;
  k.B = d.B; k.H = d.H; k.W = d.W;
  k.TW = d.TW; k.TH = d.TH; k.TB = d.TB;
  k.rows = d.TW * d.TH * d.TB;
  k.tiles_x = (d.W + d.TW - 1) / d.TW;
  k.tiles_y = (d.H + d.TH - 1) / d.TH;
  const int tiles_b = (d.B + d.TB - 1) / d.TB;
  k.m_tiles = k.tiles_x * k.tiles_y * tiles_b;
  k.n_tiles = d.Nrows / d.BN;
  k.BN = d.BN;
  k.N = d.N;
  k.stage_bytes = kATileBytes + ((d.BN * 128 + 1023) / 1024) * 1024;
  // Halo path (on by default; CHB_HALO=0 turns it off for A/B comparisons): 3x3 segments with 64-channel chunks on
  // 8-wide single-image tiles load a (TH+2)x(TW+2) halo tile once per channel chunk and feed all nine taps from it.
  int halo_mode = 1;
  if (const char* hv = getenv("CHB_HALO")) halo_mode = atoi(hv);
  k.halo_any = 0;
  k.halo_bo = 0;
  for (int s = 0; s < d.nseg; ++s) {
    const chb_conv_seg& g = d.seg[s];
    k.seg[s].halo = (halo_mode > 0 && g.taps == 9 && g.C % 64 == 0 && d.TW == 8 && d.TB == 1 && d.TH <= 16) ? 1 : 0;
    k.halo_any |= k.seg[s].halo;
  }
  bool all_halo = true;
  for (int s = 0; s < d.nseg; ++s) all_halo = all_halo && k.seg[s].halo;
  k.a_region = all_halo ? 0 : kATileBytes;
  k.hg = (k.halo_any && d.BN <= 128) ? 3 : 1;
  k.stage_bytes = k.a_region + ((d.BN * 128 * k.hg + 1023) / 1024) * 1024;
  const int budget = kSmemBudget - (k.halo_any ? 2 * kHaloBufBytes : 0);
  k.nstages = budget / k.stage_bytes;
  if (k.nstages > kMaxStages) k.nstages = kMaxStages;
  plan->smem_bytes = k.nstages * k.stage_bytes + 1024 /*align slack*/ + 1024 /*barriers*/ + kEpilogueWarps * 4096 +
                     (k.halo_any ? 2 * kHaloBufBytes : 0);
  long long ktotal = 0;
  for (int s = 0; s < d.nseg; ++s) {
    const chb_conv_seg& g = d.seg[s];
    const int kc = g.C == 32 ? 32 : 64;
    k.seg[s].taps = g.taps;
    k.seg[s].kc = kc;
    k.seg[s].nchunk = g.C / kc;
    k.seg[s].ch_off = g.ch_off;
    k.seg[s].per_image = g.per_image;
    ktotal += (long long)g.taps * g.C;
    {
      cuuint64_t dims[4] = {(cuuint64_t)g.Ca, (cuuint64_t)d.W, (cuuint64_t)d.H, (cuuint64_t)d.B};
      cuuint64_t str[3] = {(cuuint64_t)g.a_sx * 2, (cuuint64_t)g.a_sy * 2, (cuuint64_t)g.a_sb * 2};
      cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)d.TW, (cuuint32_t)d.TH, (cuuint32_t)d.TB};
      rc = encode_map(&k.tmA[s], g.a, 4, dims, str, box, kc);
      if (rc != CHB_OK) return rc;
      if (k.seg[s].halo) {
        cuuint32_t hbox[4] = {(cuuint32_t)kc, (cuuint32_t)d.TW + 2, (cuuint32_t)d.TH + 2, 1};
        rc = encode_map(&k.tmH[s], g.a, 4, dims, str, hbox, kc);
        if (rc != CHB_OK) return rc;
      }
    }
    {
      const cuuint64_t K = (cuuint64_t)g.taps * g.C;
      cuuint64_t dims[3] = {K, (cuuint64_t)d.Nrows, (cuuint64_t)(g.per_image ? d.B : 1)};
      const cuuint64_t img_stride = (g.per_image && g.w_sb > 0) ? (cuuint64_t)g.w_sb * 2 : K * 2 * (cuuint64_t)d.Nrows;
      cuuint64_t str[2] = {K * 2, img_stride};
      cuuint32_t box[3] = {(cuuint32_t)kc, (cuuint32_t)d.BN, 1};
      rc = encode_map(&k.tmW[s], g.w, 3, dims, str, box, kc);
      if (rc != CHB_OK) return rc;
    }
  }
  EpiK& e = k.e;
  e.act = d.act; e.bias = d.bias; e.bias_per_image = d.bias_per_image; e.nrows = d.Nrows;
  e.out = d.out; e.out_dtype = d.out_dtype;
  e.o_sb = d.o_sb; e.o_sy = d.o_sy; e.o_sx = d.o_sx; e.o_sn = d.o_sn;
  e.o_ngroup = d.o_ngroup; e.o_sgroup = d.o_sgroup;
  e.res = d.res; e.r_sb = d.r_sb; e.r_sy = d.r_sy; e.r_sx = d.r_sx; e.r_shift = d.r_shift;
  e.x = d.x; e.x_sb = d.x_sb; e.x_sy = d.x_sy; e.x_sx = d.x_sx; e.x_shift = d.x_shift;
  e.noise = d.noise;
  e.chan = d.chan;
  e.chan_stride = d.N / 2;
  const int total = k.m_tiles * k.n_tiles;
  const int sms = device_sm_count();
  plan->grid = total < sms ? total : sms;
  plan->flops = 2.0 * 128.0 * d.BN * (double)ktotal * (double)total;
  return CHB_OK;
}

typedef void (*ConvKernelFn)(const ConvKParams);

static ConvKernelFn pick_kernel(int epi, int act) {
  if (epi == CHB_EPI_PLAIN) {
    switch (act) {
      case CHB_ACT_RELU: return conv_igemm_kernel<CHB_EPI_PLAIN, CHB_ACT_RELU>;
      case CHB_ACT_LRELU: return conv_igemm_kernel<CHB_EPI_PLAIN, CHB_ACT_LRELU>;
      case CHB_ACT_TANH: return conv_igemm_kernel<CHB_EPI_PLAIN, CHB_ACT_TANH>;
      default: return conv_igemm_kernel<CHB_EPI_PLAIN, CHB_ACT_NONE>;
    }
  }
  switch (act) {
    case CHB_ACT_LRELU: return conv_igemm_kernel<CHB_EPI_MODULATE, CHB_ACT_LRELU>;
    case CHB_ACT_RELU: return conv_igemm_kernel<CHB_EPI_MODULATE, CHB_ACT_RELU>;
    default: return conv_igemm_kernel<CHB_EPI_MODULATE, CHB_ACT_NONE>;
  }
}

static int ensure_smem_attr() {
  static std::once_flag once;
  static cudaError_t err = cudaSuccess;
  std::call_once(once, [] {
    for (int epi = 0; epi < 2 && err == cudaSuccess; ++epi)
      for (int act = 0; act < 4 && err == cudaSuccess; ++act)
        err = cudaFuncSetAttribute(pick_kernel(epi, act), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   227 * 1024);
  });
  if (err != cudaSuccess) {
    set_error(std::string("cudaFuncSetAttribute(max dynamic smem) failed: ") + cudaGetErrorString(err));
    return CHB_ERR_CUDA;
  }
  return CHB_OK;
}

int launch_conv_plan(const ConvPlan& plan, int impl, cudaStream_t stream) {
  if (impl == CHB_IMPL_SIMT_DEBUG) {
    SimtParams sp;
    memset(&sp, 0, sizeof sp);
    const chb_conv_desc& d = plan.desc;
    sp.nseg = d.nseg; sp.B = d.B; sp.H = d.H; sp.W = d.W; sp.N = d.N; sp.BN = d.BN; sp.epi = d.epi;
    sp.e = plan.kp.e;
    for (int s = 0; s < d.nseg; ++s) {
      sp.seg[s].a = reinterpret_cast<const __half*>(d.seg[s].a);
      sp.seg[s].a_sb = d.seg[s].a_sb; sp.seg[s].a_sy = d.seg[s].a_sy; sp.seg[s].a_sx = d.seg[s].a_sx;
      sp.seg[s].ch_off = d.seg[s].ch_off; sp.seg[s].C = d.seg[s].C; sp.seg[s].taps = d.seg[s].taps;
      sp.seg[s].per_image = d.seg[s].per_image;
      sp.seg[s].w = reinterpret_cast<const __half*>(d.seg[s].w);
      sp.seg[s].w_sb = d.seg[s].w_sb;
    }
    conv_simt_kernel<<<device_sm_count() * 8, 256, 0, stream>>>(sp);
  } else {
    int rc = ensure_smem_attr();
    if (rc != CHB_OK) return rc;
    pick_kernel(plan.desc.epi, plan.desc.act)<<<plan.grid, kConvThreads, plan.smem_bytes, stream>>>(plan.kp);
  }
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    set_error(std::string("conv launch failed: ") + cudaGetErrorString(err));
    return CHB_ERR_CUDA;
  }
  return CHB_OK;
}

}  // namespace chb

extern "C" {

int chb_version(void) { return 100; }
const char* chb_last_error(void) { return chb::last_error_cstr(); }

int chb_check_device(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    chb::set_error("no CUDA device");
    return CHB_ERR_CUDA;
  }
  if (major != 10) {
    chb::set_error("ctrlhair_b200 kernels are built for sm_100a only; current device is not compute capability 10.x");
    return CHB_ERR_ARCH;
  }
  return CHB_OK;
}

int chb_conv_run(const chb_conv_desc* d, int impl, void* stream) {
  if (!d) {
    chb::set_error("chb_conv_run: NULL descriptor");
    return CHB_ERR_ARG;
  }
  int rc = chb_check_device();
  if (rc != CHB_OK) return rc;
  chb::ConvPlan plan;
  rc = chb::build_conv_plan(*d, &plan);
  if (rc != CHB_OK) return rc;
  return chb::launch_conv_plan(plan, impl, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
