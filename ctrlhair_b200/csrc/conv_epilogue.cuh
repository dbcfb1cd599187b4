// Epilogue of the implicit-GEMM conv kernel: helpers shared with the SIMT checker kernel, the per-warp staging
// block, and the epilogue warp role (TMEM -> registers -> fused math -> global).
#pragma once
#include <cuda_fp16.h>

#include "conv_igemm.cuh"
#include "ptx_sm100.cuh"

namespace chb {

// Shared-memory carve-up of one CTA (see conv_igemm.cu for the layout) and the persistent tile schedule.
struct Smem {
  uint8_t* stage_base;   // pipeline stages (A tile | weight chunks)
  uint64_t *full, *empty, *tfull, *tempty, *hfull, *hempty, *wbar;
  uint32_t* tmem_slot;
  uint32_t* ks_flag;     // split-K: "this CTA arrived last on the current tile", broadcast to the epilogue warps
  uint8_t* stg_base;     // 8 x 4 KB epilogue staging blocks
  uint8_t* halo_base;    // halo tiles (ring of p.nhalo buffers)
  uint8_t* wstat_base;   // resident weights (weight-stationary mode)
};

// t-th tile of this CTA, or -1.  Normal mode: round robin over all (n_tile, m_tile).  Weight-stationary mode: the
// CTA keeps one n_tile for its whole life (its weights stay in shared memory) and strides over the m tiles.
template <bool WSTAT>
__device__ __forceinline__ int sched_tile(const ConvKParams& p, uint32_t t) {
  if (WSTAT) {
    const int n_tile = (int)blockIdx.x % p.n_tiles;
    const long long m = (long long)((int)blockIdx.x / p.n_tiles) + (long long)t * ((int)gridDim.x / p.n_tiles);
    return m < p.m_tiles ? n_tile * p.m_tiles + (int)m : -1;
  }
  const long long tile = (long long)blockIdx.x + (long long)t * gridDim.x;
  return tile < (long long)p.tiles_mn * p.ksplit ? (int)tile : -1;
}
// Split-K: a scheduled id is (split, output tile); reduces `tile` to the output tile and returns the split.
__device__ __forceinline__ int split_of_tile(const ConvKParams& p, int& tile) {
  if (p.ksplit <= 1) return 0;
  const int ks = tile / p.tiles_mn;
  tile -= ks * p.tiles_mn;
  return ks;
}
// channel chunks [cb, ce) of a segment with `nchunk` chunks that split `ks` accumulates
__device__ __forceinline__ void split_chunks(const ConvKParams& p, int ks, int nchunk, int& cb, int& ce) {
  if (p.ksplit <= 1) { cb = 0; ce = nchunk; return; }
  cb = nchunk * ks / p.ksplit;
  ce = nchunk * (ks + 1) / p.ksplit;
}

// ------------------------------------------------------------------------------------------------
// Epilogue math shared by the tcgen05 kernel and the SIMT checker kernel.
// ------------------------------------------------------------------------------------------------
// tanh x = 1 - 2 / (exp(2x) + 1) on ex2.approx / rcp.approx: within ~3e-7 ABSOLUTE of tanhf over the whole range (exact
// limits at +-inf), in 6 instructions instead of libdevice's ~25 with a branch.  The style encoder's last layer applies
// it to 268 M values per 32 faces; that launch is MMA-bound (1.25 PFLOP/s), so this only shortens the epilogue that
// runs beside the MMAs (2 % of the encoder).
__device__ __forceinline__ float tanh_fast(float x) { return 1.f - __fdividef(2.f, __expf(2.f * x) + 1.f); }

template <int ACT>
__device__ __forceinline__ float act_t(float v) {
  if (ACT == CHB_ACT_RELU) return fmaxf(v, 0.f);
  if (ACT == CHB_ACT_LRELU) return fmaxf(v, 0.2f * v);
  if (ACT == CHB_ACT_TANH) return tanh_fast(v);
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
    const __half hi = __float2half_rn(v);
    reinterpret_cast<__half*>(e.out)[off] = hi;
    if (e.split) reinterpret_cast<__half*>(e.out)[off + e.o_lo] = __float2half_rn(v - __half2float(hi));
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
  // fast geometry (p.fast): 16 x 8 tiles, power-of-two tile grid -> shifts only
  __device__ __forceinline__ int set_tile_fast(const ConvKParams& p, int tile) {
    const int n_tile = p.n_tiles == 1 ? 0 : tile / p.m_tiles;
    const int m = tile - n_tile * p.m_tiles;
    x0 = (m & (p.tiles_x - 1)) << 3;
    y0 = ((m >> p.tx_sh) & (p.tiles_y - 1)) << 4;
    b0 = m >> (p.tx_sh + p.ty_sh);
    return n_tile;
  }
  __device__ __forceinline__ void pixel_fast(int m, int& b, int& y, int& x) const {
    b = b0; y = y0 + (m >> 3); x = x0 + (m & 7);
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

// Streaming 16-byte load that does not allocate in L1: the x / residual streams are read once, keeping them out of
// L1 leaves the per-channel constants (bias, BN fold, noise scale) resident there.
__device__ __forceinline__ uint4 ldg_stream(const void* ptr) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr));
  return v;
}
__device__ __forceinline__ float4 ldg_keep(const float4* ptr) {
  float4 v;
  asm volatile("ld.global.nc.L1::evict_last.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(ptr));
  return v;
}

// Packed fp32 pairs (FFMA2 / FADD2 / FMUL2 on sm_100): two lanes of fp32 math per issue slot.
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// Row offsets (16-byte units) of the rows a lane touches in the CH16-lanes-per-row arrangement of its warp's 32 tile
// rows: k-th access -> tile row m = q*32 + lane/CH16 + (32/CH16)*k (ty = m >> 3, tx = m & 7), tensor read at
// (y >> sh, x >> sh).  Generic geometry keeps one precomputed offset per access (~0u = row outside the image); fast
// geometry is affine in k, so it keeps a per-tile base and three strides instead of 4-8 live registers per tensor:
//   CH16 = 8: ty = q*4 + (k >> 1), tx = lane/8 + 4*(k & 1)   ->  base + (k>>2)*syB + ((k>>1)&1)*syA + (k&1)*sxs
//             with (syA, syB) = (sy, 2 sy) for sh = 0 and (0, sy) for sh = 1, sxs = (4 >> sh) * sx
//   CH16 = 4: ty = q*4 + k,        tx = lane/4               ->  base + k*sy            (sh = 0 only: outputs)
template <int CH16, bool FAST>
struct RowOff {
  uint32_t o[CH16];
  __device__ __forceinline__ uint32_t get(int k) const { return o[k]; }
  __device__ __forceinline__ bool valid(int k) const { return o[k] != 0xFFFFFFFFu; }
};
template <>
struct RowOff<8, true> {
  uint32_t lane_part, syA, syB, sxs, base;
  __device__ __forceinline__ void setup(int lane, int q, int sh, uint32_t sy16, uint32_t sx16) {
    lane_part = (uint32_t)((q * 4) >> sh) * sy16 + (uint32_t)((lane >> 3) >> sh) * sx16;
    syA = sh ? 0u : sy16;
    syB = sh ? sy16 : 2u * sy16;
    sxs = (uint32_t)(4 >> sh) * sx16;
    base = lane_part;
  }
  __device__ __forceinline__ void retile(uint32_t tile16) { base = tile16 + lane_part; }
  __device__ __forceinline__ uint32_t get(int k) const {
    return base + (uint32_t)(k >> 2) * syB + (uint32_t)((k >> 1) & 1) * syA + (uint32_t)(k & 1) * sxs;
  }
  __device__ __forceinline__ bool valid(int) const { return true; }
};
template <>
struct RowOff<4, true> {
  uint32_t lane_part, sy, base;
  __device__ __forceinline__ void setup(int lane, int q, int /*sh*/, uint32_t sy16, uint32_t sx16) {
    lane_part = (uint32_t)(q * 4) * sy16 + (uint32_t)(lane >> 2) * sx16;
    sy = sy16;
    base = lane_part;
  }
  __device__ __forceinline__ void retile(uint32_t tile16) { base = tile16 + lane_part; }
  __device__ __forceinline__ uint32_t get(int k) const { return base + (uint32_t)k * sy; }
  __device__ __forceinline__ bool valid(int) const { return true; }
};
__device__ __forceinline__ uint32_t tile_base16(const TileGeo& tg, long long sb, long long sy, long long sx, int sh,
                                                int elem_bytes) {
  return (uint32_t)((((long long)tg.b0 * sb + (long long)(tg.y0 >> sh) * sy + (long long)(tg.x0 >> sh) * sx) *
                     elem_bytes) >> 4);
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
template <int CH16, bool FAST>
__device__ __forceinline__ void gather_issue_o(const char* base, const RowOff<CH16, FAST>& ro, uint4 (&g)[CH16]) {
#pragma unroll
  for (int k = 0; k < CH16; ++k) {
    if (FAST) {
      g[k] = ldg_stream(base + ((size_t)ro.get(k) << 4));
    } else {
      g[k] = make_uint4(0u, 0u, 0u, 0u);
      if (ro.valid(k)) g[k] = ldg_stream(base + ((size_t)ro.get(k) << 4));
    }
  }
}
template <int CH16, bool FAST>
__device__ __forceinline__ void scatter_o(uint32_t stg, int lane, char* base, const RowOff<CH16, FAST>& ro,
                                          const uint4 (&regs)[CH16]) {
  constexpr int RPI = 32 / CH16;
  const int cl = lane % CH16, r0 = lane / CH16;
#pragma unroll
  for (int c = 0; c < CH16; ++c) sts128(stg + stg_off<CH16>(lane, c), regs[c]);
  __syncwarp();
#pragma unroll
  for (int k = 0; k < CH16; ++k) {
    const uint4 v = lds128(stg + stg_off<CH16>(r0 + RPI * k, cl));
    if (FAST || ro.valid(k)) *reinterpret_cast<uint4*>(base + ((size_t)ro.get(k) << 4)) = v;
  }
  __syncwarp();
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// fp16 hi + lo split of two fp32 values: hi = fp16(v), lo = fp16(v - hi)  (v is then known to ~2^-22 relative)
__device__ __forceinline__ void pack_h2_split(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// 32 accumulator columns of the PLAIN epilogue through the staging block (channels-last output, full block valid).
// ro / oo4 / oo8: per-tile row offsets (16-byte units) of the residual and of the output in the 8- and 4-lanes-per-row
// arrangements; bi = image of this lane's own row (for per-image bias).
// breg (fast geometry only): this block's 32 bias values, one per lane, fetched before the accumulator wait; they
// are broadcast with warp shuffles instead of being re-loaded (and waited for) inside the block.
template <int ACT, bool FAST>
__device__ __forceinline__ void plain_block32(const EpiK& e, uint32_t taddr, uint32_t stg, int lane, int n, int bi,
                                              const RowOff<8, FAST>& ro, const RowOff<4, FAST>& oo4,
                                              const RowOff<8, FAST>& oo8, float breg = 0.f) {
  float v[32];
  tmem_ld<32>(taddr, v);
  uint4 rr[8];
  if (e.res) {
    uint4 g[8];
    gather_issue_o<8, FAST>(reinterpret_cast<const char*>(e.res + n) + (lane & 7) * 16, ro, g);
    gather_commit<8>(stg, lane, g, rr);
  }
  tmem_ld_fence(v);
  if (FAST) {
    if (e.bias) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] += __shfl_sync(0xffffffffu, breg, i);
    }
  } else if (e.bias) {
    const float4* bp = reinterpret_cast<const float4*>(e.bias + (e.bias_per_image ? (long long)bi * e.nrows : 0) + n);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 t = ldg_keep(bp + i);
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
    char* obase = reinterpret_cast<char*>(reinterpret_cast<__half*>(e.out) + noff) + (lane & 3) * 16;
    if (e.split) {
      uint4 pk[4], pl[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        pack_h2_split(v[8 * i], v[8 * i + 1], pk[i].x, pl[i].x);
        pack_h2_split(v[8 * i + 2], v[8 * i + 3], pk[i].y, pl[i].y);
        pack_h2_split(v[8 * i + 4], v[8 * i + 5], pk[i].z, pl[i].z);
        pack_h2_split(v[8 * i + 6], v[8 * i + 7], pk[i].w, pl[i].w);
      }
      scatter_o<4, FAST>(stg, lane, obase, oo4, pk);
      scatter_o<4, FAST>(stg, lane, obase + e.o_lo * 2, oo4, pl);
    } else {
      uint4 pk[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        pk[i] = make_uint4(pack_h2(v[8 * i], v[8 * i + 1]), pack_h2(v[8 * i + 2], v[8 * i + 3]),
                           pack_h2(v[8 * i + 4], v[8 * i + 5]), pack_h2(v[8 * i + 6], v[8 * i + 7]));
      scatter_o<4, FAST>(stg, lane, obase, oo4, pk);
    }
  } else {
    uint4 pk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
      pk[i] = make_uint4(__float_as_uint(v[4 * i]), __float_as_uint(v[4 * i + 1]), __float_as_uint(v[4 * i + 2]),
                         __float_as_uint(v[4 * i + 3]));
    scatter_o<8, FAST>(stg, lane, reinterpret_cast<char*>(reinterpret_cast<float*>(e.out) + noff) + (lane & 7) * 16, oo8, pk);
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
      uint4* ol = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(e.out) + off + e.o_lo);
#pragma unroll
      for (int i = 0; i < NC / 8; ++i) {
        uint32_t pk[4], pl[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) pack_h2_split(v[8 * i + 2 * k], v[8 * i + 2 * k + 1], pk[k], pl[k]);
        op[i] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        if (e.split) ol[i] = make_uint4(pl[0], pl[1], pl[2], pl[3]);
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
// MODULATE math on 16 channels: out = act((x*a + (nz*nv + c)) * (1 + gamma + bias_g) + (beta + bias_b)), packed fp32
// pairs (FFMA2/FADD2/FMUL2), fp16 output.  g / be: accumulator columns, xr: 4 x uint4 of fp32 x values, cst: the
// five per-channel constant vectors of these 16 channels.
// ------------------------------------------------------------------------------------------------
template <int ACT, bool SPLIT>
__device__ __forceinline__ void modulate16(const float (&g)[16], const float (&be)[16], const uint4* xr, float nz,
                                           const float4 (&bg)[4], const float4 (&bb)[4], const float4 (&av)[4],
                                           const float4 (&cv)[4], const float4 (&nv)[4], uint4& h0, uint4& h1,
                                           uint4& l0, uint4& l1) {
  const uint64_t nz2 = pk2(nz, nz);
  const uint64_t slope2 = pk2(0.2f, 0.2f);
  uint32_t pk[8], pl[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint4 xq = xr[i];
    const uint64_t x01 = pk2(__uint_as_float(xq.x), __uint_as_float(xq.y));
    const uint64_t x23 = pk2(__uint_as_float(xq.z), __uint_as_float(xq.w));
    // xn = x*a + (nz*nv + c)
    const uint64_t xn01 = fma2(x01, pk2(av[i].x, av[i].y), fma2(nz2, pk2(nv[i].x, nv[i].y), pk2(cv[i].x, cv[i].y)));
    const uint64_t xn23 = fma2(x23, pk2(av[i].z, av[i].w), fma2(nz2, pk2(nv[i].z, nv[i].w), pk2(cv[i].z, cv[i].w)));
    // gamma + bias_g, beta + bias_b
    const uint64_t g01 = add2(pk2(g[4 * i], g[4 * i + 1]), pk2(bg[i].x, bg[i].y));
    const uint64_t g23 = add2(pk2(g[4 * i + 2], g[4 * i + 3]), pk2(bg[i].z, bg[i].w));
    const uint64_t b01 = add2(pk2(be[4 * i], be[4 * i + 1]), pk2(bb[i].x, bb[i].y));
    const uint64_t b23 = add2(pk2(be[4 * i + 2], be[4 * i + 3]), pk2(bb[i].z, bb[i].w));
    // xn * (1 + gamma) + beta = xn*gamma + (xn + beta)
    uint64_t o01 = fma2(xn01, g01, add2(xn01, b01));
    uint64_t o23 = fma2(xn23, g23, add2(xn23, b23));
    float o0, o1, o2, o3;
    if (ACT == CHB_ACT_LRELU) {
      float s0, s1, s2, s3;
      upk2(mul2(o01, slope2), s0, s1);
      upk2(mul2(o23, slope2), s2, s3);
      upk2(o01, o0, o1);
      upk2(o23, o2, o3);
      o0 = fmaxf(o0, s0); o1 = fmaxf(o1, s1); o2 = fmaxf(o2, s2); o3 = fmaxf(o3, s3);
    } else {
      upk2(o01, o0, o1);
      upk2(o23, o2, o3);
      o0 = act_t<ACT>(o0); o1 = act_t<ACT>(o1); o2 = act_t<ACT>(o2); o3 = act_t<ACT>(o3);
    }
    if (SPLIT) {
      pack_h2_split(o0, o1, pk[2 * i], pl[2 * i]);
      pack_h2_split(o2, o3, pk[2 * i + 1], pl[2 * i + 1]);
    } else {
      pk[2 * i] = pack_h2(o0, o1);
      pk[2 * i + 1] = pack_h2(o2, o3);
    }
  }
  h0 = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  h1 = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  if (SPLIT) {
    l0 = make_uint4(pl[0], pl[1], pl[2], pl[3]);
    l1 = make_uint4(pl[4], pl[5], pl[6], pl[7]);
  }
}

// ------------------------------------------------------------------------------------------------
// Epilogue warp role.  FAST: p.fast geometry (full 16 x 8 tiles, affine addresses, no bounds predicates).
// ------------------------------------------------------------------------------------------------
// Sum of the ksplit partial accumulators of this thread's row, columns [j0, j1), in split order (fixed: the result does
// not depend on which CTA does it), written back to TMEM.  CW columns per step, the loads of G splits in flight at once.
template <int CW, int G>
__device__ __forceinline__ void ksplit_sum(const ConvKParams& p, const float* base, size_t tile_elems, uint32_t taddr,
                                           int j0, int j1) {
  constexpr int V = CW / 4;
  for (int j = j0; j < j1; j += CW) {   // (j1 - j0) is a multiple of CW
    float a[CW];
#pragma unroll
    for (int i = 0; i < CW; ++i) a[i] = 0.f;
    for (int s0 = 0; s0 < p.ksplit; s0 += G) {
      float4 x[G][V];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float4* src = reinterpret_cast<const float4*>(base + (size_t)(s0 + g) * tile_elems) + (j >> 2) * 128;
#pragma unroll
        for (int v = 0; v < V; ++v)
          x[g][v] = s0 + g < p.ksplit ? __ldcg(src + v * 128) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int g = 0; g < G; ++g)
#pragma unroll
        for (int v = 0; v < V; ++v) {   // (adding the zeros of an absent split changes nothing)
          a[4 * v] += x[g][v].x; a[4 * v + 1] += x[g][v].y; a[4 * v + 2] += x[g][v].z; a[4 * v + 3] += x[g][v].w;
        }
    }
#pragma unroll
    for (int i = 0; i < CW; i += 8) {
      const float o[8] = {a[i], a[i + 1], a[i + 2], a[i + 3], a[i + 4], a[i + 5], a[i + 6], a[i + 7]};
      tmem_st8(taddr + (uint32_t)(j + i), o);
    }
  }
}

// Split-K combine, called by all eight epilogue warps once the tile's (partial) accumulator is complete.  Every warp
// parks its 32 rows x [j0, j1) columns in the workspace; the CTA that counts in last adds the ksplit partials of the
// tile in split order, writes the sum back to the same TMEM columns and returns true: the normal epilogue then runs
// on it unchanged.  The others return false and hand the accumulator straight back.
__device__ __forceinline__ bool ksplit_combine(const ConvKParams& p, const Smem& sm, uint32_t taddr, int tile, int ks,
                                               int warp, int lane, int row_base, int j0, int j1) {
  // Layout of a partial tile: [column / 4][row][4 columns] — a warp's 32 rows of one 4-column group are 512 contiguous
  // bytes, so the float4 stores and loads below are fully coalesced (row-major would touch 32 lines per instruction).
  const size_t tile_elems = (size_t)128 * (size_t)p.BN;
  float* base = p.ks_partial + (size_t)tile * (size_t)p.ksplit * tile_elems + (size_t)(row_base + lane) * 4;
  float* mine = base + (size_t)ks * tile_elems;
  const bool tr = tile == 0 && warp == 2 && lane == 0;   // (tuning builds: time stamps of tile 0's combine)
  if (tr) CHB_TRACE_KS(ks, 0);
  for (int j = j0; j + 8 <= j1; j += 8) {
    float v[8];
    tmem_ld<8>(taddr + (uint32_t)j, v);
    tmem_ld_fence(v);
    float4* dst = reinterpret_cast<float4*>(mine) + (j >> 2) * 128;
    dst[0] = make_float4(v[0], v[1], v[2], v[3]);
    dst[128] = make_float4(v[4], v[5], v[6], v[7]);
  }
  if (tr) CHB_TRACE_KS(ks, 1);
  fence_acq_rel_gpu();   // this thread's partial is visible device-wide before the CTA counts in
  if (tr) CHB_TRACE_KS(ks, 2);
  named_bar_sync(1, 32 * kEpilogueWarps);
  if (warp == 2 && lane == 0) {
    const unsigned int old = atomicAdd(&p.ks_counter[tile], 1u);
    const bool last = old == (unsigned int)p.ksplit - 1u;
    if (last) p.ks_counter[tile] = 0u;   // ready for the next launch that uses the workspace
    *sm.ks_flag = last ? 1u : 0u;
    if (tr) CHB_TRACE_KS(ks, 3);
  }
  named_bar_sync(1, 32 * kEpilogueWarps);
  if (tr) CHB_TRACE_KS(ks, 4);
  if (*reinterpret_cast<volatile uint32_t*>(sm.ks_flag) == 0u) return false;
  fence_acq_rel_gpu();
  if (tr) CHB_TRACE_KS(ks, 5);
  // The partials sit in L2 (~700 cycles away) and every load below is independent: issue the loads of a whole group
  // of splits before the first add, and take as many columns per step as 16 float4 registers allow.
  const int width = j1 - j0;   // 0, 16, 32, 64 or 128 columns
  if (p.ksplit <= 2 && width % 32 == 0) ksplit_sum<32, 2>(p, base, tile_elems, taddr, j0, j1);
  else if (p.ksplit <= 4 && width % 16 == 0) ksplit_sum<16, 4>(p, base, tile_elems, taddr, j0, j1);
  else ksplit_sum<8, 8>(p, base, tile_elems, taddr, j0, j1);
  if (tr) CHB_TRACE_KS(ks, 6);
  tmem_st_wait();
  if (tr) CHB_TRACE_KS(ks, 7);
  return true;
}

template <int EPI, int ACT, bool WSTAT, bool FAST>
__device__ __forceinline__ void epilogue_role(const ConvKParams& p, const Smem& sm, uint32_t tmem_base, int warp,
                                              int lane) {
    // 8 warps: TMEM lane quadrant = warp % 4 (hardware rule), column half = (warp - 2) / 4.
    const int q = warp & 3;
    const int chalf = (warp - 2) >> 2;
    const int row_base = q * 32;
    const EpiK& e = p.e;
    const uint32_t stg = smem_u32(sm.stg_base + (size_t)(warp - 2) * 4096);
    uint64_t* tfull = sm.tfull;
    uint64_t* tempty = sm.tempty;
    TileGeo tg;
    tg.init(p);

    if (EPI == CHB_EPI_PLAIN) {
      const long long osb = e.o_sb, osy = e.o_sy, osx = e.o_sx;
      const long long rsb = e.r_sb, rsy = e.r_sy, rsx = e.r_sx;
      const int rsh = e.r_shift;
      const int obytes = e.out_dtype == CHB_F16 ? 2 : 4;
      RowOff<8, FAST> ro, oo8;
      RowOff<4, FAST> oo4;
      if constexpr (FAST) {
        oo4.setup(lane, q, 0, (uint32_t)(osy >> 3), (uint32_t)(osx >> 3));
        oo8.setup(lane, q, 0, (uint32_t)(osy >> 2), (uint32_t)(osx >> 2));
        ro.setup(lane, q, rsh, (uint32_t)(rsy >> 2), (uint32_t)(rsx >> 2));
      }
      uint32_t it = 0;
      for (int sched = sched_tile<WSTAT>(p, it); sched >= 0; sched = sched_tile<WSTAT>(p, ++it)) {
        const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
        int tile = sched;
        const int ks = split_of_tile(p, tile);
        const int n_tile = FAST ? tg.set_tile_fast(p, tile) : tg.set_tile(p, tile);
        int b, y, x;
        bool valid = true;
        if (FAST) tg.pixel_fast(row_base + lane, b, y, x);
        else valid = tg.pixel(row_base + lane, b, y, x);
        // columns per warp-half (multiple of 8).  Fast geometry with a 32-column N tile: the first warp-half takes all
        // 32 columns (one staged block), the second has nothing to do but hand the accumulator back.
        const bool narrow = FAST && p.BN == 32;
        const int ch = narrow ? (chalf == 0 ? 32 : 0) : (p.BN >> 1);
        int j = narrow ? 0 : chalf * ch;
        const int jend = j + ch;
        const int n0 = n_tile * p.BN;
        const bool staged = FAST || (e.o_sn == 1 && (e.o_ngroup <= 0 || (e.o_ngroup % 32) == 0));
        if constexpr (FAST) {
          const uint32_t ob = tile_base16(tg, osb, osy, osx, 0, obytes);
          oo4.retile(ob);
          oo8.retile(ob);
          if (e.res) ro.retile(tile_base16(tg, rsb, rsy, rsx, rsh, 4));
        } else if (staged) {
          auto o_elem = [=](int bb_, int yy, int xx) { return (long long)bb_ * osb + (long long)yy * osy + (long long)xx * osx; };
          if (e.out_dtype == CHB_F16) row_offsets<4>(lane, row_base, tg, 2, o_elem, oo4.o);
          else row_offsets<8>(lane, row_base, tg, 4, o_elem, oo8.o);
          if (e.res) {
            row_offsets<8>(lane, row_base, tg, 4, [=](int bb_, int yy, int xx) {
              return (long long)bb_ * rsb + (long long)(yy >> rsh) * rsy + (long long)(xx >> rsh) * rsx;
            }, ro.o);
          }
        }
        const int bi = b < p.B ? b : p.B - 1;
        float breg[4] = {0.f, 0.f, 0.f, 0.f};
        if (FAST && e.bias) {
          const float* bp = e.bias + (e.bias_per_image ? (long long)bi * e.nrows : 0) + n0 + j + lane;
#pragma unroll
          for (int blk = 0; blk < 4; ++blk)
            if (32 * blk < ch) breg[blk] = __ldg(bp + 32 * blk);
        }
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        if (it == 0 && warp == 2 && lane == 0) { CHB_TRACE_AT(5); CHB_TRACE_CTA(1); }
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256u;
        bool run = true;
        if (!WSTAT && p.ksplit > 1) run = ksplit_combine(p, sm, taddr, tile, ks, warp, lane, row_base, j, jend);
        if (!run) {
          // another CTA finishes this tile
        } else if (FAST) {
#pragma unroll
          for (int blk = 0; blk < 4; ++blk)
            if (32 * blk < ch)
              plain_block32<ACT, FAST>(e, taddr + (uint32_t)(j + 32 * blk), stg, lane, n0 + j + 32 * blk, bi, ro, oo4, oo8,
                                       breg[blk]);
        } else {
          for (; j + 32 <= jend; j += 32) {
            if (staged && n0 + j + 32 <= p.N) {
              plain_block32<ACT, FAST>(e, taddr + (uint32_t)j, stg, lane, n0 + j, bi, ro, oo4, oo8);
            } else {
              plain_chunk<32, ACT>(p, e, taddr + (uint32_t)j, n0 + j, valid, b, y, x);
            }
          }
          for (; j + 16 <= jend; j += 16) plain_chunk<16, ACT>(p, e, taddr + (uint32_t)j, n0 + j, valid, b, y, x);
          for (; j + 8 <= jend; j += 8) plain_chunk<8, ACT>(p, e, taddr + (uint32_t)j, n0 + j, valid, b, y, x);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
        if (it == 0 && warp == 2 && lane == 0) CHB_TRACE_AT(6);
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
      RowOff<8, FAST> xo, xon;
      RowOff<4, FAST> oo;
      if constexpr (FAST) {
        xon.setup(lane, q, xsh, (uint32_t)(xsy >> 2), (uint32_t)(xsx >> 2));
        oo.setup(lane, q, 0, (uint32_t)(osy >> 3), (uint32_t)(osx >> 3));
      }
      auto x_offsets = [&](const TileGeo& t, RowOff<8, FAST>& o) {
        if constexpr (FAST) o.retile(tile_base16(t, xsb, xsy, xsx, xsh, 4));
        else row_offsets<8>(lane, row_base, t, 4, x_elem, o.o);
      };
      uint4 pf[8];
      if (sched_tile<WSTAT>(p, 0) >= 0) {
        const int nt0 = FAST ? tg.set_tile_fast(p, sched_tile<WSTAT>(p, 0)) : tg.set_tile(p, sched_tile<WSTAT>(p, 0));
        x_offsets(tg, xon);
        gather_issue_o<8, FAST>(reinterpret_cast<const char*>(e.x + nt0 * half_n + chalf * cw) + cl8 * 16, xon, pf);
      }
      uint32_t it = 0;
      for (int tile = sched_tile<WSTAT>(p, it); tile >= 0; tile = sched_tile<WSTAT>(p, ++it)) {
        const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
        const int n_tile = FAST ? tg.set_tile_fast(p, tile) : tg.set_tile(p, tile);
        const int c0 = n_tile * half_n;
        const int nrow0 = n_tile * p.BN;
        int b, y, x;
        bool valid = true;
        if (FAST) tg.pixel_fast(row_base + lane, b, y, x);
        else valid = tg.pixel(row_base + lane, b, y, x);
        float nz = 0.f;
        if (valid && e.noise) nz = __ldg(e.noise + ((long long)b * p.W + x) * p.H + y);
        const float* ca = e.chan + c0;
        xo = xon;
        if constexpr (FAST) oo.retile(tile_base16(tg, osb, osy, osx, 0, 2));
        else row_offsets<4>(lane, row_base, tg, 2, o_elem, oo.o);
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
            gather_issue_o<8, FAST>(reinterpret_cast<const char*>(e.x + c0 + j + 32) + cl8 * 16, xo, pf);
          } else if (sched_tile<WSTAT>(p, it + 1) >= 0) {
            TileGeo tn = tg;
            const int ntn = FAST ? tn.set_tile_fast(p, sched_tile<WSTAT>(p, it + 1))
                                 : tn.set_tile(p, sched_tile<WSTAT>(p, it + 1));
            x_offsets(tn, xon);
            gather_issue_o<8, FAST>(reinterpret_cast<const char*>(e.x + ntn * half_n + chalf * cw) + cl8 * 16, xon, pf);
          }
          constexpr bool split = EPI == kEpiModulateSplit;  // compile time: the lo half costs 16 more live registers
          uint4 hk[4], lk[4];
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            const int jj = j + 16 * sub;
            float g[16], be[16];
            tmem_ld<16>(taddr + (uint32_t)jj, g);
            tmem_ld<16>(taddr + (uint32_t)(half_n + jj), be);
            float4 bg[4], bb[4], av[4], cv[4], nv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              bg[i] = ldg_keep(reinterpret_cast<const float4*>(e.bias + nrow0 + jj) + i);
              bb[i] = ldg_keep(reinterpret_cast<const float4*>(e.bias + nrow0 + half_n + jj) + i);
              av[i] = ldg_keep(reinterpret_cast<const float4*>(ca + jj) + i);
              cv[i] = ldg_keep(reinterpret_cast<const float4*>(ca + cs + jj) + i);
              nv[i] = ldg_keep(reinterpret_cast<const float4*>(ca + 2 * cs + jj) + i);
            }
            tmem_ld_fence(g);
            tmem_ld_fence(be);
            modulate16<ACT, split>(g, be, &xr[4 * sub], nz, bg, bb, av, cv, nv, hk[2 * sub], hk[2 * sub + 1],
                                   lk[2 * sub], lk[2 * sub + 1]);
          }
          if (u + 1 == units) {
            // the accumulator has been fully read: hand the TMEM buffer back before the stores
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
          }
          char* obase = reinterpret_cast<char*>(reinterpret_cast<__half*>(e.out) + c0 + j) + cl4 * 16;
          scatter_o<4, FAST>(stg, lane, obase, oo, hk);
          if constexpr (split) scatter_o<4, FAST>(stg, lane, obase + e.o_lo * 2, oo, lk);
        }
      }
    }
}

}  // namespace chb
