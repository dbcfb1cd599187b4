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
#include "conv_mainloop.cuh"
#include "ptx_sm100.cuh"

#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace chb {

// Planner switches for in-box A/B experiments (tools/ab_env.sh).  They exist only in builds made with
// -DCHB_TUNING_ENV (tools/build_variant.py); the product library never reads the environment.
#ifdef CHB_TUNING_ENV
static inline int tuning_env(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}
#else
static inline int tuning_env(const char*, int dflt) { return dflt; }
#endif

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
// tcgen05 kernel.  Shared memory (1 KB aligned):
//   [nstages x stage_bytes pipeline][1 KB mbarriers + TMEM slot][8 x 4 KB epilogue staging]
//   [nhalo x halo_buf_bytes halo tiles][wstat_bytes resident weights]
// ------------------------------------------------------------------------------------------------
template <int EPI, int ACT, bool WSTAT, bool FAST>
__global__ void __launch_bounds__(kConvThreads, 1) conv_igemm_kernel(const __grid_constant__ ConvKParams p) {
  extern __shared__ uint8_t smem_raw[];
  if (threadIdx.x == 0) { CHB_TRACE_AT(0); CHB_TRACE_CTA(0); }
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Smem sm;
  sm.stage_base = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.nstages * p.stage_bytes);
  sm.full = bars;
  sm.empty = bars + kMaxStages;
  sm.tfull = bars + 2 * kMaxStages;
  sm.tempty = sm.tfull + 2;
  sm.hfull = sm.tempty + 2;
  sm.hempty = sm.hfull + kMaxHalo;
  sm.wbar = sm.hempty + kMaxHalo;
  sm.tmem_slot = reinterpret_cast<uint32_t*>(sm.wbar + 1);
  sm.ks_flag = sm.tmem_slot + 1;
  sm.stg_base = smem + (size_t)p.nstages * p.stage_bytes + 1024;
  sm.halo_base = sm.stg_base + kEpilogueWarps * 4096;
  sm.wstat_base = sm.halo_base + (size_t)p.nhalo * p.halo_buf_bytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nseg; ++s) {
      tma_prefetch_desc(&p.tmA[s]);
      tma_prefetch_desc(&p.tmW[s]);
      if (p.seg[s].halo) tma_prefetch_desc(&p.tmH[s]);
    }
    for (int i = 0; i < kMaxHalo; ++i) {
      mbar_init(&sm.hfull[i], 1);
      mbar_init(&sm.hempty[i], 1);
    }
    for (int i = 0; i < kMaxStages; ++i) {
      mbar_init(&sm.full[i], 1);
      mbar_init(&sm.empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sm.tfull[i], 1);
      mbar_init(&sm.tempty[i], kEpilogueWarps);
    }
    mbar_init(sm.wbar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(sm.tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *sm.tmem_slot;
  if (threadIdx.x == 0) CHB_TRACE_AT(1);

  if (warp == 0) {
    producer_role<WSTAT>(p, sm);
  } else if (warp == 1) {
    mma_role<WSTAT>(p, sm, tmem_base);
  } else {
    epilogue_role<EPI, ACT, WSTAT, FAST>(p, sm, tmem_base, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) CHB_TRACE_AT(7);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  if (threadIdx.x == 32) { CHB_TRACE_AT(8); CHB_TRACE_CTA(2); }
}

// ------------------------------------------------------------------------------------------------
// SIMT checker kernel: same contract, one thread per output element, no tensor cores / TMA.
// Used only by tests (impl = CHB_IMPL_SIMT_DEBUG) to bisect the tcgen05 main loop from the epilogue.
// ------------------------------------------------------------------------------------------------
struct SimtSeg {
  const __half* a;
  long long a_sb, a_sy, a_sx;
  int ch_off, C, Cw, taps, per_image, a_pad;
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
    const int K = sg.taps * sg.Cw;
    const __half* wrow = sg.w + (sg.per_image ? (long long)b * (sg.w_sb > 0 ? sg.w_sb : (long long)p.e.nrows * K) : 0) + (long long)nrow * K;
    for (int tap = 0; tap < sg.taps; ++tap) {
      const int yy = y + sg.a_pad + (sg.taps == 9 ? tap / 3 - 1 : 0);
      const int xx = x + sg.a_pad + (sg.taps == 9 ? tap % 3 - 1 : 0);
      if (yy < 0 || yy >= p.H + 2 * sg.a_pad || xx < 0 || xx >= p.W + 2 * sg.a_pad) continue;
      const __half* ap = sg.a + (long long)b * sg.a_sb + (long long)yy * sg.a_sy + (long long)xx * sg.a_sx + sg.ch_off;
      const __half* wp = wrow + tap * sg.Cw;
      for (int c = 0; c < sg.C; ++c) acc = fmaf(__half2float(ap[c]), __half2float(wp[c >= sg.Cw ? c - sg.Cw : c]), acc);
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
      __half* op = reinterpret_cast<__half*>(p.e.out) + (long long)b * p.e.o_sb + (long long)y * p.e.o_sy +
                   (long long)x * p.e.o_sx + n;
      const __half hi = __float2half_rn(o);
      *op = hi;
      if (p.e.split) op[p.e.o_lo] = __float2half_rn(o - __half2float(hi));
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
  CHB_REQUIRE(d.nseg >= 1 && d.nseg <= kMaxSeg, "1..4 segments");
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
    CHB_REQUIRE(g.a_pad >= 0 && g.a_pad <= 8, "a_pad in 0..8");
    CHB_REQUIRE(g.w_dup == 0 || g.w_dup == 1 || (g.w_dup == 2 && g.C % 128 == 0),
                "w_dup is 0, 1 or 2; a hi+lo split segment needs C % 128 == 0");
    CHB_REQUIRE((reinterpret_cast<uintptr_t>(g.a) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.w) & 15) == 0,
                "operands must be 16-byte aligned");
  }
  if (d.epi == CHB_EPI_MODULATE) {
    CHB_REQUIRE(d.x && d.chan && d.bias, "modulate epilogue needs x, chan and bias");
    CHB_REQUIRE(d.BN % 128 == 0 && d.N == d.Nrows, "modulate epilogue needs BN % 128 == 0 and N == Nrows");
    CHB_REQUIRE(d.act != CHB_ACT_TANH, "modulate epilogue supports none / relu / lrelu");
    CHB_REQUIRE(d.o_sn == 1 || d.o_sn == 0, "modulate output is channels-last");
  }
  if (d.o_split) {
    CHB_REQUIRE(d.out_dtype == CHB_F16 && d.o_sn <= 1 && d.o_ngroup <= 0 && d.o_lo_off > 0 && d.o_lo_off % 8 == 0 &&
                    d.N == d.Nrows && d.BN % 64 == 0,
                "o_split needs fp16 channels-last output, N == Nrows, BN % 64 == 0 and o_lo_off % 8 == 0");
  }
  if (d.ksplit > 1) {
    int min_chunks = 1 << 30;
    for (int s = 0; s < d.nseg; ++s) {
      const int nch = d.seg[s].C / (d.seg[s].C == 32 ? 32 : 64);
      if (nch < min_chunks) min_chunks = nch;
    }
    const long long tiles = (long long)((d.W + d.TW - 1) / d.TW) * ((d.H + d.TH - 1) / d.TH) *
                            ((d.B + d.TB - 1) / d.TB) * (d.Nrows / d.BN);
    CHB_REQUIRE(d.epi == CHB_EPI_PLAIN && d.ks_ws && d.ksplit <= 16 && d.ksplit <= min_chunks && d.BN % 16 == 0 &&
                    tiles * (long long)sizeof(unsigned) <= (long long)kKsCounterBytes &&
                    (reinterpret_cast<uintptr_t>(d.ks_ws) & 15) == 0,
                "ksplit needs the PLAIN epilogue, a 16-byte aligned workspace, ksplit <= 16 and <= the channel chunks "
                "of every segment, and at most 1024 output tiles");
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
  // Halo path (on by default; CHB_HALO=0 turns it off for A/B comparisons): 3x3 segments with 64-channel chunks on
  // 8-wide single-image tiles load a (TH+2)x(TW+2) halo tile once per channel chunk and feed all nine taps from it.
  const int halo_mode = tuning_env("CHB_HALO", 1), wstat_mode = tuning_env("CHB_WSTAT", 2);
  const bool halo_geo = d.TW == 8 && d.TB == 1 && d.TH <= 16;
  // Weight-stationary mode: every segment can take its A operand from halo tiles (3x3 or 1x1-as-centre-tap, 32- or
  // 64-channel chunks), weights are shared by all images and the [BN x K] slab fits next to the halo ring.
  long long wbytes_total = 0;
  bool wstat_ok = halo_mode > 0 && wstat_mode > 0 && halo_geo;
  int kc_max = 32;
  for (int s = 0; s < d.nseg; ++s) {
    const chb_conv_seg& g = d.seg[s];
    const int kc = g.C == 32 ? 32 : 64;
    if (kc > kc_max) kc_max = kc;
    if (g.per_image) wstat_ok = false;
    if (kc == 32 && wstat_mode < 2) wstat_ok = false;  // CHB_WSTAT=1: one-hot (64-byte row) layers streamed (slower since the nine taps of a resident chunk go out in one asm block)
    wbytes_total += (long long)g.taps * (g.w_dup == 2 ? g.C / 2 : g.C) * d.BN * 2;
  }
  // ring buffer of the A operand: a (TH+2)x(TW+2) halo tile, or — when every segment is 1x1 — the exact 128-row tile,
  // which lets a thin 1x1 GEMM (conv_img.taps: K = 192, N = 32) keep 3-4 pixel tiles in flight instead of 2
  bool all_1x1 = true;
  for (int s = 0; s < d.nseg; ++s) all_1x1 = all_1x1 && d.seg[s].taps == 1;
  const int halo_buf = all_1x1 ? (kc_max == 64 ? kATileBytes : kATileBytes / 2)
                               : (kc_max == 64 ? kHaloBufBytes : kHaloBufBytes / 2);
  const int kRegion = 193 * 1024;  // 227 KB - align slack - barriers - epilogue staging
  const int wstat_min_halo = tuning_env("CHB_WSTAT_MINHALO", 2);
  if (wbytes_total + (long long)wstat_min_halo * halo_buf > kRegion || d.Nrows / d.BN > device_sm_count()) wstat_ok = false;
  if (d.ksplit > 1) wstat_ok = false;  // resident weights and split-K answer opposite problems (many / few pixel tiles)
  k.wstat = wstat_ok ? 1 : 0;
  k.halo_any = 0;
  k.halo_bo = 0;
  bool all_halo = true;
  for (int s = 0; s < d.nseg; ++s) {
    const chb_conv_seg& g = d.seg[s];
    k.seg[s].halo = (k.wstat || (halo_mode > 0 && halo_geo && g.taps == 9 && g.C % 64 == 0)) ? 1 : 0;
    k.halo_any |= k.seg[s].halo;
    all_halo = all_halo && k.seg[s].halo;
  }
  k.halo_buf_bytes = halo_buf;
  k.a_region = all_halo ? 0 : kATileBytes;
  k.hg = (k.halo_any && d.BN <= 128) ? 3 : 1;
  if (k.wstat) {
    k.wstat_bytes = (int)wbytes_total;
    k.stage_bytes = 0;
    k.nstages = 0;
    k.nhalo = (int)((kRegion - wbytes_total) / halo_buf);
    if (k.nhalo > kMaxHalo) k.nhalo = kMaxHalo;
  } else {
    k.wstat_bytes = 0;
    k.nhalo = k.halo_any ? 2 : 0;
    if (k.halo_any) {
      const int n = tuning_env(d.BN <= 64 ? "CHB_NHALO64" : (d.BN <= 128 ? "CHB_NHALO128" : "CHB_NHALO256"), 0);
      if (n >= 2 && n <= kMaxHalo) k.nhalo = n;
    }
    k.stage_bytes = k.a_region + ((d.BN * 128 * k.hg + 1023) / 1024) * 1024;
    const int budget = kSmemBudget - k.nhalo * halo_buf;
    k.nstages = budget / k.stage_bytes;
    if (k.nstages > kMaxStages) k.nstages = kMaxStages;
  }
  plan->smem_bytes = k.nstages * k.stage_bytes + 1024 /*align slack*/ + 1024 /*barriers*/ + kEpilogueWarps * 4096 +
                     k.nhalo * halo_buf + k.wstat_bytes;
  long long ktotal = 0, wk_total = 0;
  for (int s = 0; s < d.nseg; ++s) {
    const chb_conv_seg& g = d.seg[s];
    const int kc = g.C == 32 ? 32 : 64;
    k.seg[s].taps = g.taps;
    k.seg[s].kc = kc;
    k.seg[s].nchunk = g.C / kc;
    const int Cw = g.w_dup == 2 ? g.C / 2 : g.C;  // channels the weights span
    k.seg[s].nchunk_w = Cw / kc;
    k.seg[s].ch_off = g.ch_off;
    k.seg[s].per_image = g.per_image;
    k.seg[s].xy_off = g.a_pad;
    k.seg[s].wofs = (int)(wk_total * d.BN * 2);  // offset of this segment inside the resident weight slab
    ktotal += (long long)g.taps * g.C;
    wk_total += (long long)g.taps * Cw;
    {
      cuuint64_t dims[4] = {(cuuint64_t)g.Ca, (cuuint64_t)(d.W + 2 * g.a_pad), (cuuint64_t)(d.H + 2 * g.a_pad),
                            (cuuint64_t)d.B};
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
      const cuuint64_t K = (cuuint64_t)g.taps * Cw;
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
  e.split = d.o_split ? 1 : 0; e.o_lo = d.o_lo_off;
  e.res = d.res; e.r_sb = d.r_sb; e.r_sy = d.r_sy; e.r_sx = d.r_sx; e.r_shift = d.r_shift;
  e.x = d.x; e.x_sb = d.x_sb; e.x_sy = d.x_sy; e.x_sx = d.x_sx; e.x_shift = d.x_shift;
  e.noise = d.noise;
  e.chan = d.chan;
  e.chan_stride = d.N / 2;
  // Fast tile geometry (see ConvKParams::fast).  CHB_FAST=0 turns it off for A/B comparisons.
  {
    auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
    auto ilog2 = [](int v) { int s = 0; while ((1 << s) < v) ++s; return s; };
    const long long ob = d.out_dtype == CHB_F16 ? 2 : 4;
    bool fast = d.TB == 1 && d.TW == 8 && d.TH == 16 && d.H % 16 == 0 && d.W % 8 == 0 && pow2(k.tiles_x) &&
                pow2(k.tiles_y) && d.N == d.Nrows && (d.BN % 64 == 0 || (d.BN == 32 && d.epi == CHB_EPI_PLAIN)) &&
                (d.o_sn == 1 || d.epi == CHB_EPI_MODULATE) &&
                (d.o_ngroup <= 0 || d.o_ngroup % 32 == 0) && (d.o_sb * ob) % 16 == 0 && (d.o_sy * ob) % 16 == 0 &&
                (d.o_sx * ob) % 16 == 0 && (reinterpret_cast<uintptr_t>(d.out) & 15) == 0;
    if (d.epi == CHB_EPI_PLAIN && d.res)
      fast = fast && d.r_sb % 4 == 0 && d.r_sy % 4 == 0 && d.r_sx % 4 == 0 && (d.r_shift == 0 || d.r_shift == 1) &&
             (reinterpret_cast<uintptr_t>(d.res) & 15) == 0;
    if (d.epi == CHB_EPI_MODULATE)
      fast = fast && d.out_dtype == CHB_F16 && d.x_sb % 4 == 0 && d.x_sy % 4 == 0 && d.x_sx % 4 == 0 &&
             (d.x_shift == 0 || d.x_shift == 1) && (reinterpret_cast<uintptr_t>(d.x) & 15) == 0;
    fast = fast && tuning_env("CHB_FAST", 1) != 0;
    k.fast = fast ? 1 : 0;
    k.tx_sh = ilog2(k.tiles_x);
    k.ty_sh = ilog2(k.tiles_y);
  }
  const int total = k.m_tiles * k.n_tiles;
  const int sms = device_sm_count();
  k.ksplit = d.ksplit > 1 ? d.ksplit : 1;
  k.tiles_mn = total;
  k.ks_counter = reinterpret_cast<unsigned int*>(d.ks_ws);
  k.ks_partial = d.ks_ws ? reinterpret_cast<float*>(reinterpret_cast<char*>(d.ks_ws) + kKsCounterBytes) : nullptr;
  plan->grid = total * k.ksplit < sms ? total * k.ksplit : sms;
  if (k.wstat) {
    int per = sms / k.n_tiles;  // CTAs per n_tile
    if (per > k.m_tiles) per = k.m_tiles;
    plan->grid = per * k.n_tiles;
  }
  plan->flops = 2.0 * 128.0 * d.BN * (double)ktotal * (double)total;
  return CHB_OK;
}

typedef void (*ConvKernelFn)(const ConvKParams);

template <bool WSTAT, bool FAST>
static ConvKernelFn pick_kernel_wf(int epi, int act) {
  if (epi == CHB_EPI_PLAIN) {
    switch (act) {
      case CHB_ACT_RELU: return conv_igemm_kernel<CHB_EPI_PLAIN, CHB_ACT_RELU, WSTAT, FAST>;
      case CHB_ACT_LRELU: return conv_igemm_kernel<CHB_EPI_PLAIN, CHB_ACT_LRELU, WSTAT, FAST>;
      case CHB_ACT_TANH: return conv_igemm_kernel<CHB_EPI_PLAIN, CHB_ACT_TANH, WSTAT, FAST>;
      default: return conv_igemm_kernel<CHB_EPI_PLAIN, CHB_ACT_NONE, WSTAT, FAST>;
    }
  }
  if (epi == kEpiModulateSplit) {
    switch (act) {
      case CHB_ACT_LRELU: return conv_igemm_kernel<kEpiModulateSplit, CHB_ACT_LRELU, WSTAT, FAST>;
      case CHB_ACT_RELU: return conv_igemm_kernel<kEpiModulateSplit, CHB_ACT_RELU, WSTAT, FAST>;
      default: return conv_igemm_kernel<kEpiModulateSplit, CHB_ACT_NONE, WSTAT, FAST>;
    }
  }
  switch (act) {
    case CHB_ACT_LRELU: return conv_igemm_kernel<CHB_EPI_MODULATE, CHB_ACT_LRELU, WSTAT, FAST>;
    case CHB_ACT_RELU: return conv_igemm_kernel<CHB_EPI_MODULATE, CHB_ACT_RELU, WSTAT, FAST>;
    default: return conv_igemm_kernel<CHB_EPI_MODULATE, CHB_ACT_NONE, WSTAT, FAST>;
  }
}
static ConvKernelFn pick_kernel(int epi, int act, int wstat, int fast) {
  if (fast) return wstat ? pick_kernel_wf<true, true>(epi, act) : pick_kernel_wf<false, true>(epi, act);
  return wstat ? pick_kernel_wf<true, false>(epi, act) : pick_kernel_wf<false, false>(epi, act);
}

static int ensure_smem_attr() {
  static std::once_flag once;
  static cudaError_t err = cudaSuccess;
  std::call_once(once, [] {
    for (int epi = 0; epi < 3 && err == cudaSuccess; ++epi)
      for (int act = 0; act < 4 && err == cudaSuccess; ++act)
        for (int w = 0; w < 4 && err == cudaSuccess; ++w)
          err = cudaFuncSetAttribute(pick_kernel(epi, act, w & 1, w >> 1), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024);
  });
  if (err != cudaSuccess) {
    set_error(std::string("cudaFuncSetAttribute(max dynamic smem) failed: ") + cudaGetErrorString(err));
    return CHB_ERR_CUDA;
  }
  return CHB_OK;
}

int ensure_conv_kernels_ready() { return ensure_smem_attr(); }

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
      sp.seg[s].Cw = d.seg[s].w_dup == 2 ? d.seg[s].C / 2 : d.seg[s].C;
      sp.seg[s].per_image = d.seg[s].per_image;
      sp.seg[s].a_pad = d.seg[s].a_pad;
      sp.seg[s].w = reinterpret_cast<const __half*>(d.seg[s].w);
      sp.seg[s].w_sb = d.seg[s].w_sb;
    }
    conv_simt_kernel<<<device_sm_count() * 8, 256, 0, stream>>>(sp);
  } else {
    int rc = ensure_smem_attr();
    if (rc != CHB_OK) return rc;
    const int epi = (plan.desc.epi == CHB_EPI_MODULATE && plan.desc.o_split) ? kEpiModulateSplit : plan.desc.epi;
    pick_kernel(epi, plan.desc.act, plan.kp.wstat, plan.kp.fast)<<<plan.grid, kConvThreads, plan.smem_bytes, stream>>>(plan.kp);
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

int chb_struct_size(int which) {
  switch (which) {
    case 0: return (int)sizeof(chb_conv_seg);
    case 1: return (int)sizeof(chb_conv_desc);
    case 2: return (int)sizeof(chb_gen_config);
    case 3: return (int)sizeof(chb_mlp_layer);
    default: return -1;
  }
}
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

#ifdef CHB_TRACE
extern "C" int chb_debug_trace_read(unsigned long long* host16) {
  return cudaMemcpyFromSymbol(host16, chb::chb_trace_buf, sizeof(unsigned long long) * 16) == cudaSuccess ? 0 : -1;
}
extern "C" int chb_debug_trace_ks_read(unsigned long long* host128) {
  return cudaMemcpyFromSymbol(host128, chb::chb_trace_ks, sizeof(unsigned long long) * 128) == cudaSuccess ? 0 : -1;
}
extern "C" int chb_debug_trace_cta_read(unsigned long long* host480) {
  return cudaMemcpyFromSymbol(host480, chb::chb_trace_cta, sizeof(unsigned long long) * 480) == cudaSuccess ? 0 : -1;
}
#endif

int64_t chb_conv_ksplit_workspace_bytes(int max_ctas) {
  return max_ctas > 0 ? (int64_t)chb::ksplit_workspace_bytes(max_ctas) : 0;
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
