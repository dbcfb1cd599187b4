// Internal (C++) interface of the implicit-GEMM convolution operator.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <string>

#include "../../include/ctrlhair_b200.h"

namespace chb {

constexpr int kMaxSeg = 4;
constexpr int kEpiModulateSplit = 2;  // kernel-internal: CHB_EPI_MODULATE with chb_conv_desc.o_split (own instantiation)
constexpr int kATileBytes = 16384;  // 128 rows x 128 B
constexpr int kMaxStages = 8;
constexpr int kMaxHalo = 12;
constexpr int kHaloBufBytes = 24576;  // (16+2) x (8+2) rows x 128 B, rounded up to 1 KB
constexpr int kSmemBudget = 192 * 1024;  // pipeline stages; + 32 KB epilogue staging + barriers <= 227 KB
constexpr int kEpilogueWarps = 8;
constexpr int kConvThreads = 64 + 32 * kEpilogueWarps;  // warp0 TMA, warp1 MMA, warps 2-9 epilogue

struct SegK {
  int taps, nchunk, kc, ch_off, per_image;
  int xy_off;  // explicit input border (chb_conv_seg.a_pad): added to the TMA x / y coordinates
  int wofs;  // weight-stationary mode: byte offset of this segment inside the resident weight slab
  int halo;  // 1: the A operand of this 3x3 segment is loaded once per channel chunk as a (TH+2)x(TW+2) halo tile
  int nchunk_w;  // channel chunks the WEIGHTS have per tap: nchunk, or nchunk / 2 for a hi+lo split activation
                 // (chb_conv_seg.w_dup == 2: chunk c of the activation uses weight chunk c % nchunk_w)
};

struct EpiK {
  int act;
  const float* bias;
  int bias_per_image;
  int nrows;
  // plain
  void* out;
  int out_dtype;
  long long o_sb, o_sy, o_sx, o_sn;
  int o_ngroup;
  long long o_sgroup;
  int split;           // fp16 hi+lo split output: lo = fp16(v - hi) goes o_lo elements after hi
  long long o_lo;
  const float* res;
  long long r_sb, r_sy, r_sx;
  int r_shift;
  // modulate
  const float* x;
  long long x_sb, x_sy, x_sx;
  int x_shift;
  const float* noise;
  const float* chan;   // planar [3][C]: a = rstd | c = -mean*rstd | nv = noise_var*rstd
  int chan_stride;     // C
};

struct ConvKParams {
  CUtensorMap tmA[kMaxSeg];
  CUtensorMap tmW[kMaxSeg];
  CUtensorMap tmH[kMaxSeg];  // halo boxes (segments with halo = 1)
  SegK seg[kMaxSeg];
  int wstat;          // weight-stationary mode
  int wstat_bytes;    // resident weight slab size
  int nhalo;          // halo buffers in the ring
  int halo_buf_bytes;
  int halo_any;   // some segment uses the halo path: two halo buffers follow the pipeline stages
  int halo_bo;    // descriptor base-offset mode for non-1024B-aligned tap views (0: none, 1: (addr >> 7) & 7)
  int nseg;
  int B, H, W;
  int TW, TH, TB, rows;
  int tiles_x, tiles_y, m_tiles, n_tiles;
  int BN, N;
  int nstages, stage_bytes;
  int a_region;  // bytes of the per-stage A tile (0 when every segment takes its A operand from halo tiles)
  int hg;        // taps per pipeline stage on the halo path (1 or 3)
  // Fast tile geometry: every tile is a full 16 x 8 pixel block of one image, the tile grid is a power of two in x
  // and y and all epilogue tensors have 16-byte aligned rows, so per-tile addresses are an affine function of the
  // tile coordinates (no per-row pixel decode, no bounds predicates).
  int fast, tx_sh, ty_sh;
  // Split-K (chb_conv_desc.ksplit > 1): tile ids run over ksplit * tiles_mn; split ks of an output tile accumulates the
  // channel chunks [nchunk * ks / ksplit, nchunk * (ks + 1) / ksplit) of every segment, parks its fp32 accumulator in
  // ks_partial and counts in on ks_counter[tile]; the CTA that arrives last sums the partials IN SPLIT ORDER (a fixed
  // order: the result does not depend on which CTA that is), writes the sum back to TMEM and runs the epilogue.
  int ksplit, tiles_mn;
  float* ks_partial;        // [tiles_mn][ksplit][BN / 4][128 rows][4]
  unsigned int* ks_counter; // [tiles_mn], zero between launches (the last CTA resets its tile's counter)
  EpiK e;
};

constexpr size_t kKsCounterBytes = 4096;   // head of a split-K workspace: one counter per output tile (<= 1024 tiles)
// bytes of the split-K workspace that serves any launch with ksplit * tiles_mn <= max_ctas (BN <= 256)
inline size_t ksplit_workspace_bytes(int max_ctas) { return kKsCounterBytes + (size_t)max_ctas * 128 * 256 * 4; }

struct ConvPlan {
  ConvKParams kp;
  chb_conv_desc desc;  // kept for the SIMT checker kernel
  int grid;
  int smem_bytes;
  double flops;  // tensor-core FLOPs issued (padding included)
};

#ifdef CHB_TRACE
// Tuning builds only (tools/build_variant.py with VARIANT_FLAGS=-DCHB_TRACE): SM-clock timestamps of CTA 0's first tile,
// read back with chb_debug_trace_read (tools/gpu_trace_small.py).
static __device__ unsigned long long chb_trace_buf[16];   // (one copy per translation unit; conv_igemm.cu's is the live one)
#define CHB_TRACE_AT(k) do { if (blockIdx.x == 0) chb_trace_buf[k] = clock64(); } while (0)
static __device__ unsigned long long chb_trace_cta[3 * 160];   // per CTA: globaltimer at entry, epilogue start, exit
__device__ __forceinline__ unsigned long long chb_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define CHB_TRACE_CTA(k) do { if (blockIdx.x < 160) chb_trace_cta[3 * blockIdx.x + (k)] = chb_globaltimer(); } while (0)
static __device__ unsigned long long chb_trace_ks[16 * 8];   // split-K combine of output tile 0: [split][point]
#define CHB_TRACE_KS(ks, k) do { chb_trace_ks[8 * (ks) + (k)] = chb_globaltimer(); } while (0)
#else
#define CHB_TRACE_AT(k) do { } while (0)
#define CHB_TRACE_CTA(k) do { } while (0)
#define CHB_TRACE_KS(ks, k) do { } while (0)
#endif

void set_error(const std::string& msg);
int build_conv_plan(const chb_conv_desc& d, ConvPlan* plan);
int launch_conv_plan(const ConvPlan& plan, int impl, cudaStream_t stream);
int ensure_conv_kernels_ready();  // one-time cudaFuncSetAttribute of every instantiation (not legal inside a capture)
int device_sm_count();
int onehot_pyramid_ones(const uint8_t* labels, int B, int S, int nlevels, const int* shifts, void* const* outs,
                        int nclass, int ones_ch0, int ones_n, void* stream);
extern "C" const void* chb_noise_fill_kernel_address(void);
int img_from_taps(const float* y, const float* bias, float* out, int B, int S, cudaStream_t stream);
int codes_cast_transpose(const float* in, void* out, int B, int NC, int L, cudaStream_t stream);

}  // namespace chb
