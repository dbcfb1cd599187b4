// One colour/texture training sub-step (config 045) on the GPU: forward, losses, backward (including the WGAN-GP
// double backward) and Adam, fp32.  Reference: color_texture_branch/train.py:115-148 (loop body), solver.py:85-117
// (forward), :218-245 + :186-216 (forward_d / forward_general_dis), :119-166 (forward_g), my_torchlib/train_utils.py:
// 54-89 (train), model_eigengan.py:14-89, model.py:86-127, predictor/predictor_model.py:14-41.
//
// The three nets are 256-wide MLPs and the per-rank batch is 16-256 rows, so a sub-step is ~90 GEMMs of a few MFLOP
// each: it is bound by the latency of its dependent chain, not by FLOPs or bytes.  The design answer is (a) no autograd
// tape - the backward of every path is written out, with the leaky-ReLU masks re-derived from the stored
// pre-activations, (b) one strided fp32 GEMM kernel for all three contraction shapes (X W^T, dY W, dY^T X) whose tile
// is built for latency (register double-buffered staging, 4 x 4 register tiles), (c) the sub-step recorded as an
// operation list whose read / write ranges give the REAL dependencies, and executed either as an explicit CUDA graph
// (one node per operation, one edge per dependency: weight-gradient GEMMs, the two generator passes and the three
// discriminator passes run beside the data path) or by one persistent cooperative kernel with grid barriers.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ctrlhair_b200.h"
#include "conv_igemm.cuh"

namespace chb {
namespace ctt {

constexpr int kCode = 512, kHid = 256, kNoise = 8, kSub = 2, kGLayers = 4, kDOut = 11, kGIn = 5, kPOut = 4, kCHid = 32;
constexpr int kXinLd = 8;  // padded row length of the generator's 5-wide input
constexpr float kSlope = 0.2f;
constexpr int kMaxBatch = 1024;

__device__ __forceinline__ float lrelu(float v) { return v > 0.f ? v : kSlope * v; }
__device__ __forceinline__ float dlrelu(float pre) { return pre > 0.f ? 1.f : kSlope; }

// C(m,n) (+)= sum_k A(m,k) B(k,n)  [+ bias(n)]  [* lrelu'(mask(m,n))], arbitrary strides, optional lrelu on either operand.
struct Gemm {
  const float* A; long a_sm, a_sk;
  const float* B; long b_sk, b_sn;  // B == nullptr: all ones
  float* C; long c_sm;
  int M, N, K;
  const float* bias;
  const float* mask; long mask_sm;
  int a_act, b_act, accumulate;
};

// 32 x 32 output tile per CTA, K consumed in chunks of 128.  These GEMMs are a few MFLOP each and sit on a dependent
// chain, so what matters is the latency of one tile:
//   * staging: every thread has 32 independent global loads in flight per chunk, and the loads of chunk c + 1 are
//     issued before chunk c is computed (register double buffer);
//   * compute: the 256 threads split the chunk's K range four ways (64 threads per quarter), each thread owning a
//     4 x 4 register tile fed by two 16-byte shared-memory reads per k (the round-1 loop read five words per four FMAs
//     and was bound by shared-memory issue: ~3 us per chunk); the four partial tiles are summed through shared memory.
constexpr int kGemmKC = 128;
constexpr int kGemmLd = 36;   // row pitch of the staged operands (floats): a multiple of 4, so rows stay 16-byte aligned
__device__ __forceinline__ void gemm_tile(const Gemm& g, int bx, int by, float (*As)[kGemmLd], float (*Bs)[kGemmLd]) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;                       // staging / epilogue layout
  const int slice = threadIdx.x >> 6, ty4 = (threadIdx.x & 63) >> 3, tx4 = threadIdx.x & 7;   // compute layout
  const int m0 = by * 32, n0 = bx * 32;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float av[16], bv[16];
  auto load_chunk = [&](int k0) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      int m, k;
      if (g.a_sk == 1) { m = ty + 8 * (j >> 2); k = tx + 32 * (j & 3); } else { m = tx; k = ty + 8 * j; }
      float v = 0.f;
      if (m0 + m < g.M && k0 + k < g.K) v = g.A[(long)(m0 + m) * g.a_sm + (long)(k0 + k) * g.a_sk];
      av[j] = v;
      int n, kb;
      if (g.b_sn == 1) { kb = ty + 8 * j; n = tx; } else { n = ty + 8 * (j >> 2); kb = tx + 32 * (j & 3); }
      float w = 0.f;
      if (n0 + n < g.N && k0 + kb < g.K) w = g.B ? g.B[(long)(k0 + kb) * g.b_sk + (long)(n0 + n) * g.b_sn] : 1.f;
      bv[j] = w;
    }
  };
  load_chunk(0);
  for (int k0 = 0; k0 < g.K; k0 += kGemmKC) {
    __syncthreads();   // the previous chunk has been consumed
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      int m, k;
      if (g.a_sk == 1) { m = ty + 8 * (j >> 2); k = tx + 32 * (j & 3); } else { m = tx; k = ty + 8 * j; }
      As[k][m] = g.a_act ? lrelu(av[j]) : av[j];
      int n, kb;
      if (g.b_sn == 1) { kb = ty + 8 * j; n = tx; } else { n = ty + 8 * (j >> 2); kb = tx + 32 * (j & 3); }
      Bs[kb][n] = g.b_act ? lrelu(bv[j]) : bv[j];
    }
    __syncthreads();
    if (k0 + kGemmKC < g.K) load_chunk(k0 + kGemmKC);   // in flight while this chunk is computed
    // rows / columns beyond K were staged as zeros: the quarter can always run its 32 steps
#pragma unroll 8
    for (int kk = slice * 32; kk < slice * 32 + 32; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty4 * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx4 * 4]);
      const float ar[4] = {a.x, a.y, a.z, a.w}, br[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
  }
  __syncthreads();
  float* red = &As[0][0];   // 4 partial 32 x 33 tiles (4 224 floats) in the 4 608 floats of As
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) red[(slice * 32 + ty4 * 4 + i) * 33 + tx4 * 4 + j] = acc[i][j];
  __syncthreads();
  const int n = n0 + tx;
  if (n < g.N) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int ml = ty + 8 * r, m = m0 + ml;
      if (m >= g.M) continue;
      float v = (red[ml * 33 + tx] + red[(32 + ml) * 33 + tx]) + (red[(64 + ml) * 33 + tx] + red[(96 + ml) * 33 + tx]);
      if (g.bias) v += g.bias[n];
      if (g.mask) v *= dlrelu(g.mask[(long)m * g.mask_sm + n]);
      float* p = g.C + (long)m * g.c_sm + n;
      *p = g.accumulate ? *p + v : v;
    }
  }
  __syncthreads();   // `red` aliases the staging buffer of the next tile
}

__global__ void __launch_bounds__(256) gemm_kernel(const Gemm g) {
  __shared__ __align__(16) float As[kGemmKC][kGemmLd];
  __shared__ __align__(16) float Bs[kGemmKC][kGemmLd];
  gemm_tile(g, blockIdx.x, blockIdx.y, As, Bs);
}

// ---------------------------------------------------------------------------------------------------------------
struct StageArgs {
  chb_cttrain_batch in;
  float *code, *rgb, *pca, *noise, *nc, *label, *alpha, *ones, *e0;
  int *p1, *p2, *p3, *flag;
  int B;
};

__global__ void stage_kernel(const StageArgs a) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  for (int i = t; i < a.B * kCode; i += nt) a.code[i] = a.in.code[i];
  for (int i = t; i < a.B * 3; i += nt) a.rgb[i] = a.in.rgb_mean[i];
  for (int i = t; i < a.B * kNoise; i += nt) a.noise[i] = a.in.noise[i];
  for (int i = t; i < a.B * kDOut; i += nt) a.e0[i] = (i % kDOut) == 0 ? 1.f : 0.f;
  for (int i = t; i < a.B; i += nt) {
    a.pca[i] = a.in.pca_std[i];
    a.nc[i] = a.in.noise_curliness[i];
    a.label[i] = a.in.curliness_label[i];
    a.alpha[i] = a.in.alpha_gp ? a.in.alpha_gp[i] : 0.f;
    a.ones[i] = 1.f;
    a.p1[i] = a.in.perm_rgb[i];
    a.p2[i] = a.in.perm_curliness[i];
    a.p3[i] = a.in.perm_noise[i];
  }
  if (t == 0) *a.flag = a.in.noise_from_encoder;
}

// Generator inputs of both passes (solver.py:89-111): auto-encoder pass reads the discriminator's real-code outputs,
// GAN pass reads the shuffled batch (or, on the encoder-noise coin, the shuffled and detached encoder noise).
struct PrepArgs {
  const float *r, *rgb, *pca, *noise, *nc, *label;
  const int *p1, *p2, *p3, *flag;
  const float* Lvec[kGLayers];  // SubspaceLayer.L of each layer
  float *xinA, *xinG, *zG, *labG;
  float* LzA[kGLayers];
  float* LzG[kGLayers];
  int B;
};

__device__ __forceinline__ void prep_row(const PrepArgs& a, int b) {
  const float* rb = a.r + (long)b * kDOut;
  float* xa = a.xinA + (long)b * kXinLd;
  xa[0] = rb[1 + kNoise];
  xa[1] = a.rgb[b * 3 + 0]; xa[2] = a.rgb[b * 3 + 1]; xa[3] = a.rgb[b * 3 + 2];
  xa[4] = a.pca[b];
  xa[5] = xa[6] = xa[7] = 0.f;
  const int i1 = a.p1[b], i2 = a.p2[b], i3 = a.p3[b];
  float* xg = a.xinG + (long)b * kXinLd;
  xg[0] = a.nc[i2];
  xg[1] = a.rgb[i1 * 3 + 0]; xg[2] = a.rgb[i1 * 3 + 1]; xg[3] = a.rgb[i1 * 3 + 2];
  xg[4] = a.pca[i1];
  xg[5] = xg[6] = xg[7] = 0.f;
  a.labG[b] = a.label[i2];
  const bool enc = *a.flag != 0;
  for (int j = 0; j < kNoise; ++j) {
    const float za = rb[1 + j];
    const float zg = enc ? a.r[(long)i3 * kDOut + 1 + j] : a.noise[(long)i3 * kNoise + j];
    a.zG[(long)b * kNoise + j] = zg;
    const int l = j / kSub, k = j % kSub;
    a.LzA[l][b * kSub + k] = a.Lvec[l][k] * za;
    a.LzG[l][b * kSub + k] = a.Lvec[l][k] * zg;
  }
}
__global__ void prep_kernel(const PrepArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < a.B) prep_row(a, b);
}

__global__ void xhat_kernel(const float* __restrict__ code, const float* __restrict__ f, const float* __restrict__ alpha,
                            float* __restrict__ xh, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float al = alpha[i / kCode];
  xh[i] = al * code[i] + (1.f - al) * f[i];
}

// ---- block-wide helpers (1024 threads) -----------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ float block_sum(float v, float* red) {  // red: 33 floats of shared memory
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    float t = lane < nw ? red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

struct LossArgs {
  chb_cttrain_config cfg;
  int B;
  const float *code, *r, *q, *ae, *zG, *xinG, *labG, *g, *pout, *cout;
  float *dR, *dQ, *dAE, *Gt, *dP, *dC, *losses;
  const float* U[kGLayers];
  float* dU[kGLayers];
};

// shared pieces of both loss kernels: info / rec / info_curliness (solver.py:122-123,147, :223,225,245)
__device__ void common_losses(const LossArgs& a, float* red, float& info, float& rec, float& ic) {
  const int B = a.B, t = threadIdx.x, nt = blockDim.x;
  float s = 0.f;
  for (int i = t; i < B * kNoise; i += nt) {
    const int b = i / kNoise, j = i % kNoise;
    const float d = a.q[b * kDOut + 1 + j] - a.zG[i];
    s += d * d;
    a.dQ[b * kDOut + 1 + j] = a.cfg.lambda_info * 2.f * d / (float)(kNoise * B);
  }
  info = block_sum(s, red) / (float)(kNoise * B);
  s = 0.f;
  for (int i = t; i < B * kCode; i += nt) {
    const float d = a.ae[i] - a.code[i];
    s += d * d;
    a.dAE[i] = a.cfg.lambda_rec * 2.f * d / (float)(kCode * B);
  }
  rec = block_sum(s, red) / (float)(kCode * B);
  s = 0.f;
  for (int b = t; b < B; b += nt) {
    const float d = a.q[b * kDOut + 1 + kNoise] - a.xinG[b * kXinLd];
    s += d * d;
    a.dQ[b * kDOut + 1 + kNoise] = a.cfg.lambda_info_curliness * 2.f * d / (float)B;
    a.dQ[b * kDOut + kDOut - 1] = 0.f;
  }
  ic = block_sum(s, red) / (float)B;
}

struct LossSmem {
  float red[33];
  float nrm[kMaxBatch];
  float m1[1 + kNoise], m2[1 + kNoise];
};

__device__ void loss_d_body(const LossArgs& a, LossSmem& sm) {
  float* red = sm.red;
  float* nrm = sm.nrm;
  float* m1 = sm.m1;
  float* m2 = sm.m2;
  const int B = a.B, t = threadIdx.x, nt = blockDim.x, lane = t & 31, w = t >> 5, nw = nt >> 5;
  const float invB = 1.f / (float)B;
  // WGAN critic loss (solver.py:195-196)
  float s = 0.f;
  for (int b = t; b < B; b += nt) {
    s += a.q[b * kDOut] - a.r[b * kDOut];
    a.dQ[b * kDOut] = a.cfg.lambda_adv * invB;
    a.dR[b * kDOut] = -a.cfg.lambda_adv * invB;
    a.dR[b * kDOut + kDOut - 1] = 0.f;
  }
  const float adv = block_sum(s, red) * invB;
  // gradient penalty (solver.py:204-216): per-sample L2 norm of d out_hat / d x_hat
  for (int b = w; b < B; b += nw) {
    float v = 0.f;
    for (int i = lane; i < kCode; i += 32) { const float x = a.g[b * kCode + i]; v += x * x; }
    v = warp_sum(v);
    if (lane == 0) nrm[b] = sqrtf(v);
  }
  __syncthreads();
  s = 0.f;
  for (int b = t; b < B; b += nt) { const float d = nrm[b] - 1.f; s += d * d; }
  const float gp = block_sum(s, red) * invB;
  for (int i = t; i < B * kCode; i += nt) {
    const float n = nrm[i / kCode];
    a.Gt[i] = a.cfg.lambda_gp * 2.f * invB * (n - 1.f) / n * a.g[i];
  }
  float info, rec, ic;
  common_losses(a, red, info, rec, ic);
  // moment losses on cat([noise_curliness, noise]) of the encoder outputs (solver.py:233-242)
  for (int c = w; c < 1 + kNoise; c += nw) {   // one warp per column (the persistent form runs this with 8 warps)
    const int col = c == 0 ? 1 + kNoise : c;
    float s1 = 0.f, s2 = 0.f;
    for (int b = lane; b < B; b += 32) { const float v = a.r[b * kDOut + col]; s1 += v; s2 += v * v; }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (lane == 0) { m1[c] = s1 * invB; m2[c] = s2 * invB; }
  }
  __syncthreads();
  float M1 = 0.f, M2 = 0.f;
  for (int c = 0; c < 1 + kNoise; ++c) { M1 += m1[c] * m1[c]; M2 += (m2[c] - 1.f) * (m2[c] - 1.f); }
  const float nc = (float)(1 + kNoise);
  M1 /= nc; M2 /= nc;
  for (int i = t; i < B * (1 + kNoise); i += nt) {
    const int b = i / (1 + kNoise), c = i % (1 + kNoise);
    const int col = c == 0 ? 1 + kNoise : c;
    const float v = a.r[b * kDOut + col];
    a.dR[b * kDOut + col] = a.cfg.lambda_moment_1 * 2.f * m1[c] * invB / nc +
                            a.cfg.lambda_moment_2 * 2.f * (m2[c] - 1.f) * 2.f * v * invB / nc;
  }
  if (t == 0) {
    float* L = a.losses;
    for (int i = 0; i < CHB_CTT_NUM_LOSSES; ++i) L[i] = 0.f;
    L[CHB_CTT_L_ADV] = adv; L[CHB_CTT_L_GP] = gp; L[CHB_CTT_L_INFO] = info; L[CHB_CTT_L_REC] = rec;
    L[CHB_CTT_L_MOMENT_1] = M1; L[CHB_CTT_L_MOMENT_2] = M2; L[CHB_CTT_L_INFO_CURLINESS] = ic;
    L[CHB_CTT_L_TOTAL] = a.cfg.lambda_adv * adv + a.cfg.lambda_gp * gp + a.cfg.lambda_info * info +
                         a.cfg.lambda_rec * rec + a.cfg.lambda_moment_1 * M1 + a.cfg.lambda_moment_2 * M2 +
                         a.cfg.lambda_info_curliness * ic;
  }
}
__global__ void __launch_bounds__(1024) loss_d_kernel(const LossArgs a) {
  __shared__ LossSmem sm;
  loss_d_body(a, sm);
}

__device__ void loss_g_body(const LossArgs& a, LossSmem& sm) {
  float* red = sm.red;
  const int B = a.B, t = threadIdx.x, nt = blockDim.x;
  const float invB = 1.f / (float)B;
  float s = 0.f;
  for (int b = t; b < B; b += nt) {
    s += a.q[b * kDOut];
    a.dQ[b * kDOut] = -a.cfg.lambda_adv * invB;
  }
  const float adv = -block_sum(s, red) * invB;  // solver.py:175-176
  float info, rec, ic;
  common_losses(a, red, info, rec, ic);
  // rgb / pca_std through the frozen predictor (solver.py:125-140)
  s = 0.f;
  for (int i = t; i < B * 3; i += nt) {
    const int b = i / 3, j = i % 3;
    const float d = a.pout[b * kPOut + j] - a.xinG[b * kXinLd + 1 + j];
    s += d * d;
    a.dP[b * kPOut + j] = a.cfg.lambda_rgb * 2.f * d / (float)(3 * B);
  }
  const float rgb = block_sum(s, red) / (float)(3 * B);
  s = 0.f;
  for (int b = t; b < B; b += nt) {
    const float d = a.pout[b * kPOut + 3] - a.xinG[b * kXinLd + 4];
    s += d * d;
    a.dP[b * kPOut + 3] = a.cfg.lambda_pca_std * 2.f * d * invB;
  }
  const float pca = block_sum(s, red) * invB;
  // weighted BCE of the frozen curliness classifier (solver.py:142-156, curliness_with_weight)
  s = 0.f;
  for (int b = t; b < B; b += nt) s += fabsf(a.xinG[b * kXinLd]);
  const float sw = block_sum(s, red);
  s = 0.f;
  for (int b = t; b < B; b += nt) {
    const float wgt = fabsf(a.xinG[b * kXinLd]) / sw * (float)B;
    const float tg = a.labG[b] * 0.5f + 0.5f;
    const float p = 1.f / (1.f + expf(-a.cout[b]));
    const float lp = fmaxf(logf(p), -100.f), lq = fmaxf(logf(1.f - p), -100.f);
    s += -wgt * (tg * lp + (1.f - tg) * lq);
    a.dC[b] = a.cfg.lambda_cls_curliness * wgt * (p - tg) * invB;
  }
  const float cls = block_sum(s, red) * invB;
  // orthogonal regulariser of the subspace bases (model_eigengan.py:27-31, :85-89), n_basis = 2
  float orth = 0.f;
  for (int l = 0; l < kGLayers; ++l) {
    const float* U = a.U[l];
    float s00 = 0.f, s01 = 0.f, s11 = 0.f;
    for (int i = t; i < kHid; i += nt) { const float u0 = U[i], u1 = U[kHid + i]; s00 += u0 * u0; s01 += u0 * u1; s11 += u1 * u1; }
    const float e00 = block_sum(s00, red) - 1.f, e01 = block_sum(s01, red), e11 = block_sum(s11, red) - 1.f;
    orth += (e00 * e00 + 2.f * e01 * e01 + e11 * e11) * 0.25f;
    for (int i = t; i < kHid; i += nt) {
      const float u0 = U[i], u1 = U[kHid + i];
      a.dU[l][i] += a.cfg.lambda_orthogonal * (e00 * u0 + e01 * u1);
      a.dU[l][kHid + i] += a.cfg.lambda_orthogonal * (e01 * u0 + e11 * u1);
    }
  }
  if (t == 0) {
    float* L = a.losses;
    for (int i = 0; i < CHB_CTT_NUM_LOSSES; ++i) L[i] = 0.f;
    L[CHB_CTT_L_ADV] = adv; L[CHB_CTT_L_INFO] = info; L[CHB_CTT_L_REC] = rec; L[CHB_CTT_L_RGB] = rgb;
    L[CHB_CTT_L_PCA_STD] = pca; L[CHB_CTT_L_INFO_CURLINESS] = ic; L[CHB_CTT_L_CLS_CURLINESS] = cls;
    L[CHB_CTT_L_ORTHOGONAL] = orth;
    L[CHB_CTT_L_TOTAL] = a.cfg.lambda_adv * adv + a.cfg.lambda_info * info + a.cfg.lambda_rec * rec +
                         a.cfg.lambda_rgb * rgb + a.cfg.lambda_pca_std * pca + a.cfg.lambda_info_curliness * ic +
                         a.cfg.lambda_cls_curliness * cls + a.cfg.lambda_orthogonal * orth;
  }
}
__global__ void __launch_bounds__(1024) loss_g_kernel(const LossArgs a) {
  __shared__ LossSmem sm;
  loss_g_body(a, sm);
}

// Gradients of the auto-encoder pass w.r.t. the generator's *inputs* that came out of the discriminator
// (noise -> subspace coordinates, noise_curliness -> first input column): they continue into D's real-code pass.
struct GInArgs {
  const float* ds[kGLayers];
  const float* U[kGLayers];
  const float* Lvec[kGLayers];
  const float* Win;  // [256,5]
  float* dR;
  int B;
};
__device__ void g_input_grads_row(const GInArgs& a, int b, float* red) {
  const int t = threadIdx.x;
  float v = a.ds[0][(long)b * kHid + t] * a.Win[t * kGIn + 0];
  v = block_sum(v, red);
  if (t == 0) a.dR[b * kDOut + 1 + kNoise] += v;
  for (int l = 0; l < kGLayers; ++l)
    for (int k = 0; k < kSub; ++k) {
      float u = a.ds[l][(long)b * kHid + t] * a.U[l][k * kHid + t];
      u = block_sum(u, red);
      if (t == 0) a.dR[b * kDOut + 1 + l * kSub + k] += a.Lvec[l][k] * u;
    }
}
__global__ void __launch_bounds__(256) g_input_grads_kernel(const GInArgs a) {
  __shared__ float red[33];
  g_input_grads_row(a, blockIdx.x, red);
}

// dL[l][k] += sum_b z[b, 2l+k] * <ds_l[b,:], U_l[k,:]>   (SubspaceLayer.L, model_eigengan.py:24)
struct LGradArgs {
  const float* ds[kGLayers];
  const float* U[kGLayers];
  float* dL[kGLayers];
  const float* z; int ldz; int B;
};
__device__ void subspace_l_grad_one(const LGradArgs& a, int idx, float* red) {
  const int l = idx / kSub, k = idx % kSub, t = threadIdx.x;
  const float u = a.U[l][k * kHid + t];
  float s = 0.f;
  for (int b = 0; b < a.B; ++b) s += a.z[(long)b * a.ldz + l * kSub + k] * a.ds[l][(long)b * kHid + t] * u;
  s = block_sum(s, red);
  if (t == 0) a.dL[l][k] += s;
}
__global__ void __launch_bounds__(256) subspace_l_grad_kernel(const LGradArgs a) {
  __shared__ float red[33];
  subspace_l_grad_one(a, blockIdx.x, red);
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent form of a sub-step (round 2): the same ~95 operations, in the same order and with the same per-tile
// arithmetic, executed by ONE cooperative kernel that walks an operation list in device memory.  A grid-wide barrier
// separates dependent operations only (the host analyses the read / write ranges when it builds the list), so the
// cost of a dependency drops from a graph-node boundary (~7 us: drain, launch, fill) to an atomic barrier (~1.5 us),
// and independent GEMMs (weight gradients next to the data path, the two generator passes) share a phase.
enum { OP_GEMM = 0, OP_ZERO, OP_PREP, OP_XHAT, OP_LOSS_D, OP_LOSS_G, OP_GIN, OP_LGRAD };
struct XhatArgs { const float *code, *f, *alpha; float* xh; int n; };
struct ZeroArgs { float* p; long n; };
struct Op {
  int type, sync;   // sync: grid barrier before this operation
  int rot, pad_;    // first CTA of this operation's tile loop (spreads the small GEMMs of one phase over the grid)
  union {
    Gemm gemm;
    ZeroArgs zero;
    PrepArgs prep;
    XhatArgs xhat;
    LossArgs loss;
    GInArgs gin;
    LGradArgs lgrad;
  };
};

__device__ __forceinline__ void grid_barrier(unsigned long long* bar, unsigned long long target) {
  // release / acquire at gpu scope (no sequentially-consistent fence: MEMBAR.SC costs microseconds).  bar.sync orders
  // the CTA's writes before thread 0's release (cumulativity); the acquire load is followed by an L1 invalidation
  // (CCTL.IVALL in SASS), so the weak loads the CTA issues after the second bar.sync read the other CTAs' data from L2.
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(bar), "l"(1ull) : "memory");
    unsigned long long seen;
    do {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(bar) : "memory");
    } while (seen < target);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) ctt_persistent_kernel(const Op* __restrict__ ops, int nops,
                                                             unsigned long long* bar, unsigned long long base) {
  __shared__ __align__(16) float As[kGemmKC][kGemmLd];
  __shared__ __align__(16) float Bs[kGemmKC][kGemmLd];
  __shared__ LossSmem lsm;
  __shared__ Op opbuf[2];   // descriptor i + 1 is fetched while operation i (and its barrier) run
  const int G = gridDim.x, cta = blockIdx.x;
  unsigned long long target = base;
  constexpr int kWords = (int)(sizeof(Op) / 4);
  if (threadIdx.x < kWords)
    reinterpret_cast<unsigned*>(&opbuf[0])[threadIdx.x] = __ldg(reinterpret_cast<const unsigned*>(ops) + threadIdx.x);
  for (int i = 0; i < nops; ++i) {
    __syncthreads();   // descriptor i has landed; everyone is done with descriptor i - 1 (the buffer refilled below)
    const Op& op = opbuf[i & 1];
    if (i + 1 < nops && threadIdx.x < kWords)
      reinterpret_cast<unsigned*>(&opbuf[(i + 1) & 1])[threadIdx.x] =
          __ldg(reinterpret_cast<const unsigned*>(ops + i + 1) + threadIdx.x);
    if (op.sync) {
      target += (unsigned long long)G;
      grid_barrier(bar, target);
    }
    switch (op.type) {
      case OP_GEMM: {
        const int ntx = (op.gemm.N + 31) / 32, nty = (op.gemm.M + 31) / 32;
        for (int tile = (cta - op.rot + G) % G; tile < ntx * nty; tile += G)
          gemm_tile(op.gemm, tile % ntx, tile / ntx, As, Bs);
        break;
      }
      case OP_ZERO:
        for (long j = (long)cta * blockDim.x + threadIdx.x; j < op.zero.n; j += (long)G * blockDim.x) op.zero.p[j] = 0.f;
        break;
      case OP_PREP:
        for (int b = cta * blockDim.x + threadIdx.x; b < op.prep.B; b += G * blockDim.x) prep_row(op.prep, b);
        break;
      case OP_XHAT:
        for (int j = cta * blockDim.x + threadIdx.x; j < op.xhat.n; j += G * blockDim.x) {
          const float al = op.xhat.alpha[j / kCode];
          op.xhat.xh[j] = al * op.xhat.code[j] + (1.f - al) * op.xhat.f[j];
        }
        break;
      case OP_LOSS_D:
        if (cta == 0) loss_d_body(op.loss, lsm);
        break;
      case OP_LOSS_G:
        if (cta == 0) loss_g_body(op.loss, lsm);
        break;
      case OP_GIN:
        for (int b = cta; b < op.gin.B; b += G) g_input_grads_row(op.gin, b, lsm.red);
        break;
      case OP_LGRAD:
        for (int idx = cta; idx < kGLayers * kSub; idx += G) subspace_l_grad_one(op.lgrad, idx, lsm.red);
        break;
    }
  }
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long n, float b1, float b2, float eps, float step_size,
                            float bc2_sqrt) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i];
  const float mi = m[i] + (gi - m[i]) * (1.f - b1);          // exp_avg.lerp_(grad, 1 - beta1)
  const float vi = v[i] * b2 + (1.f - b2) * gi * gi;          // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  m[i] = mi;
  v[i] = vi;
  p[i] -= step_size * mi / (sqrtf(vi) / bc2_sqrt + eps);     // param.addcdiv_(exp_avg, denom, value=-step_size)
}

// ---------------------------------------------------------------------------------------------------------------
struct Entry { std::string name; int64_t offset, numel; int group; };

struct Mlp {  // plain chain: Linear -> lrelu -> ... -> Linear
  int nl = 0;         // number of Linear layers
  int in_dim = 0;
  int dims[6] = {};
  int64_t w[6] = {}, b[6] = {};  // offsets (floats) inside the net's region
};

}  // namespace ctt
}  // namespace chb

using namespace chb::ctt;

struct chb_cttrain {
  chb_cttrain_config cfg;
  int B = 0;
  std::vector<Entry> table;
  int64_t nD = 0, nG = 0, nF = 0;
  Mlp D, P, C;
  // generator offsets (inside the G region)
  int64_t g_win = 0, g_bin = 0, g_w[kGLayers] = {}, g_b[kGLayers] = {}, g_U[kGLayers] = {}, g_L[kGLayers] = {},
          g_mu[kGLayers] = {};
  int64_t off_P = 0, off_C = 0;  // inside the frozen region
  float* state = nullptr;
  char* ws = nullptr;
  int64_t ws_bytes = 0;
  long step_count[2] = {0, 0};
  cudaGraphExec_t exec[2] = {nullptr, nullptr};
  int launches[2] = {0, 0};
  // persistent form (cfg.use_graph == 2): operation lists in library-owned device memory, one barrier counter
  void* dev_ops[2] = {nullptr, nullptr};
  int n_ops[2] = {0, 0}, n_sync[2] = {0, 0};
  unsigned long long* dev_bar = nullptr;
  unsigned long long bar_total = 0;
  int persist_grid = 0;
  bool failed = false;
  // workspace pointers
  float *code, *rgb, *pca, *noise, *nc, *label, *alpha, *ones, *e0;
  int *p1, *p2, *p3, *flag;
  float *zR[4], *r, *xinA, *LzA[kGLayers], *sA[kGLayers], *ae;
  float *xinG, *zG, *labG, *LzG[kGLayers], *sG[kGLayers], *f;
  float *zF[4], *q, *xh, *zH[4], *uH[4], *g, *zP[3], *pout, *zC[3], *cout;
  float *dR, *dQ, *dAE, *Gt, *df, *dP, *dC, *dz[4], *dzc[3], *dsA[kGLayers], *dsG[kGLayers], *dt[4], *losses;

  float* par(int group) const { return state + (group == CHB_CTT_D ? 0 : nD); }
  float* grad(int group) const { return state + nD + nG + (group == CHB_CTT_D ? 0 : nD); }
  float* adam_m(int group) const { return state + 2 * (nD + nG) + (group == CHB_CTT_D ? 0 : nD); }
  float* adam_v(int group) const { return state + 3 * (nD + nG) + (group == CHB_CTT_D ? 0 : nD); }
  float* frozen() const { return state + 4 * (nD + nG); }
};

namespace {

int64_t add_tensor(chb_cttrain* t, const std::string& name, int64_t numel, int group, int64_t& cursor) {
  const int64_t off = cursor;
  t->table.push_back({name, off, numel, group});
  cursor += numel;
  return off;
}

void build_mlp(chb_cttrain* t, Mlp& m, const char* prefix, int group, int in_dim, int hidden, int n_hidden, int out_dim,
               int64_t& cursor) {
  m.nl = n_hidden + 1;
  m.in_dim = in_dim;
  for (int l = 0; l < m.nl; ++l) {
    const int id = l == 0 ? in_dim : hidden, od = l < n_hidden ? hidden : out_dim;
    m.dims[l] = od;
    const std::string base = std::string(prefix) + "net." + std::to_string(l) + ".fc.";
    m.w[l] = add_tensor(t, base + "weight", (int64_t)od * id, group, cursor);
    m.b[l] = add_tensor(t, base + "bias", od, group, cursor);
  }
}

struct Rec {  // records launches on a stream (plain or under capture), or operations of the persistent form
  cudaStream_t s;
  int n = 0;
  cudaError_t err = cudaSuccess;
  std::vector<Op>* ops = nullptr;
  void check() { if (err == cudaSuccess) err = cudaGetLastError(); }
};

Op new_op(int type) {
  Op o;
  memset(&o, 0, sizeof o);
  o.type = type;
  o.sync = 1;
  return o;
}

void gemm(Rec& rc, const Gemm& g) {
  rc.n++;
  if (rc.ops) {
    Op o = new_op(OP_GEMM);
    o.gemm = g;
    rc.ops->push_back(o);
    return;
  }
  dim3 grid((g.N + 31) / 32, (g.M + 31) / 32);
  gemm_kernel<<<grid, 256, 0, rc.s>>>(g);
  rc.check();
}

// Y[M,N] (+)= act?(X)[M,K] W[N,K]^T + bias, optionally * lrelu'(mask)
void lin_fwd(Rec& rc, const float* X, long ldx, int x_act, const float* W, int K, int N, const float* bias, float* Y,
             long ldy, int M, const float* mask = nullptr, long mask_ld = 0, int acc = 0) {
  Gemm g{};
  g.A = X; g.a_sm = ldx; g.a_sk = 1; g.B = W; g.b_sk = 1; g.b_sn = K; g.C = Y; g.c_sm = ldy;
  g.M = M; g.N = N; g.K = K; g.bias = bias; g.mask = mask; g.mask_sm = mask_ld; g.a_act = x_act; g.accumulate = acc;
  gemm(rc, g);
}
// dX[M,K] (+)= dY[M,N] W[N,K], optionally * lrelu'(pre)
void lin_bwd_data(Rec& rc, const float* dY, long ldy, const float* W, int K, int N, float* dX, long ldx, int M,
                  const float* pre, long ldpre, int acc) {
  Gemm g{};
  g.A = dY; g.a_sm = ldy; g.a_sk = 1; g.B = W; g.b_sk = K; g.b_sn = 1; g.C = dX; g.c_sm = ldx;
  g.M = M; g.N = K; g.K = N; g.mask = pre; g.mask_sm = ldpre; g.accumulate = acc;
  gemm(rc, g);
}
// dW[N,K] += dY[M,N]^T act?(X)[M,K];  db[N] += column sums of dY
void lin_bwd_w(Rec& rc, const float* dY, long ldy, const float* X, long ldx, int x_act, float* dW, float* db, int M, int N,
               int K) {
  Gemm g{};
  g.A = dY; g.a_sm = 1; g.a_sk = ldy; g.B = X; g.b_sk = ldx; g.b_sn = 1; g.C = dW; g.c_sm = K;
  g.M = N; g.N = K; g.K = M; g.b_act = x_act; g.accumulate = 1;
  gemm(rc, g);
  if (db) {
    Gemm h{};
    h.A = dY; h.a_sm = 1; h.a_sk = ldy; h.B = nullptr; h.b_sk = 1; h.b_sn = 1; h.C = db; h.c_sm = 1;
    h.M = N; h.N = 1; h.K = M; h.accumulate = 1;
    gemm(rc, h);
  }
}

void mlp_fwd(Rec& rc, const Mlp& m, const float* P, const float* X, long ldx, float* const* z, float* out, int B,
             bool skip_last = false) {
  for (int l = 0; l < m.nl; ++l) {
    if (l == m.nl - 1 && skip_last) break;
    const int id = l == 0 ? m.in_dim : m.dims[l - 1];
    const float* xin = l == 0 ? X : z[l - 1];
    float* y = l == m.nl - 1 ? out : z[l];
    lin_fwd(rc, xin, l == 0 ? ldx : id, l > 0, P + m.w[l], id, m.dims[l], P + m.b[l], y, m.dims[l], B);
  }
}

// backward of the chain from dOut; dz[l] receives the gradient w.r.t. pre-activation l (l < nl-1)
void mlp_bwd(Rec& rc, const Mlp& m, const float* P, float* G, const float* dOut, const float* X, long ldx,
             float* const* z, float* const* dz, float* dX, int dX_acc, int B) {
  for (int l = m.nl - 1; l >= 0; --l) {
    const int id = l == 0 ? m.in_dim : m.dims[l - 1], od = m.dims[l];
    const float* dy = l == m.nl - 1 ? dOut : dz[l];
    const float* xin = l == 0 ? X : z[l - 1];
    if (G) lin_bwd_w(rc, dy, od, xin, l == 0 ? ldx : id, l > 0, G + m.w[l], G + m.b[l], B, od, id);
    if (l > 0)
      lin_bwd_data(rc, dy, od, P + m.w[l], id, od, dz[l - 1], id, B, z[l - 1], id, 0);
    else if (dX)
      lin_bwd_data(rc, dy, od, P + m.w[l], id, od, dX, id, B, nullptr, 0, dX_acc);
  }
}

void gen_fwd(Rec& rc, const chb_cttrain* t, const float* xin, float* const* Lz, float* const* s, float* code) {
  const float* P = t->par(CHB_CTT_G);
  const int B = t->B;
  lin_fwd(rc, xin, kXinLd, 0, P + t->g_win, kGIn, kHid, P + t->g_bin, s[0], kHid, B);
  for (int l = 0; l < kGLayers; ++l) {
    // s_l += (L * z_l) @ U_l + mu_l   (model_eigengan.py:24, :77-78)
    Gemm g{};
    g.A = Lz[l]; g.a_sm = kSub; g.a_sk = 1; g.B = P + t->g_U[l]; g.b_sk = kHid; g.b_sn = 1; g.C = s[l]; g.c_sm = kHid;
    g.M = B; g.N = kHid; g.K = kSub; g.bias = P + t->g_mu[l]; g.accumulate = 1;
    gemm(rc, g);
    const int od = l < kGLayers - 1 ? kHid : kCode;
    float* y = l < kGLayers - 1 ? s[l + 1] : code;
    lin_fwd(rc, s[l], kHid, 1, P + t->g_w[l], kHid, od, P + t->g_b[l], y, od, B);  // Linear(LeakyReLU(x)) (:50-53, :79)
  }
}

void gen_bwd(Rec& rc, chb_cttrain* t, const float* dcode, const float* xin, float* const* Lz, const float* z, int ldz,
             float* const* s, float* const* ds, bool wg) {
  const float* P = t->par(CHB_CTT_G);
  float* G = wg ? t->grad(CHB_CTT_G) : nullptr;
  const int B = t->B;
  for (int l = kGLayers - 1; l >= 0; --l) {
    const int od = l < kGLayers - 1 ? kHid : kCode;
    const float* dout = l < kGLayers - 1 ? ds[l + 1] : dcode;
    if (G) lin_bwd_w(rc, dout, od, s[l], kHid, 1, G + t->g_w[l], G + t->g_b[l], B, od, kHid);
    lin_bwd_data(rc, dout, od, P + t->g_w[l], kHid, od, ds[l], kHid, B, s[l], kHid, 0);
  }
  if (!G) return;
  lin_bwd_w(rc, ds[0], kHid, xin, kXinLd, 0, G + t->g_win, G + t->g_bin, B, kHid, kGIn);
  LGradArgs la{};
  for (int l = 0; l < kGLayers; ++l) {
    // dU_l += (L z_l)^T ds_l ; dmu_l += column sums of ds_l
    Gemm g{};
    g.A = Lz[l]; g.a_sm = 1; g.a_sk = kSub; g.B = ds[l]; g.b_sk = kHid; g.b_sn = 1; g.C = G + t->g_U[l]; g.c_sm = kHid;
    g.M = kSub; g.N = kHid; g.K = B; g.accumulate = 1;
    gemm(rc, g);
    Gemm h{};
    h.A = t->ones; h.a_sm = 0; h.a_sk = 1; h.B = ds[l]; h.b_sk = kHid; h.b_sn = 1; h.C = G + t->g_mu[l]; h.c_sm = kHid;
    h.M = 1; h.N = kHid; h.K = B; h.accumulate = 1;
    gemm(rc, h);
    la.ds[l] = ds[l]; la.U[l] = P + t->g_U[l]; la.dL[l] = G + t->g_L[l];
  }
  la.z = z; la.ldz = ldz; la.B = B;
  rc.n++;
  if (rc.ops) {
    Op o = new_op(OP_LGRAD);
    o.lgrad = la;
    rc.ops->push_back(o);
    return;
  }
  subspace_l_grad_kernel<<<kGLayers * kSub, 256, 0, rc.s>>>(la);
  rc.check();
}

void record_step(chb_cttrain* t, int which, Rec& rc) {
  const int B = t->B;
  const float* PD = t->par(CHB_CTT_D);
  const float* PG = t->par(CHB_CTT_G);
  const float* PF = t->frozen();
  const int64_t ng = which == CHB_CTT_D ? t->nD : t->nG;
  if (rc.ops) {
    Op o = new_op(OP_ZERO);
    o.zero.p = t->grad(which);
    o.zero.n = (long)ng;
    rc.ops->push_back(o);
  } else if (rc.err == cudaSuccess) {
    rc.err = cudaMemsetAsync(t->grad(which), 0, ng * sizeof(float), rc.s);
  }
  // ---- Solver.forward (solver.py:85-117)
  mlp_fwd(rc, t->D, PD, t->code, kCode, t->zR, t->r, B);
  PrepArgs pa{};
  pa.r = t->r; pa.rgb = t->rgb; pa.pca = t->pca; pa.noise = t->noise; pa.nc = t->nc; pa.label = t->label;
  pa.p1 = t->p1; pa.p2 = t->p2; pa.p3 = t->p3; pa.flag = t->flag;
  pa.xinA = t->xinA; pa.xinG = t->xinG; pa.zG = t->zG; pa.labG = t->labG; pa.B = B;
  for (int l = 0; l < kGLayers; ++l) { pa.Lvec[l] = PG + t->g_L[l]; pa.LzA[l] = t->LzA[l]; pa.LzG[l] = t->LzG[l]; }
  rc.n++;
  if (rc.ops) {
    Op o = new_op(OP_PREP);
    o.prep = pa;
    rc.ops->push_back(o);
  } else {
    prep_kernel<<<(B + 127) / 128, 128, 0, rc.s>>>(pa);
    rc.check();
  }
  gen_fwd(rc, t, t->xinA, t->LzA, t->sA, t->ae);
  gen_fwd(rc, t, t->xinG, t->LzG, t->sG, t->f);
  mlp_fwd(rc, t->D, PD, t->f, kCode, t->zF, t->q, B);

  LossArgs la{};
  la.cfg = t->cfg; la.B = B; la.code = t->code; la.r = t->r; la.q = t->q; la.ae = t->ae; la.zG = t->zG; la.xinG = t->xinG;
  la.labG = t->labG; la.g = t->g; la.pout = t->pout; la.cout = t->cout; la.dR = t->dR; la.dQ = t->dQ; la.dAE = t->dAE;
  la.Gt = t->Gt; la.dP = t->dP; la.dC = t->dC; la.losses = t->losses;
  for (int l = 0; l < kGLayers; ++l) { la.U[l] = PG + t->g_U[l]; la.dU[l] = t->grad(CHB_CTT_G) + t->g_U[l]; }

  if (which == CHB_CTT_D) {
    // ---- forward_general_dis: x_hat, D(x_hat), d out_hat / d x_hat with the graph kept (solver.py:198-216)
    rc.n++;
    if (rc.ops) {
      Op o = new_op(OP_XHAT);
      o.xhat.code = t->code; o.xhat.f = t->f; o.xhat.alpha = t->alpha; o.xhat.xh = t->xh; o.xhat.n = B * kCode;
      rc.ops->push_back(o);
    } else {
      xhat_kernel<<<(B * kCode + 255) / 256, 256, 0, rc.s>>>(t->code, t->f, t->alpha, t->xh, B * kCode);
      rc.check();
    }
    mlp_fwd(rc, t->D, PD, t->xh, kCode, t->zH, nullptr, B, /*skip_last=*/true);
    mlp_bwd(rc, t->D, PD, nullptr, t->e0, t->xh, kCode, t->zH, t->uH, t->g, 0, B);
    rc.n++;
    if (rc.ops) {
      Op o = new_op(OP_LOSS_D);
      o.loss = la;
      rc.ops->push_back(o);
    } else {
      loss_d_kernel<<<1, 1024, 0, rc.s>>>(la);
      rc.check();
    }
    float* GD = t->grad(CHB_CTT_D);
    // D(fake) pass: parameter gradients only (the generator is not stepped here)
    mlp_bwd(rc, t->D, PD, GD, t->dQ, t->f, kCode, t->zF, t->dz, nullptr, 0, B);
    // rec loss -> generator (auto-encoder pass) -> its discriminator-produced inputs -> D(real) pass
    gen_bwd(rc, t, t->dAE, t->xinA, t->LzA, nullptr, 0, t->sA, t->dsA, /*wg=*/false);
    GInArgs ga{};
    for (int l = 0; l < kGLayers; ++l) { ga.ds[l] = t->dsA[l]; ga.U[l] = PG + t->g_U[l]; ga.Lvec[l] = PG + t->g_L[l]; }
    ga.Win = PG + t->g_win; ga.dR = t->dR; ga.B = B;
    if (rc.ops) {
      Op o = new_op(OP_GIN);
      o.gin = ga;
      rc.ops->push_back(o);
    } else {
      g_input_grads_kernel<<<B, 256, 0, rc.s>>>(ga);
      rc.check();
    }
    rc.n++;
    mlp_bwd(rc, t->D, PD, GD, t->dR, t->code, kCode, t->zR, t->dz, nullptr, 0, B);
    // gradient penalty, second-order pass.  With u_l = lrelu'(z_l) * (u_{l+1} W_{l+1}) and g = u_1 W_1 the masks are
    // piecewise constant, so  dW_l += u_l^T dt_{l-1},  dt_l = (dt_{l-1} W_l^T) * lrelu'(z_l),  dt_0 = dGP/dg.
    const float* xprev = t->Gt;
    for (int l = 0; l < t->D.nl - 1; ++l) {
      const int id = l == 0 ? kCode : kHid;
      lin_bwd_w(rc, t->uH[l], kHid, xprev, id, 0, GD + t->D.w[l], nullptr, B, kHid, id);
      lin_fwd(rc, xprev, id, 0, PD + t->D.w[l], id, kHid, nullptr, t->dt[l], kHid, B, t->zH[l], kHid);
      xprev = t->dt[l];
    }
    const int last = t->D.nl - 1;  // row 0 of the output layer: u_4 = lrelu'(z_4) * W_5[0,:]
    Gemm h{};
    h.A = t->ones; h.a_sm = 0; h.a_sk = 1; h.B = xprev; h.b_sk = kHid; h.b_sn = 1; h.C = GD + t->D.w[last]; h.c_sm = kHid;
    h.M = 1; h.N = kHid; h.K = B; h.accumulate = 1;
    gemm(rc, h);
  } else {
    mlp_fwd(rc, t->P, PF + t->off_P, t->f, kCode, t->zP, t->pout, B);
    mlp_fwd(rc, t->C, PF + t->off_C, t->f, kCode, t->zC, t->cout, B);
    rc.n++;
    if (rc.ops) {
      Op o = new_op(OP_LOSS_G);
      o.loss = la;
      rc.ops->push_back(o);
    } else {
      loss_g_kernel<<<1, 1024, 0, rc.s>>>(la);
      rc.check();
    }
    // d loss / d fake code: through D (adv, info, info_curliness), the rgb predictor and the curliness classifier
    mlp_bwd(rc, t->D, PD, nullptr, t->dQ, t->f, kCode, t->zF, t->dz, t->df, 0, B);
    mlp_bwd(rc, t->P, PF + t->off_P, nullptr, t->dP, t->f, kCode, t->zP, t->dz, t->df, 1, B);
    mlp_bwd(rc, t->C, PF + t->off_C, nullptr, t->dC, t->f, kCode, t->zC, t->dzc, t->df, 1, B);
    gen_bwd(rc, t, t->df, t->xinG, t->LzG, t->zG, kNoise, t->sG, t->dsG, /*wg=*/true);
    gen_bwd(rc, t, t->dAE, t->xinA, t->LzA, t->r + 1, kDOut, t->sA, t->dsA, /*wg=*/true);
  }
}

// ---- persistent form: dependency analysis over byte ranges, then upload
struct Range { const char* lo; const char* hi; };
static void add_range(std::vector<Range>& v, const void* p, long floats) {
  if (p && floats > 0) v.push_back({static_cast<const char*>(p), static_cast<const char*>(p) + floats * 4});
}
static bool overlaps(const std::vector<Range>& a, const std::vector<Range>& b) {
  for (const Range& x : a)
    for (const Range& y : b)
      if (x.lo < y.hi && y.lo < x.hi) return true;
  return false;
}

// Byte ranges an operation reads / writes (hulls of its strided accesses).  They drive both schedules: the grid
// barriers of the persistent kernel and the edges of the explicit CUDA graph.
static void op_ranges(const Op& o, int B, std::vector<Range>& r, std::vector<Range>& w) {
  auto rw = [&](const void* p, long n) { add_range(r, p, n); add_range(w, p, n); };
  switch (o.type) {
    case OP_GEMM: {
      const Gemm& g = o.gemm;
      add_range(r, g.A, (long)(g.M - 1) * g.a_sm + (long)(g.K - 1) * g.a_sk + 1);
      if (g.B) add_range(r, g.B, (long)(g.K - 1) * g.b_sk + (long)(g.N - 1) * g.b_sn + 1);
      add_range(r, g.bias, g.N);
      add_range(r, g.mask, (long)(g.M - 1) * g.mask_sm + g.N);
      add_range(w, g.C, (long)(g.M - 1) * g.c_sm + g.N);
      if (g.accumulate) add_range(r, g.C, (long)(g.M - 1) * g.c_sm + g.N);
      break;
    }
    case OP_ZERO:
      add_range(w, o.zero.p, o.zero.n);
      break;
    case OP_PREP: {
      const PrepArgs& a = o.prep;
      add_range(r, a.r, (long)B * kDOut); add_range(r, a.rgb, (long)B * 3); add_range(r, a.pca, B);
      add_range(r, a.noise, (long)B * kNoise); add_range(r, a.nc, B); add_range(r, a.label, B);
      add_range(r, a.p1, B); add_range(r, a.p2, B); add_range(r, a.p3, B); add_range(r, a.flag, 1);
      add_range(w, a.xinA, (long)B * kXinLd); add_range(w, a.xinG, (long)B * kXinLd);
      add_range(w, a.zG, (long)B * kNoise); add_range(w, a.labG, B);
      for (int l = 0; l < kGLayers; ++l) {
        add_range(r, a.Lvec[l], kSub);
        add_range(w, a.LzA[l], (long)B * kSub); add_range(w, a.LzG[l], (long)B * kSub);
      }
      break;
    }
    case OP_XHAT:
      add_range(r, o.xhat.code, o.xhat.n); add_range(r, o.xhat.f, o.xhat.n); add_range(r, o.xhat.alpha, B);
      add_range(w, o.xhat.xh, o.xhat.n);
      break;
    case OP_LOSS_D:
    case OP_LOSS_G: {
      const LossArgs& a = o.loss;
      add_range(r, a.code, (long)B * kCode); add_range(r, a.r, (long)B * kDOut); add_range(r, a.q, (long)B * kDOut);
      add_range(r, a.ae, (long)B * kCode); add_range(r, a.zG, (long)B * kNoise); add_range(r, a.xinG, (long)B * kXinLd);
      add_range(r, a.labG, B); add_range(r, a.g, (long)B * kCode); add_range(r, a.pout, (long)B * kPOut);
      add_range(r, a.cout, B);
      rw(a.dR, (long)B * kDOut); rw(a.dQ, (long)B * kDOut); rw(a.dAE, (long)B * kCode); rw(a.Gt, (long)B * kCode);
      rw(a.dP, (long)B * kPOut); rw(a.dC, B); rw(a.losses, 64);
      for (int l = 0; l < kGLayers; ++l) {
        add_range(r, a.U[l], (long)kSub * kHid);
        if (o.type == OP_LOSS_G) rw(a.dU[l], (long)kSub * kHid);
      }
      break;
    }
    case OP_GIN: {
      const GInArgs& a = o.gin;
      for (int l = 0; l < kGLayers; ++l) {
        add_range(r, a.ds[l], (long)B * kHid); add_range(r, a.U[l], (long)kSub * kHid); add_range(r, a.Lvec[l], kSub);
      }
      add_range(r, a.Win, (long)kHid * kGIn);
      rw(a.dR, (long)B * kDOut);
      break;
    }
    case OP_LGRAD: {
      const LGradArgs& a = o.lgrad;
      for (int l = 0; l < kGLayers; ++l) {
        add_range(r, a.ds[l], (long)B * kHid); add_range(r, a.U[l], (long)kSub * kHid);
        rw(a.dL[l], kSub);
      }
      add_range(r, a.z, (long)(B - 1) * a.ldz + kNoise);
      break;
    }
  }
}

static bool ops_conflict(const std::vector<Range>& r1, const std::vector<Range>& w1, const std::vector<Range>& r2,
                         const std::vector<Range>& w2) {
  return overlaps(w1, r2) || overlaps(w1, w2) || overlaps(r1, w2);
}

// Sets Op::sync / Op::rot: a barrier before an operation that conflicts with anything since the last barrier.
void schedule_ops(std::vector<Op>& ops, int G, int B, int* n_sync) {
  std::vector<Range> pr, pw;   // reads / writes of the current phase
  long tiles_in_phase = 0;
  *n_sync = 0;
  for (size_t i = 0; i < ops.size(); ++i) {
    Op& o = ops[i];
    std::vector<Range> r, w;
    op_ranges(o, B, r, w);
    long tiles = 1;
    if (o.type == OP_GEMM) tiles = (long)((o.gemm.N + 31) / 32) * ((o.gemm.M + 31) / 32);
    const bool conflict = i > 0 && ops_conflict(pr, pw, r, w);
    if (conflict) {
      o.sync = 1;
      ++*n_sync;
      pr.clear(); pw.clear();
      tiles_in_phase = 0;
    } else {
      o.sync = 0;
    }
    o.rot = (int)(tiles_in_phase % G);
    tiles_in_phase += tiles;
    pr.insert(pr.end(), r.begin(), r.end());
    pw.insert(pw.end(), w.begin(), w.end());
  }
}

// The same operation list as an explicit CUDA graph: one kernel (or memset) node per operation, an edge from every
// earlier operation it conflicts with.  Stream capture (round 1) could only produce a linear chain of ~95 nodes;
// with real dependencies the weight-gradient GEMMs, the two generator passes and the three discriminator passes run
// beside the data path, and the dependent chain shrinks to the critical path.
cudaError_t build_graph(chb_cttrain* t, int which) {
  std::vector<Op> ops;
  Rec rc{nullptr};
  rc.ops = &ops;
  record_step(t, which, rc);
  const int B = t->B;
  std::vector<std::vector<Range>> R(ops.size()), W(ops.size());
  for (size_t i = 0; i < ops.size(); ++i) op_ranges(ops[i], B, R[i], W[i]);
  cudaGraph_t gr = nullptr;
  cudaError_t e = cudaGraphCreate(&gr, 0);
  std::vector<cudaGraphNode_t> node(ops.size());
  int edges = 0;
  for (size_t i = 0; i < ops.size() && e == cudaSuccess; ++i) {
    std::vector<cudaGraphNode_t> deps;
    for (size_t j = 0; j < i; ++j)
      if (ops_conflict(R[j], W[j], R[i], W[i])) deps.push_back(node[j]);
    edges += (int)deps.size();
    Op& o = ops[i];
    if (o.type == OP_ZERO) {
      cudaMemsetParams mp{};
      mp.dst = o.zero.p; mp.value = 0; mp.elementSize = 4; mp.width = (size_t)o.zero.n; mp.height = 1; mp.pitch = 0;
      e = cudaGraphAddMemsetNode(&node[i], gr, deps.data(), deps.size(), &mp);
      continue;
    }
    cudaKernelNodeParams kp{};
    void* args[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    int xn = 0;
    kp.blockDim = dim3(256);
    kp.gridDim = dim3(1);
    switch (o.type) {
      case OP_GEMM:
        kp.func = reinterpret_cast<void*>(gemm_kernel);
        kp.gridDim = dim3((o.gemm.N + 31) / 32, (o.gemm.M + 31) / 32);
        args[0] = &o.gemm;
        break;
      case OP_PREP:
        kp.func = reinterpret_cast<void*>(prep_kernel);
        kp.gridDim = dim3((B + 127) / 128); kp.blockDim = dim3(128);
        args[0] = &o.prep;
        break;
      case OP_XHAT:
        kp.func = reinterpret_cast<void*>(xhat_kernel);
        kp.gridDim = dim3((o.xhat.n + 255) / 256);
        xn = o.xhat.n;
        args[0] = &o.xhat.code; args[1] = &o.xhat.f; args[2] = &o.xhat.alpha; args[3] = &o.xhat.xh; args[4] = &xn;
        break;
      case OP_LOSS_D:
        kp.func = reinterpret_cast<void*>(loss_d_kernel); kp.blockDim = dim3(1024); args[0] = &o.loss;
        break;
      case OP_LOSS_G:
        kp.func = reinterpret_cast<void*>(loss_g_kernel); kp.blockDim = dim3(1024); args[0] = &o.loss;
        break;
      case OP_GIN:
        kp.func = reinterpret_cast<void*>(g_input_grads_kernel); kp.gridDim = dim3(B); args[0] = &o.gin;
        break;
      case OP_LGRAD:
        kp.func = reinterpret_cast<void*>(subspace_l_grad_kernel); kp.gridDim = dim3(kGLayers * kSub); args[0] = &o.lgrad;
        break;
    }
    kp.kernelParams = args;
    e = cudaGraphAddKernelNode(&node[i], gr, deps.data(), deps.size(), &kp);
  }
  if (e == cudaSuccess) e = cudaGraphInstantiate(&t->exec[which], gr, 0);
  if (gr) cudaGraphDestroy(gr);
  t->launches[which] = (int)ops.size() + 1;
  t->n_ops[which] = (int)ops.size();
  t->n_sync[which] = edges;
  return e;
}

cudaError_t build_persistent(chb_cttrain* t, int which) {
  if (!t->persist_grid) {
    int dev = 0, sms = 0, per_sm = 0, coop = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ctt_persistent_kernel, 256, 0);
    if (e != cudaSuccess) return e;
    if (!coop || per_sm < 1 || sms < 1) return cudaErrorNotSupported;
    t->persist_grid = sms;   // one CTA per SM: every CTA is resident, the barrier cannot deadlock
    e = cudaMalloc(&t->dev_bar, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(t->dev_bar, 0, sizeof(unsigned long long));
    if (e != cudaSuccess) return e;
  }
  std::vector<Op> ops;
  Rec rc{nullptr};
  rc.ops = &ops;
  record_step(t, which, rc);
  schedule_ops(ops, t->persist_grid, t->B, &t->n_sync[which]);
  cudaError_t e = cudaMalloc(&t->dev_ops[which], ops.size() * sizeof(Op));
  if (e == cudaSuccess) e = cudaMemcpy(t->dev_ops[which], ops.data(), ops.size() * sizeof(Op), cudaMemcpyHostToDevice);
  t->n_ops[which] = (int)ops.size();
  t->launches[which] = 2;   // stage kernel + the persistent kernel
  return e;
}

}  // namespace

extern "C" {

int chb_cttrain_create(const chb_cttrain_config* cfg, chb_cttrain** out) {
  if (!cfg || !out || cfg->batch < 1 || cfg->batch > kMaxBatch) {
    chb::set_error("chb_cttrain_create: need 1 <= batch <= 1024");
    return CHB_ERR_ARG;
  }
  if (int rc = chb_check_device()) return rc;
  chb_cttrain* t = new chb_cttrain();
  t->cfg = *cfg;
  t->B = cfg->batch;
  int64_t cur = 0;
  build_mlp(t, t->D, "D.", CHB_CTT_D, kCode, kHid, 4, kDOut, cur);
  t->nD = cur;
  cur = 0;
  t->g_win = add_tensor(t, "G.main_layer_in.weight", (int64_t)kHid * kGIn, CHB_CTT_G, cur);
  t->g_bin = add_tensor(t, "G.main_layer_in.bias", kHid, CHB_CTT_G, cur);
  for (int l = 0; l < kGLayers; ++l) {
    const int od = l < kGLayers - 1 ? kHid : kCode;
    const std::string b = "G.main_layer_mid." + std::to_string(l) + ".1.";
    t->g_w[l] = add_tensor(t, b + "weight", (int64_t)od * kHid, CHB_CTT_G, cur);
    t->g_b[l] = add_tensor(t, b + "bias", od, CHB_CTT_G, cur);
  }
  for (int l = 0; l < kGLayers; ++l) {
    const std::string b = "G.subspaces." + std::to_string(l) + ".";
    t->g_U[l] = add_tensor(t, b + "U", (int64_t)kSub * kHid, CHB_CTT_G, cur);
    t->g_L[l] = add_tensor(t, b + "L", kSub, CHB_CTT_G, cur);
    t->g_mu[l] = add_tensor(t, b + "mu", kHid, CHB_CTT_G, cur);
  }
  t->nG = cur;
  cur = 0;
  t->off_P = 0;
  build_mlp(t, t->P, "P.", CHB_CTT_FROZEN, kCode, kHid, 3, kPOut, cur);
  t->off_C = 0;  // offsets of C's tensors are absolute inside the frozen region, like P's
  build_mlp(t, t->C, "C.", CHB_CTT_FROZEN, kCode, kCHid, 3, 1, cur);
  t->nF = cur;
  // workspace arena
  const int64_t B = t->B;
  int64_t off = 0;
  auto take = [&](int64_t floats) { const int64_t o = off; off += (floats + 63) / 64 * 64; return o; };
  struct Slot { float** p; int64_t n; };
  std::vector<Slot> slots;
  auto F = [&](float*& p, int64_t n) { slots.push_back({&p, n}); };
  F(t->code, B * kCode); F(t->rgb, B * 3); F(t->pca, B); F(t->noise, B * kNoise); F(t->nc, B); F(t->label, B);
  F(t->alpha, B); F(t->ones, B); F(t->e0, B * kDOut);
  for (int l = 0; l < 4; ++l) { F(t->zR[l], B * kHid); F(t->zF[l], B * kHid); F(t->zH[l], B * kHid); F(t->uH[l], B * kHid);
                                F(t->dz[l], B * kHid); F(t->dt[l], B * kHid); }
  for (int l = 0; l < kGLayers; ++l) { F(t->LzA[l], B * kSub); F(t->LzG[l], B * kSub); F(t->sA[l], B * kHid);
                                       F(t->sG[l], B * kHid); F(t->dsA[l], B * kHid); F(t->dsG[l], B * kHid); }
  for (int l = 0; l < 3; ++l) { F(t->zP[l], B * kHid); F(t->zC[l], B * kCHid); F(t->dzc[l], B * kCHid); }
  F(t->r, B * kDOut); F(t->q, B * kDOut); F(t->xinA, B * kXinLd); F(t->xinG, B * kXinLd); F(t->zG, B * kNoise);
  F(t->labG, B); F(t->ae, B * kCode); F(t->f, B * kCode); F(t->xh, B * kCode); F(t->g, B * kCode);
  F(t->pout, B * kPOut); F(t->cout, B); F(t->dR, B * kDOut); F(t->dQ, B * kDOut); F(t->dAE, B * kCode);
  F(t->Gt, B * kCode); F(t->df, B * kCode); F(t->dP, B * kPOut); F(t->dC, B); F(t->losses, 64);
  std::vector<int64_t> offs;
  for (auto& s : slots) offs.push_back(take(s.n));
  const int64_t ints = take(4 * B + 64);
  t->ws_bytes = off * (int64_t)sizeof(float);
  // stash offsets as fake pointers until bind
  for (size_t i = 0; i < slots.size(); ++i) *slots[i].p = reinterpret_cast<float*>(offs[i] * sizeof(float));
  t->p1 = reinterpret_cast<int*>(ints * sizeof(float));
  *out = t;
  return CHB_OK;
}

void chb_cttrain_destroy(chb_cttrain* t) {
  if (!t) return;
  for (int i = 0; i < 2; ++i) {
    if (t->exec[i]) cudaGraphExecDestroy(t->exec[i]);
    if (t->dev_ops[i]) cudaFree(t->dev_ops[i]);
  }
  if (t->dev_bar) cudaFree(t->dev_bar);
  delete t;
}

int chb_cttrain_num_tensors(const chb_cttrain* t) { return t ? (int)t->table.size() : 0; }

int chb_cttrain_tensor_info(const chb_cttrain* t, int i, char* name, int name_cap, int64_t* offset, int64_t* numel,
                            int* group) {
  if (!t || i < 0 || i >= (int)t->table.size() || !name || name_cap < 1) {
    chb::set_error("chb_cttrain_tensor_info: index out of range");
    return CHB_ERR_ARG;
  }
  const Entry& e = t->table[i];
  std::strncpy(name, e.name.c_str(), name_cap - 1);
  name[name_cap - 1] = 0;
  if (offset) *offset = e.offset;
  if (numel) *numel = e.numel;
  if (group) *group = e.group;
  return CHB_OK;
}

int64_t chb_cttrain_state_floats(const chb_cttrain* t) { return t ? 4 * (t->nD + t->nG) + t->nF : 0; }

int chb_cttrain_region(const chb_cttrain* t, int region, int group, int64_t* offset, int64_t* numel) {
  if (!t || !offset || !numel || group < 0 || group > 2 || (group != CHB_CTT_FROZEN && (region < 0 || region > 3))) {
    chb::set_error("chb_cttrain_region: bad arguments");
    return CHB_ERR_ARG;
  }
  if (group == CHB_CTT_FROZEN) {
    *offset = 4 * (t->nD + t->nG);
    *numel = t->nF;
  } else {
    *offset = region * (t->nD + t->nG) + (group == CHB_CTT_D ? 0 : t->nD);
    *numel = group == CHB_CTT_D ? t->nD : t->nG;
  }
  return CHB_OK;
}

int64_t chb_cttrain_workspace_bytes(const chb_cttrain* t) { return t ? t->ws_bytes : 0; }

int chb_cttrain_bind(chb_cttrain* t, float* state, void* workspace) {
  if (!t || !state || !workspace || (reinterpret_cast<uintptr_t>(workspace) & 255) || t->ws) {
    chb::set_error("chb_cttrain_bind: NULL / misaligned (256 B) pointers, or already bound");
    return CHB_ERR_ARG;
  }
  t->state = state;
  t->ws = static_cast<char*>(workspace);
  auto fix = [&](float*& p) { p = reinterpret_cast<float*>(t->ws + reinterpret_cast<uintptr_t>(p)); };
  fix(t->code); fix(t->rgb); fix(t->pca); fix(t->noise); fix(t->nc); fix(t->label); fix(t->alpha); fix(t->ones); fix(t->e0);
  for (int l = 0; l < 4; ++l) { fix(t->zR[l]); fix(t->zF[l]); fix(t->zH[l]); fix(t->uH[l]); fix(t->dz[l]); fix(t->dt[l]); }
  for (int l = 0; l < kGLayers; ++l) { fix(t->LzA[l]); fix(t->LzG[l]); fix(t->sA[l]); fix(t->sG[l]); fix(t->dsA[l]); fix(t->dsG[l]); }
  for (int l = 0; l < 3; ++l) { fix(t->zP[l]); fix(t->zC[l]); fix(t->dzc[l]); }
  fix(t->r); fix(t->q); fix(t->xinA); fix(t->xinG); fix(t->zG); fix(t->labG); fix(t->ae); fix(t->f); fix(t->xh); fix(t->g);
  fix(t->pout); fix(t->cout); fix(t->dR); fix(t->dQ); fix(t->dAE); fix(t->Gt); fix(t->df); fix(t->dP); fix(t->dC);
  fix(t->losses);
  t->p1 = reinterpret_cast<int*>(t->ws + reinterpret_cast<uintptr_t>(t->p1));
  t->p2 = t->p1 + t->B;
  t->p3 = t->p2 + t->B;
  t->flag = t->p3 + t->B;
  return CHB_OK;
}

int chb_cttrain_step(chb_cttrain* t, int which, const chb_cttrain_batch* b, float* losses_out, void* stream) {
  if (!t || !t->ws || !b || !losses_out || (which != CHB_CTT_D && which != CHB_CTT_G)) {
    chb::set_error("chb_cttrain_step: NULL argument, unbound trainer or bad sub-step id");
    return CHB_ERR_ARG;
  }
  if (!b->code || !b->rgb_mean || !b->pca_std || !b->noise || !b->noise_curliness || !b->curliness_label ||
      !b->perm_rgb || !b->perm_curliness || !b->perm_noise || (which == CHB_CTT_D && !b->alpha_gp)) {
    chb::set_error("chb_cttrain_step: NULL batch field");
    return CHB_ERR_ARG;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  StageArgs sa{};
  sa.in = *b;
  sa.code = t->code; sa.rgb = t->rgb; sa.pca = t->pca; sa.noise = t->noise; sa.nc = t->nc; sa.label = t->label;
  sa.alpha = t->alpha; sa.ones = t->ones; sa.e0 = t->e0; sa.p1 = t->p1; sa.p2 = t->p2; sa.p3 = t->p3; sa.flag = t->flag;
  sa.B = t->B;
  stage_kernel<<<std::min(64, (t->B * kCode + 255) / 256), 256, 0, s>>>(sa);
  cudaError_t err = cudaGetLastError();
  const bool graph = t->cfg.use_graph == 1 && s != nullptr && s != cudaStreamLegacy;
  if (err == cudaSuccess && t->cfg.use_graph == 2) {
    if (!t->dev_ops[which]) err = build_persistent(t, which);
    if (err == cudaSuccess) {
      const Op* ops = static_cast<const Op*>(t->dev_ops[which]);
      int nops = t->n_ops[which];
      unsigned long long* bar = t->dev_bar;
      unsigned long long base = t->bar_total;
      void* args[4] = {&ops, &nops, &bar, &base};
      err = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(ctt_persistent_kernel), dim3(t->persist_grid), dim3(256),
                                        args, 0, s);
      t->bar_total += (unsigned long long)t->n_sync[which] * (unsigned long long)t->persist_grid;
    }
  } else if (err == cudaSuccess && graph) {
    if (!t->exec[which]) err = build_graph(t, which);
    if (err == cudaSuccess) err = cudaGraphLaunch(t->exec[which], s);
  } else if (err == cudaSuccess) {
    Rec rc{s};
    record_step(t, which, rc);
    err = rc.err;
    t->launches[which] = rc.n + 1;
  }
  if (err == cudaSuccess)
    err = cudaMemcpyAsync(losses_out, t->losses, CHB_CTT_NUM_LOSSES * sizeof(float), cudaMemcpyDeviceToDevice, s);
  if (err != cudaSuccess) {
    chb::set_error(std::string("chb_cttrain_step failed: ") + cudaGetErrorString(err));
    return CHB_ERR_CUDA;
  }
  return CHB_OK;
}

int chb_cttrain_adam(chb_cttrain* t, int which, void* stream) {
  if (!t || !t->ws || (which != CHB_CTT_D && which != CHB_CTT_G)) {
    chb::set_error("chb_cttrain_adam: unbound trainer or bad net id");
    return CHB_ERR_ARG;
  }
  const long step = ++t->step_count[which];
  const double bc1 = 1.0 - std::pow((double)t->cfg.beta1, (double)step);
  const double bc2 = 1.0 - std::pow((double)t->cfg.beta2, (double)step);
  const int64_t n = which == CHB_CTT_D ? t->nD : t->nG;
  adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      t->par(which), t->grad(which), t->adam_m(which), t->adam_v(which), n, t->cfg.beta1, t->cfg.beta2, t->cfg.eps,
      (float)(t->cfg.lr / bc1), (float)std::sqrt(bc2));
  const cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    chb::set_error(std::string("chb_cttrain_adam launch failed: ") + cudaGetErrorString(err));
    return CHB_ERR_CUDA;
  }
  return CHB_OK;
}

int chb_cttrain_launches(const chb_cttrain* t, int which) {
  return (t && (which == 0 || which == 1)) ? t->launches[which] : 0;
}

/* Operations and dependency edges (graph) / grid barriers (persistent kernel) of one sub-step (0 before first use). */
int chb_cttrain_schedule(const chb_cttrain* t, int which, int* n_ops, int* n_barriers) {
  if (!t || (which != 0 && which != 1)) return CHB_ERR_ARG;
  if (n_ops) *n_ops = t->n_ops[which];
  if (n_barriers) *n_barriers = t->n_sync[which];
  return CHB_OK;
}

}  // extern "C"
