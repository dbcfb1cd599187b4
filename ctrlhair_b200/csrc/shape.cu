// Shape branch generator (mask encoders / decoders): shape_branch/model.py:69-199, layers from
// my_torchlib/module.py (Conv2dBlock :67-137, custom LayerNorm :177-205, LinearBlock :16-64).
//
// Every convolution and both fully-connected layers run on the tcgen05 implicit-GEMM kernel:
//   * encoder conv4x4 stride 2 pad 1  ==  conv3x3 stride 1 pad 1 over the space-to-depth map [H/2, W/2, 4C]
//     (input row 2y-1+ky lives in block row y-1 / y / y / y+1 with parity 1 / 0 / 1 / 0); the packer scatters the 16
//     taps into the 9 x 4 (tap, parity) slots and leaves the impossible combinations zero
//   * decoder [nearest up2, conv3x3]: the LayerNorm-apply kernel writes the upsampled fp16 map the conv reads
//   * fc layers are 1x1 "convs" over a [1, B] image; the packer permutes their rows/columns between the reference's
//     NCHW flatten order and the NHWC order used here
// LayerNorm (per-sample mean / UNBIASED std over C*H*W, (x-mean)/(std+eps)*gamma_c+beta_c) + LeakyReLU(0.2) is a
// statistics kernel (double accumulation) plus an apply kernel that also produces the next conv's layout.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/ctrlhair_b200.h"
#include "conv_igemm.cuh"

namespace chb {

constexpr int kPosCh = 40;  // 4 * pos_encoding_order (shape_branch/config.py:20, model.py:18-30)

// positional embedding table fp16 [S][S][40]: channel k < 20: sin(2^(k/2) pi c), else cos; even k -> x, odd k -> y
// (np.meshgrid(c, c) stacks [x-grid, y-grid]; gamma1 = sin(nums * bi) reshaped to [-1, S, S] -> channel = 2*o + axis)
__global__ void shape_pos_kernel(__half* pos, int S) {
  const long long total = (long long)S * S * kPosCh;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % kPosCh);
    const int x = (int)((i / kPosCh) % S);
    const int y = (int)(i / ((long long)kPosCh * S));
    const int kk = k % 20;
    const int order = kk / 2, axis = kk % 2;
    const double c = (double)(axis == 0 ? x : y) / (double)S;
    const double arg = ldexp(3.14159265358979323846, order) * c;
    const float v = (float)(k < 20 ? sin(arg) : cos(arg));
    pos[i] = __float2half_rn(v);
  }
}

// Encoder input: concat [mask | positional map] (model.py:96-101) in the space-to-depth layout the first conv reads:
// out[b, by, bx, par * Cin + c] = in[c][2 by + par / 2][2 bx + par % 2], channels >= 4 Cin zero.  One block per
// (32 output pixels of a row, row, image): the gather runs with the pixel index across the lanes (coalesced reads of
// the NCHW mask planes), a padded shared tile turns it into contiguous 16-byte stores.  HBM-bound: 4 B read per mask
// element + 2 B written per output element.
constexpr int kPrepBx = 32;
// With `labels` (uint8 [B,S,S], 255 = none) instead of `mask`, the one-hot channels are synthesised on the fly
// (shape_util.py:6-26: channel c of the hair net is class 13, of the face net class c + (c >= 13)), so a caller that
// holds a label map needs neither the [B,19,S,S] one-hot tensor nor the hair / face split.
__global__ void __launch_bounds__(256) shape_prep_kernel(const float* __restrict__ mask,
                                                         const uint8_t* __restrict__ labels,
                                                         const __half* __restrict__ pos, __half* __restrict__ out, int B,
                                                         int Cm, int S, int Cpad) {
  extern __shared__ __align__(16) unsigned char prep_smem[];
  __half* tile = reinterpret_cast<__half*>(prep_smem);   // [kPrepBx][Cpad + 2]
  const int pitch = Cpad + 2;
  const int Cin = Cm + kPosCh, Hh = S / 2;
  const int bx0 = blockIdx.x * kPrepBx, by = blockIdx.y, b = blockIdx.z;
  // mask planes: pixel index across the lanes (8-byte stride in x), one 4-byte load per element
  for (int idx = threadIdx.x; idx < 4 * Cm * kPrepBx; idx += 256) {
    const int chm = idx / kPrepBx, bxl = idx - chm * kPrepBx;
    const int par = chm / Cm, c = chm - par * Cm;
    const int bx = bx0 + bxl;
    __half v = __float2half_rn(0.f);
    if (bx < Hh) {
      const int y = 2 * by + (par >> 1), x = 2 * bx + (par & 1);
      if (labels) {
        const int cls = Cm == 1 ? 13 : (c < 13 ? c : c + 1);
        v = __float2half_rn(labels[((long long)b * S + y) * S + x] == cls ? 1.f : 0.f);
      } else {
        v = __float2half_rn(__ldg(mask + (((long long)b * Cm + c) * S + y) * S + x));
      }
    }
    tile[bxl * pitch + par * Cin + c] = v;
  }
  // positional map: 40 contiguous halves per source pixel = five 16-byte loads
  static_assert(kPosCh % 8 == 0, "positional channels are read in 16-byte pieces");
  constexpr int kQ = kPosCh / 8;
  for (int idx = threadIdx.x; idx < 4 * kPrepBx * kQ; idx += 256) {
    const int q = idx % kQ, t = idx / kQ;
    const int bxl = t % kPrepBx, par = t / kPrepBx;
    const int bx = bx0 + bxl;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (bx < Hh) {
      const int y = 2 * by + (par >> 1), x = 2 * bx + (par & 1);
      v = __ldg(reinterpret_cast<const uint4*>(pos + ((long long)y * S + x) * kPosCh) + q);
    }
    const __half* h = reinterpret_cast<const __half*>(&v);
    __half* dst = tile + bxl * pitch + par * Cin + Cm + q * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[j] = h[j];
  }
  // zero padding channels
  const int npad = Cpad - 4 * Cin;
  for (int idx = threadIdx.x; idx < npad * kPrepBx; idx += 256) {
    const int bxl = idx / npad, k = idx - bxl * npad;
    tile[bxl * pitch + 4 * Cin + k] = __float2half_rn(0.f);
  }
  __syncthreads();
  const int c8 = Cpad / 8;
  __half* orow = out + (((long long)b * Hh + by) * Hh + bx0) * Cpad;
  for (int idx = threadIdx.x; idx < kPrepBx * c8; idx += 256) {
    const int bxl = idx / c8, k = idx - bxl * c8;
    if (bx0 + bxl >= Hh) continue;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(tile + bxl * pitch + k * 8);   // pitch is even: 4-byte aligned
    *reinterpret_cast<uint4*>(orow + (long long)bxl * Cpad + k * 8) = make_uint4(src[0], src[1], src[2], src[3]);
  }
}

// Sum and sum of squares of each image (the custom LayerNorm of my_torchlib/module.py:177-205 normalises over C*H*W).
// 16-byte loads, four in flight per thread, fp32 partials flushed to double every 64 values, one atomic per warp.
// The last block of an image to finish turns the moments into {1 / (std + eps), -mean / (std + eps)} (unbiased std, eps
// added to the std: module.py:189-199), so no separate finalize launch is needed.
__global__ void __launch_bounds__(256) ln_stats_kernel(const float* __restrict__ x, double* __restrict__ sums,
                                                       unsigned* __restrict__ done, float2* __restrict__ ss,
                                                       long long n) {
  const int b = blockIdx.y;
  const float4* xb = reinterpret_cast<const float4*>(x + (long long)b * n);
  const long long n4 = n >> 2;   // n = H*W*C is a multiple of 8
  double d1 = 0.0, d2 = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += 4 * stride) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      v[u] = (i + u * stride < n4) ? __ldg(xb + i + u * stride) : make_float4(0.f, 0.f, 0.f, 0.f);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      s1 += (v[u].x + v[u].y) + (v[u].z + v[u].w);
      s2 = fmaf(v[u].x, v[u].x, s2); s2 = fmaf(v[u].y, v[u].y, s2);
      s2 = fmaf(v[u].z, v[u].z, s2); s2 = fmaf(v[u].w, v[u].w, s2);
    }
    d1 += s1; d2 += s2;
  }
  for (int o = 16; o > 0; o >>= 1) {
    d1 += __shfl_xor_sync(0xffffffffu, d1, o);
    d2 += __shfl_xor_sync(0xffffffffu, d2, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&sums[b * 2], d1);
    atomicAdd(&sums[b * 2 + 1], d2);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(&done[b], 1u);
    if (t == gridDim.x - 1) {   // every other block's atomics are visible: it fenced before it counted
      __threadfence();
      const double m1 = atomicAdd(&sums[b * 2], 0.0), m2 = atomicAdd(&sums[b * 2 + 1], 0.0);
      const double dn = (double)n;
      const double mean = m1 / dn;
      const double var = (m2 - dn * mean * mean) / (dn - 1.0);
      const double inv = 1.0 / (sqrt(var > 0.0 ? var : 0.0) + 1e-5);
      ss[b] = make_float2((float)inv, (float)(-mean * inv));
    }
  }
}

// LayerNorm apply (+ per-channel affine + LeakyReLU 0.2) and relayout to the fp16 tensor the next conv reads.
// mode 0: same layout; mode 1: space-to-depth (next conv is 4x4 stride 2); mode 2: nearest 2x upsample.
// norm 0: plain fp32 -> fp16 conversion.  One thread per 8 channels of a pixel: 16-byte loads and stores.
__global__ void __launch_bounds__(256) ln_apply_kernel(const float* __restrict__ x, const float2* __restrict__ ss,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       __half* __restrict__ out, int B, int H, int W, int C, int mode,
                                                       int norm) {
  const int c8 = C >> 3;
  const unsigned per_img = (unsigned)H * W * c8;
  const long long total = (long long)B * per_img;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / per_img);
    unsigned p = (unsigned)(i - (long long)b * per_img);
    const int c = (int)(p % c8) * 8;
    p /= c8;
    const int xx = (int)(p % W), yy = (int)(p / W);
    const float4* src = reinterpret_cast<const float4*>(x + i * 8);
    const float4 a = __ldg(src), bq = __ldg(src + 1);
    float v[8] = {a.x, a.y, a.z, a.w, bq.x, bq.y, bq.z, bq.w};
    if (norm) {
      const float2 s = __ldg(ss + b);
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float be[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float t = fmaf(fmaf(v[k], s.x, s.y), g[k], be[k]);
        v[k] = t > 0.f ? t : 0.2f * t;
      }
    }
    uint32_t pk[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      __half2 h = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
      pk[k] = *reinterpret_cast<uint32_t*>(&h);
    }
    const uint4 q = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    if (mode == 0) {
      *reinterpret_cast<uint4*>(out + i * 8) = q;
    } else if (mode == 1) {
      const int par = (yy & 1) * 2 + (xx & 1);
      *reinterpret_cast<uint4*>(out + (((long long)b * (H / 2) + (yy >> 1)) * (W / 2) + (xx >> 1)) * (4LL * C) +
                                (long long)par * C + c) = q;
    } else {
      __half* o = out + (((long long)b * 2 * H + 2 * yy) * 2 * W + 2 * xx) * C + c;
      *reinterpret_cast<uint4*>(o) = q;
      *reinterpret_cast<uint4*>(o + C) = q;
      *reinterpret_cast<uint4*>(o + 2LL * W * C) = q;
      *reinterpret_cast<uint4*>(o + 2LL * W * C + C) = q;
    }
  }
}

// code fp32 [B, K] -> fp16 [B, Kpad] (zero padded): decoder fc input
__global__ void code_pad_kernel(const float* __restrict__ a, int Ka, const float* __restrict__ b2, int Kb,
                                __half* __restrict__ out, int B, int Kpad) {
  const long long total = (long long)B * Kpad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % Kpad);
    const int b = (int)(i / Kpad);
    float v = 0.f;
    if (k < Ka) v = a[(long long)b * Ka + k];
    else if (k < Ka + Kb) v = b2[(long long)b * Kb + (k - Ka)];
    out[i] = __float2half_rn(v);
  }
}

// mask = softmax([face[:13], hair, face[13:]])  (model.py:184-187) -> fp32 NCHW [B,19,S,S]
// `labels` (optional): mask_one_hot_to_label of the same probabilities (shape_util.py:17-20: first maximum; a softmax
// row is never all zero, so 255 cannot occur) — written next to, or instead of, the [B,19,S,S] probabilities.
__global__ void shape_softmax_kernel(const float* __restrict__ hair /*[B,S,S,32]*/, const float* __restrict__ face
                                     /*[B,S,S,32]*/, float* __restrict__ out, uint8_t* __restrict__ labels, int B,
                                     int S) {
  const long long total = (long long)B * S * S;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    float l[19];
    const float* f = face + i * 32;
#pragma unroll
    for (int k = 0; k < 13; ++k) l[k] = f[k];
    l[13] = hair[i * 32];
#pragma unroll
    for (int k = 13; k < 18; ++k) l[k + 1] = f[k];
    float m = l[0];
#pragma unroll
    for (int k = 1; k < 19; ++k) m = fmaxf(m, l[k]);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 19; ++k) {
      l[k] = expf(l[k] - m);
      s += l[k];
    }
    const float inv = 1.f / s;
    const long long b = i / ((long long)S * S), pix = i % ((long long)S * S);
    if (out) {
#pragma unroll
      for (int k = 0; k < 19; ++k) out[(b * 19 + k) * (long long)S * S + pix] = l[k] * inv;
    }
    if (labels) {   // argmax of the fp32 probabilities the reference would see (ties of the rounded products included)
      float best = l[0] * inv;
      int arg = 0;
#pragma unroll
      for (int k = 1; k < 19; ++k) {
        const float v = l[k] * inv;
        if (v > best) { best = v; arg = k; }
      }
      labels[i] = (uint8_t)arg;
    }
  }
}

// decoder logits fp32 NHWC [B,S,S,ld] (first C channels valid) -> fp32 NCHW [B,C,S,S]  (what MaskDecoder returns)
__global__ void logits_nchw_kernel(const float* __restrict__ in, int ld, int C, float* __restrict__ out, int B, int S) {
  const long long total = (long long)B * S * S;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / ((long long)S * S), pix = i % ((long long)S * S);
    for (int k = 0; k < C; ++k) out[(b * C + k) * (long long)S * S + pix] = in[i * ld + k];
  }
}

// forward_decoder on caller-provided logits (model.py:184-187): hair fp32 [B,1,S,S], face fp32 [B,18,S,S] (NCHW)
__global__ void softmax_nchw_kernel(const float* __restrict__ hair, const float* __restrict__ face,
                                    float* __restrict__ out, int B, int S) {
  const long long plane = (long long)S * S, total = (long long)B * plane;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / plane, pix = i % plane;
    float l[19];
#pragma unroll
    for (int k = 0; k < 13; ++k) l[k] = face[(b * 18 + k) * plane + pix];
    l[13] = hair[b * plane + pix];
#pragma unroll
    for (int k = 13; k < 18; ++k) l[k + 1] = face[(b * 18 + k) * plane + pix];
    float m = l[0];
#pragma unroll
    for (int k = 1; k < 19; ++k) m = fmaxf(m, l[k]);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 19; ++k) {
      l[k] = expf(l[k] - m);
      s += l[k];
    }
    const float inv = 1.f / s;
#pragma unroll
    for (int k = 0; k < 19; ++k) out[(b * 19 + k) * plane + pix] = l[k] * inv;
  }
}

static int sgrid(long long n, int threads) {
  long long g = (n + threads - 1) / threads;
  const long long cap = (long long)device_sm_count() * 16;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

struct STensor {
  std::string name;
  int64_t offset, nbytes;
  int dtype;
};

struct NetLayout {
  int w[8], b[8], g[8], be[8];  // conv weights / bias / LN gamma / beta per layer (enc: 0..6; dec: 0..6 + out conv at 7)
  int fcw, fcb;
};

}  // namespace chb

using namespace chb;

struct chb_shape {
  chb_shape_config cfg;
  std::vector<STensor> tensors;
  int64_t blob_bytes = 0, ws_bytes = 0;
  NetLayout enc[2], dec[2];  // 0 = hair, 1 = face
  int64_t ws_pos, ws_in, ws_conv, ws_act, ws_feat, ws_fcout, ws_code16, ws_logit[2], ws_sums, ws_io, ws_ks;
  const uint8_t* blob = nullptr;
  uint8_t* ws = nullptr;
  bool pos_ready = false;
  std::map<int, std::vector<ConvPlan>> enc_plans[2], dec_plans[2];
};

namespace chb {
static const int kEncCm[2] = {1, 18};
static const int kEncPad0[2] = {192, 256};   // 4*(Cm+40) rounded up to a multiple of 64
static const int kEncOut[2] = {32, 1024};    // hair: mean(16) ++ std(16); face: 1024
static const int kDecIn[2] = {1088, 1024};   // hair decoder input 1024+16 padded to a multiple of 64
static const int kDecOutRows[2] = {32, 32};  // out conv rows (1 / 18 valid): 32 puts both on the fast epilogue geometry
                                             // (the 16-row hair layer on the generic path took 0.29 ms, the face one 0.13)

static int sadd(chb_shape* z, const std::string& name, int64_t nbytes, int dtype) {
  STensor t{name, z->blob_bytes, nbytes, dtype};
  z->tensors.push_back(t);
  z->blob_bytes += (nbytes + 255) / 256 * 256;
  return (int)z->tensors.size() - 1;
}
static int64_t sws(chb_shape* z, int64_t n) {
  const int64_t o = z->ws_bytes;
  z->ws_bytes += (n + 1023) / 1024 * 1024;
  return o;
}
static int enc_cout(int i) { return 32 << i > 2048 ? 2048 : 32 << i; }
static int dec_cout(int i) { return 32 << (6 - i) > 2048 ? 2048 : 32 << (6 - i); }

// Few output tiles and a long K (the 2x2 .. 8x8 layers stream 9-151 MB of weights, the fully connected layers 17 MB):
// pick the N tile and a split-K factor together.  A wide N tile keeps the re-reads of the activation tile low (every
// n-tile CTA reads all of A), split-K (<= 8 CTAs per output tile, >= 2 channel chunks each) supplies the CTAs; the N
// tile only narrows when both together stay under ~96 CTAs.
static void few_tiles_plan(chb_conv_desc* d, int m_tiles, int nchunk, void* ks_ws) {
  const int sms = device_sm_count();
  const int steps = d->seg[0].taps * nchunk;   // K steps of 64 a CTA walks per tile (~0.2 us each)
  int bn = d->BN;
  for (;; bn /= 2) {
    const int tiles = m_tiles * (d->Nrows / bn);
    int s = sms / tiles;
    if (s > 8) s = 8;
    if (s > nchunk / 2) s = nchunk / 2;
    if (steps < 96) s = 1;                       // the combine costs ~7 us: short K does not split ...
    else if (s > steps / 24) s = steps / 24;     // ... and a split keeps at least 24 K steps
    if (s < 1) s = 1;
    const bool narrower = bn > 32 && d->Nrows % (bn / 2) == 0;
    if (tiles * s >= 96 || !narrower) {
      d->BN = bn;
      if (s >= 2 && tiles <= 1024) { d->ksplit = s; d->ks_ws = ks_ws; }
      return;
    }
  }
}

static chb_conv_desc conv_desc(int B, int H, int W, const void* a, int C, const void* w, const float* bias, int N, int BN,
                               void* out, void* ks_ws) {
  chb_conv_desc d;
  memset(&d, 0, sizeof d);
  d.B = B; d.H = H; d.W = W;
  d.TW = W < 8 ? W : 8;
  d.TH = H < 16 ? H : 16;
  d.TB = 1;
#ifdef CHB_TUNING_ENV
  static const bool batch_tiles = [] { const char* v = getenv("CHB_SHAPE_TB"); return !v || atoi(v) != 0; }();
#else
  const bool batch_tiles = true;
#endif
  if (batch_tiles && d.TW * d.TH <= 64) {
    // tiny maps (2x2 .. 8x8): several images share one 128-row MMA tile, so the 2048 x 36864 weight matrices of the
    // deep layers are streamed once per N tile instead of once per image ...
    d.TB = 128 / (d.TW * d.TH);
    if (d.TB > B) d.TB = B;
  }
  d.nseg = 1;
  chb_conv_seg& s = d.seg[0];
  s.a = a; s.Ca = C; s.C = C; s.taps = 9; s.w = w;
  s.a_sx = C; s.a_sy = (int64_t)W * C; s.a_sb = (int64_t)H * W * C;
  d.N = d.Nrows = N; d.BN = BN;
  d.epi = CHB_EPI_PLAIN; d.act = CHB_ACT_NONE; d.bias = bias;
  d.out = out; d.out_dtype = CHB_F32;
  d.o_sn = 1; d.o_sx = N; d.o_sy = (int64_t)W * N; d.o_sb = (int64_t)H * W * N;
  {
    const int m_tiles = ((B + d.TB - 1) / d.TB) * ((H + d.TH - 1) / d.TH) * ((W + d.TW - 1) / d.TW);
    if (m_tiles * (N / BN) < 96) few_tiles_plan(&d, m_tiles, C / (C == 32 ? 32 : 64), ks_ws);
  }
  return d;
}
static chb_conv_desc fc_desc(int B, const void* a, int K, const void* w, const float* bias, int N, int BN, void* out,
                             void* ks_ws) {
  chb_conv_desc d;
  memset(&d, 0, sizeof d);
  d.B = 1; d.H = 1; d.W = B;
  d.TW = B >= 128 ? 128 : (B + 7) / 8 * 8; d.TH = 1; d.TB = 1;
  d.nseg = 1;
  chb_conv_seg& s = d.seg[0];
  s.a = a; s.Ca = K; s.C = K; s.taps = 1; s.w = w;
  s.a_sx = K; s.a_sy = (int64_t)B * K; s.a_sb = (int64_t)B * K;
  d.N = d.Nrows = N; d.BN = BN;
  d.epi = CHB_EPI_PLAIN; d.act = CHB_ACT_NONE; d.bias = bias;
  d.out = out; d.out_dtype = CHB_F32;
  d.o_sn = 1; d.o_sx = N; d.o_sy = 0; d.o_sb = 0;
  {
    const int m_tiles = (B + d.TW - 1) / d.TW;
    if (m_tiles * (N / BN) < 96) few_tiles_plan(&d, m_tiles, K / 64, ks_ws);
  }
  return d;
}
static const void* bp(const chb_shape* z, int t) { return z->blob + z->tensors[t].offset; }
static const float* bpf(const chb_shape* z, int t) { return reinterpret_cast<const float*>(z->blob + z->tensors[t].offset); }

static int build_enc_plans(chb_shape* z, int net, int B, std::vector<ConvPlan>& plans) {
  const int S = z->cfg.crop;
  plans.resize(8);
  int cin = kEncPad0[net];
  for (int i = 0; i < 7; ++i) {
    const int Hh = (S / 2) >> i, co = enc_cout(i);
    // layer i reads the space-to-depth map [B,Hh,Hh,cin] and writes fp32 [B,Hh,Hh,co]
    chb_conv_desc d = conv_desc(B, Hh, Hh, z->ws + (i == 0 ? z->ws_in : z->ws_act), cin, bp(z, z->enc[net].w[i]),
                                bpf(z, z->enc[net].b[i]), co, co < 256 ? co : 256, z->ws + z->ws_conv, z->ws + z->ws_ks);
    int rc = build_conv_plan(d, &plans[i]);
    if (rc != CHB_OK) return rc;
    cin = 4 * co;
  }
  chb_conv_desc f = fc_desc(B, z->ws + z->ws_feat, 8192, bp(z, z->enc[net].fcw), bpf(z, z->enc[net].fcb), kEncOut[net],
                            kEncOut[net] < 256 ? kEncOut[net] : 256, z->ws + z->ws_fcout, z->ws + z->ws_ks);
  return build_conv_plan(f, &plans[7]);
}

static int build_dec_plans(chb_shape* z, int net, int B, std::vector<ConvPlan>& plans) {
  const int S = z->cfg.crop;
  plans.resize(9);
  chb_conv_desc f = fc_desc(B, z->ws + z->ws_code16, kDecIn[net], bp(z, z->dec[net].fcw), bpf(z, z->dec[net].fcb), 8192,
                            256, z->ws + z->ws_fcout, z->ws + z->ws_ks);
  int rc = build_conv_plan(f, &plans[0]);
  if (rc != CHB_OK) return rc;
  int cin = 2048;
  for (int i = 0; i < 7; ++i) {
    const int r = 4 << i, co = dec_cout(i);
    chb_conv_desc d = conv_desc(B, r, r, z->ws + z->ws_act, cin, bp(z, z->dec[net].w[i]), bpf(z, z->dec[net].b[i]), co,
                                co < 256 ? co : 256, z->ws + z->ws_conv, z->ws + z->ws_ks);
    if ((rc = build_conv_plan(d, &plans[1 + i])) != CHB_OK) return rc;
    cin = co;
  }
  chb_conv_desc o = conv_desc(B, S, S, z->ws + z->ws_act, 32, bp(z, z->dec[net].w[7]), bpf(z, z->dec[net].b[7]),
                              kDecOutRows[net], kDecOutRows[net], z->ws + z->ws_logit[net], z->ws + z->ws_ks);
  return build_conv_plan(o, &plans[8]);
}

static void ln(chb_shape* z, cudaStream_t st, int B, int H, int W, int C, const float* gamma, const float* beta,
               __half* out, int mode, int norm) {
  const float* x = reinterpret_cast<const float*>(z->ws + (norm == 2 ? z->ws_fcout : z->ws_conv));
  double* sums = reinterpret_cast<double*>(z->ws + z->ws_sums);
  const long long n = (long long)H * W * C;
  const size_t mb = (size_t)z->cfg.max_batch;
  unsigned* done = reinterpret_cast<unsigned*>(sums + mb * 2);             // [max_batch] block counters (8 B slots)
  float2* ss = reinterpret_cast<float2*>(sums + mb * 3);
  if (norm == 1) {
    cudaMemsetAsync(sums, 0, mb * 3 * sizeof(double), st);                 // moments and counters
    long long blocks = (n / 4 + 256 * 16 - 1) / (256 * 16);
    const long long want = ((long long)device_sm_count() * 8 + B - 1) / B;
    if (blocks > want) blocks = want;
    if (blocks < 1) blocks = 1;
    ln_stats_kernel<<<dim3((unsigned)blocks, (unsigned)B), 256, 0, st>>>(x, sums, done, ss, n);
  }
  ln_apply_kernel<<<sgrid((long long)B * (n / 8), 256), 256, 0, st>>>(x, ss, gamma, beta, out, B, H, W, C, mode,
                                                                      norm == 1 ? 1 : 0);
}
}  // namespace chb

extern "C" {

int chb_shape_create(const chb_shape_config* cfg, chb_shape** out) {
  if (!cfg || !out || cfg->crop != 256 || cfg->max_batch <= 0) {
    set_error("chb_shape_create: need crop == 256 (the reference's fixed mask size) and max_batch > 0");
    return CHB_ERR_ARG;
  }
  chb_shape* z = new chb_shape();
  z->cfg = *cfg;
  const char* nm[2] = {"hair", "face"};
  for (int net = 0; net < 2; ++net) {
    int cin = kEncPad0[net];
    const std::string p = std::string(nm[net]) + "_encoder.";
    for (int i = 0; i < 7; ++i) {
      const int co = enc_cout(i);
      z->enc[net].w[i] = sadd(z, p + std::to_string(i) + ".w", (int64_t)co * 9 * cin * 2, CHB_F16);
      z->enc[net].b[i] = sadd(z, p + std::to_string(i) + ".b", (int64_t)co * 4, CHB_F32);
      z->enc[net].g[i] = sadd(z, p + std::to_string(i) + ".gamma", (int64_t)co * 4, CHB_F32);
      z->enc[net].be[i] = sadd(z, p + std::to_string(i) + ".beta", (int64_t)co * 4, CHB_F32);
      cin = 4 * co;
    }
    z->enc[net].fcw = sadd(z, p + "fc.w", (int64_t)kEncOut[net] * 8192 * 2, CHB_F16);
    z->enc[net].fcb = sadd(z, p + "fc.b", (int64_t)kEncOut[net] * 4, CHB_F32);
  }
  for (int net = 0; net < 2; ++net) {
    const std::string p = std::string(nm[net]) + "_decoder.";
    z->dec[net].fcw = sadd(z, p + "fc.w", (int64_t)8192 * kDecIn[net] * 2, CHB_F16);
    z->dec[net].fcb = sadd(z, p + "fc.b", (int64_t)8192 * 4, CHB_F32);
    int cin = 2048;
    for (int i = 0; i < 7; ++i) {
      const int co = dec_cout(i);
      z->dec[net].w[i] = sadd(z, p + std::to_string(i) + ".w", (int64_t)co * 9 * cin * 2, CHB_F16);
      z->dec[net].b[i] = sadd(z, p + std::to_string(i) + ".b", (int64_t)co * 4, CHB_F32);
      z->dec[net].g[i] = sadd(z, p + std::to_string(i) + ".gamma", (int64_t)co * 4, CHB_F32);
      z->dec[net].be[i] = sadd(z, p + std::to_string(i) + ".beta", (int64_t)co * 4, CHB_F32);
      cin = co;
    }
    z->dec[net].w[7] = sadd(z, p + "out.w", (int64_t)kDecOutRows[net] * 9 * 32 * 2, CHB_F16);
    z->dec[net].b[7] = sadd(z, p + "out.b", (int64_t)kDecOutRows[net] * 4, CHB_F32);
  }
  const int64_t B = cfg->max_batch, S = cfg->crop;
  z->ws_pos = sws(z, S * S * kPosCh * 2);
  z->ws_in = sws(z, B * (S / 2) * (S / 2) * 256 * 2);
  z->ws_conv = sws(z, B * S * S * 32 * 4);          // largest fp32 conv output: 256x256x32 (== 128x128x32x... all smaller)
  z->ws_act = sws(z, B * S * S * 32 * 2 * 2);       // largest fp16 activation: 256x256x32 plain / 128x128x64 upsampled
  z->ws_feat = sws(z, B * 8192 * 2);
  z->ws_fcout = sws(z, B * 8192 * 4);
  z->ws_code16 = sws(z, B * 1088 * 2);
  z->ws_logit[0] = sws(z, B * S * S * 32 * 4);
  z->ws_logit[1] = sws(z, B * S * S * 32 * 4);
  z->ws_sums = sws(z, B * 2 * 8 + B * 8 + B * 8);   // double moments [B][2], block counters [B], float2 constants [B]
  z->ws_io = sws(z, B * 19 * S * S * 4);
  z->ws_ks = sws(z, (int64_t)ksplit_workspace_bytes(device_sm_count()));   // split-K counters + partial accumulators
  *out = z;
  return CHB_OK;
}

void chb_shape_destroy(chb_shape* z) { delete z; }
int chb_shape_num_tensors(const chb_shape* z) { return z ? (int)z->tensors.size() : 0; }
int chb_shape_tensor_info(const chb_shape* z, int i, char* name, int cap, int64_t* offset, int64_t* nbytes, int* dtype) {
  if (!z || i < 0 || i >= (int)z->tensors.size()) {
    set_error("chb_shape_tensor_info: index out of range");
    return CHB_ERR_ARG;
  }
  const STensor& t = z->tensors[i];
  if (name && cap > 0) snprintf(name, cap, "%s", t.name.c_str());
  if (offset) *offset = t.offset;
  if (nbytes) *nbytes = t.nbytes;
  if (dtype) *dtype = t.dtype;
  return CHB_OK;
}
int64_t chb_shape_blob_bytes(const chb_shape* z) { return z ? z->blob_bytes : 0; }
int64_t chb_shape_workspace_bytes(const chb_shape* z) { return z ? z->ws_bytes : 0; }

int chb_shape_bind(chb_shape* z, const void* blob, void* workspace) {
  if (!z || !blob || !workspace || (reinterpret_cast<uintptr_t>(blob) & 255) ||
      (reinterpret_cast<uintptr_t>(workspace) & 1023)) {
    set_error("chb_shape_bind: NULL or misaligned pointers (blob 256 B, workspace 1024 B)");
    return CHB_ERR_ARG;
  }
  int rc = chb_check_device();
  if (rc != CHB_OK) return rc;
  z->blob = reinterpret_cast<const uint8_t*>(blob);
  z->ws = reinterpret_cast<uint8_t*>(workspace);
  for (int n = 0; n < 2; ++n) {
    z->enc_plans[n].clear();
    z->dec_plans[n].clear();
  }
  shape_pos_kernel<<<sgrid((long long)z->cfg.crop * z->cfg.crop * kPosCh, 256), 256>>>(
      reinterpret_cast<__half*>(z->ws + z->ws_pos), z->cfg.crop);
  cudaError_t err = cudaMemset(z->ws + z->ws_ks, 0, kKsCounterBytes);   // split-K tile counters start at zero
  if (err == cudaSuccess) err = cudaDeviceSynchronize();
  if (err != cudaSuccess) {
    set_error(std::string("chb_shape_bind: positional table kernel failed: ") + cudaGetErrorString(err));
    return CHB_ERR_CUDA;
  }
  return CHB_OK;
}

// net: 0 = hair encoder (mask [B,1,S,S] -> out [B,32] = mean(16) ++ |std|(16)), 1 = face encoder ([B,18,S,S] -> [B,1024])
static int shape_encode_impl(chb_shape* z, int net, const float* mask, const uint8_t* labels, float* out, int B,
                             void* stream_);

int chb_shape_encode(chb_shape* z, int net, const float* mask, float* out, int B, void* stream_) {
  if (!mask) {
    chb::set_error("chb_shape_encode: bad arguments or unbound object");
    return CHB_ERR_ARG;
  }
  return shape_encode_impl(z, net, mask, nullptr, out, B, stream_);
}

int chb_shape_encode_labels(chb_shape* z, int net, const uint8_t* labels, float* out, int B, void* stream_) {
  if (!labels) {
    chb::set_error("chb_shape_encode_labels: bad arguments or unbound object");
    return CHB_ERR_ARG;
  }
  return shape_encode_impl(z, net, nullptr, labels, out, B, stream_);
}

static int shape_encode_impl(chb_shape* z, int net, const float* mask, const uint8_t* labels, float* out, int B,
                             void* stream_) {
  if (!z || !out || !z->ws || net < 0 || net > 1 || B <= 0 || B > z->cfg.max_batch) {
    set_error("chb_shape_encode: bad arguments or unbound object");
    return CHB_ERR_ARG;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  auto it = z->enc_plans[net].find(B);
  if (it == z->enc_plans[net].end()) {
    std::vector<ConvPlan> pl;
    int rc = build_enc_plans(z, net, B, pl);
    if (rc != CHB_OK) return rc;
    it = z->enc_plans[net].emplace(B, std::move(pl)).first;
  }
  const std::vector<ConvPlan>& pl = it->second;
  const int S = z->cfg.crop;
  shape_prep_kernel<<<dim3((unsigned)((S / 2 + kPrepBx - 1) / kPrepBx), (unsigned)(S / 2), (unsigned)B), 256,
                      (size_t)kPrepBx * (kEncPad0[net] + 2) * sizeof(__half), st>>>(
      mask, labels, reinterpret_cast<const __half*>(z->ws + z->ws_pos), reinterpret_cast<__half*>(z->ws + z->ws_in), B,
      kEncCm[net], S, kEncPad0[net]);
  for (int i = 0; i < 7; ++i) {
    int rc = launch_conv_plan(pl[i], CHB_IMPL_TCGEN05, st);
    if (rc != CHB_OK) return rc;
    const int Hh = (S / 2) >> i, co = enc_cout(i);
    // layers 0..5 feed another 4x4/s2 conv (space-to-depth); layer 6 feeds the fc (plain NHWC flatten)
    ln(z, st, B, Hh, Hh, co, bpf(z, z->enc[net].g[i]), bpf(z, z->enc[net].be[i]),
       reinterpret_cast<__half*>(z->ws + (i < 6 ? z->ws_act : z->ws_feat)), i < 6 ? 1 : 0, 1);
  }
  int rc = launch_conv_plan(pl[7], CHB_IMPL_TCGEN05, st);
  if (rc != CHB_OK) return rc;
  cudaError_t err = cudaMemcpyAsync(out, z->ws + z->ws_fcout, (size_t)B * kEncOut[net] * 4, cudaMemcpyDeviceToDevice, st);
  if (err != cudaSuccess) {
    set_error(std::string("chb_shape_encode: ") + cudaGetErrorString(err));
    return CHB_ERR_CUDA;
  }
  return CHB_OK;
}

// One MaskDecoder (model.py:116-143): net 0 = hair decoder on cat([face_code, hair_code]) (model.py:175-178),
// net 1 = face decoder on face_code (:180-182).  Leaves the logits fp32 NHWC in ws_logit[net].
static int run_decoder(chb_shape* z, int net, const float* hair_code, const float* face_code, int B, cudaStream_t st) {
  auto it = z->dec_plans[net].find(B);
  if (it == z->dec_plans[net].end()) {
    std::vector<ConvPlan> pl;
    int rc = build_dec_plans(z, net, B, pl);
    if (rc != CHB_OK) return rc;
    it = z->dec_plans[net].emplace(B, std::move(pl)).first;
  }
  const std::vector<ConvPlan>& pl = it->second;
  code_pad_kernel<<<sgrid((long long)B * kDecIn[net], 256), 256, 0, st>>>(
      face_code, 1024, hair_code, net == 0 ? 16 : 0, reinterpret_cast<__half*>(z->ws + z->ws_code16), B, kDecIn[net]);
  int rc = launch_conv_plan(pl[0], CHB_IMPL_TCGEN05, st);
  if (rc != CHB_OK) return rc;
  // fc output [B,8192] is already NHWC [B,2,2,2048] (rows permuted by the packer): cast + upsample to 4x4
  ln(z, st, B, 2, 2, 2048, nullptr, nullptr, reinterpret_cast<__half*>(z->ws + z->ws_act), 2, 2);
  for (int i = 0; i < 7; ++i) {
    if ((rc = launch_conv_plan(pl[1 + i], CHB_IMPL_TCGEN05, st)) != CHB_OK) return rc;
    const int r = 4 << i, co = dec_cout(i);
    ln(z, st, B, r, r, co, bpf(z, z->dec[net].g[i]), bpf(z, z->dec[net].be[i]),
       reinterpret_cast<__half*>(z->ws + z->ws_act), i < 6 ? 2 : 0, 1);
  }
  return launch_conv_plan(pl[8], CHB_IMPL_TCGEN05, st);
}

static int last_launch(const char* who) {
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    set_error(std::string(who) + ": " + cudaGetErrorString(err));
    return CHB_ERR_CUDA;
  }
  return CHB_OK;
}

// hair_code fp32 [B,16], face_code fp32 [B,1024] -> mask probabilities fp32 [B,19,S,S]  (model.py:195-199)
int chb_shape_decode(chb_shape* z, const float* hair_code, const float* face_code, float* mask_out, int B,
                     void* stream_) {
  if (!z || !hair_code || !face_code || !mask_out || !z->ws || B <= 0 || B > z->cfg.max_batch) {
    set_error("chb_shape_decode: bad arguments or unbound object");
    return CHB_ERR_ARG;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  const int S = z->cfg.crop;
  for (int net = 0; net < 2; ++net) {
    int rc = run_decoder(z, net, hair_code, face_code, B, st);
    if (rc != CHB_OK) return rc;
  }
  shape_softmax_kernel<<<sgrid((long long)B * S * S, 256), 256, 0, st>>>(
      reinterpret_cast<const float*>(z->ws + z->ws_logit[0]), reinterpret_cast<const float*>(z->ws + z->ws_logit[1]),
      mask_out, nullptr, B, S);
  return last_launch("chb_shape_decode");
}

// forward_decode_by_code + mask_one_hot_to_label (ui/backend.py:89-90,312-313) in one call: uint8 labels [B,S,S].
int chb_shape_decode_labels(chb_shape* z, const float* hair_code, const float* face_code, uint8_t* labels_out, int B,
                            void* stream_) {
  if (!z || !hair_code || !face_code || !labels_out || !z->ws || B <= 0 || B > z->cfg.max_batch) {
    set_error("chb_shape_decode_labels: bad arguments or unbound object");
    return CHB_ERR_ARG;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  const int S = z->cfg.crop;
  for (int net = 0; net < 2; ++net) {
    int rc = run_decoder(z, net, hair_code, face_code, B, st);
    if (rc != CHB_OK) return rc;
  }
  shape_softmax_kernel<<<sgrid((long long)B * S * S, 256), 256, 0, st>>>(
      reinterpret_cast<const float*>(z->ws + z->ws_logit[0]), reinterpret_cast<const float*>(z->ws + z->ws_logit[1]),
      nullptr, labels_out, B, S);
  return last_launch("chb_shape_decode_labels");
}

// forward_hair_decoder (net 0, model.py:175-178) / forward_face_decoder (net 1, :180-182): logits fp32 NCHW,
// [B,1,S,S] for the hair decoder, [B,18,S,S] for the face decoder.  hair_code is ignored for net 1 (may be NULL).
int chb_shape_decode_logits(chb_shape* z, int net, const float* hair_code, const float* face_code, float* logits_out,
                            int B, void* stream_) {
  if (!z || (net != 0 && net != 1) || (net == 0 && !hair_code) || !face_code || !logits_out || !z->ws || B <= 0 ||
      B > z->cfg.max_batch) {
    set_error("chb_shape_decode_logits: bad arguments or unbound object");
    return CHB_ERR_ARG;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  const int S = z->cfg.crop;
  int rc = run_decoder(z, net, hair_code, face_code, B, st);
  if (rc != CHB_OK) return rc;
  logits_nchw_kernel<<<sgrid((long long)B * S * S, 256), 256, 0, st>>>(
      reinterpret_cast<const float*>(z->ws + z->ws_logit[net]), 32, net == 0 ? 1 : 18, logits_out, B, S);
  return last_launch("chb_shape_decode_logits");
}

// forward_decoder (model.py:184-187): softmax over [face[:13], hair, face[13:]]; inputs and output fp32 NCHW.
int chb_shape_softmax(chb_shape* z, const float* hair_logit, const float* face_logit, float* mask_out, int B,
                      void* stream_) {
  if (!z || !hair_logit || !face_logit || !mask_out || B <= 0) {
    set_error("chb_shape_softmax: bad arguments");
    return CHB_ERR_ARG;
  }
  int rc = chb_check_device();
  if (rc != CHB_OK) return rc;
  const int S = z->cfg.crop;
  softmax_nchw_kernel<<<sgrid((long long)B * S * S, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      hair_logit, face_logit, mask_out, B, S);
  return last_launch("chb_shape_softmax");
}

}  // extern "C"
