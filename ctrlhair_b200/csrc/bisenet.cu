// BiSeNet (ResNet-18) face parsing on the GPU: external_code/face_parsing/model.py:230-274 (BiSeNet.forward, first head),
// :105-146 (ContextPath), :81-103 (AttentionRefinementModule), :196-231 (FeatureFusionModule), :37-56 (BiSeNetOutput),
// resnet.py:21-93 (BasicBlock / Resnet18) and the caller my_parsing_util.py:31-54 (normalise, argmax, label swap) +
// hair_editor.py:331-335 (nearest resize of the label map).  The reference runs this network on the CPU at 512x512 for
// every input and target image (`.cuda()` commented out, my_parsing_util.py:37,41).
//
// Every conv but the 7x7 stem runs on the tcgen05 implicit-GEMM operator (conv_igemm.cu): fp16 NHWC activations, eval
// BatchNorm folded into weights + bias by the host packer, ReLU in the epilogue.
//   * identity shortcuts ride in the conv's GEMM as a 1x1 K-segment with identity weights (exact: 1.0 * x in fp16 with
//     fp32 accumulation), learned 1x1/s2 shortcuts as a 1x1 K-segment over the subsampled block input
//   * stride-2 3x3 convs are computed at the input resolution and subsampled (3 small layers; 4x their FLOPs is ~25 % of
//     the network's 21 GFLOP per image)
//   * the FFM's concat is two K-segments
// Small HBM-bound helpers around it (this file): stem 7x7/s2 (+ input normalisation, SIMT), 3x3/s2 max-pool, 2x
// subsample, global average pool, per-image dense layers (attention vectors), scale/add/nearest-upsample, and the tail
// (bilinear x8 align_corners + argmax + label swap + nearest resize) which never materialises the 19 x 512 x 512 logits.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/ctrlhair_b200.h"
#include "conv_igemm.cuh"

namespace chb {

constexpr int kPoolSlabs = 16;
static inline unsigned bgrid(long long n, int block) {
  long long g = (n + block - 1) / block;
  const long long cap = (long long)device_sm_count() * 32;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

// ---------------------------------------------------------------- stem: normalise + conv 7x7 s2 p3 (3 -> 64) + BN + ReLU
// img u8 [B,S,S,3] RGB (what cv2 / PIL hand over after the 512x512 resize); x = (img/255 - mean) / std
// (my_parsing_util.py:25-28), evaluated in the reference's operation order (scale, subtract, divide).  One thread = one
// output pixel x all 64 output channels: the 7 x 7 x 3 patch is fetched and normalised once, every weight read is a
// shared-memory broadcast (a warp reads one 16-byte word for four FMAs; rows of 21 weights are padded to 24).
// A CTA of 128 threads takes a 16 x 8 tile of output pixels: warps 0-1 compute channels 0..31 of them, warps 2-3
// channels 32..63 (a warp shares one channel half, so every weight read is a broadcast), and a thread owns TWO pixels
// (rows ty and ty + 8) so that each 16-byte weight word feeds eight FMAs.  The 37 x 21 x 3 input patch
// of the tile is normalised ONCE into shared memory (each pixel on its own would redo 147 fp32 divisions and byte loads:
// a third of the instructions), and the FMAs go out as packed pairs (fma.rn.f32x2, two output channels per issue slot;
// the weights sit [ky][k][co] so that a 16-byte word holds four output channels of one tap).  Each output is still the
// same chain of 147 fp32 FMAs in the same order — bit for bit the result of the first version of this kernel (one thread
// per pixel x 64 channels, 1.87 ms at B = 32), which is what keeps the label agreement of the parser where it was.
__device__ __forceinline__ unsigned long long stem_pk2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ unsigned long long stem_fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
constexpr int kStemTile = 8, kStemIn = 2 * kStemTile + 5, kStemPitch = 68;   // 21 input columns, rows of 63 (+5) floats
constexpr int kStemTileH = 16, kStemInH = 2 * kStemTileH + 5;                 // 37 input rows
__global__ void __launch_bounds__(128) bisenet_stem_kernel(const uint8_t* __restrict__ img,
                                                           const float* __restrict__ w /*[64][148]: co, (ky,kx,ci)*/,
                                                           const float* __restrict__ bias, __half* __restrict__ out,
                                                           int B, int S) {
  __shared__ __align__(16) float sw[7 * 21 * 64];            // [ky][k = kx*3+ci][co]
  __shared__ __align__(16) float sin_[kStemInH * kStemPitch];  // normalised input patch of the tile
  __shared__ float sb[64];
  for (int i = threadIdx.x; i < 7 * 21 * 64; i += blockDim.x) {
    const int co = i & 63, r = i >> 6, ky = r / 21, k = r - ky * 21;
    sw[i] = w[co * 148 + ky * 21 + k];
  }
  if (threadIdx.x < 64) sb[threadIdx.x] = bias[threadIdx.x];
  const int So = S / 2, tiles = So / kStemTile, tiles_y = So / kStemTileH;
  const float nm[3] = {0.485f, 0.456f, 0.406f}, ns[3] = {0.229f, 0.224f, 0.225f};
  const int half = threadIdx.x >> 6;                 // channel half of this warp
  const int lp = threadIdx.x & 63, ty = lp >> 3, tx = lp & 7;
  for (long long t = blockIdx.x; t < (long long)B * tiles * tiles_y; t += gridDim.x) {
    const int tX = (int)(t % tiles), tY = (int)((t / tiles) % tiles_y), b = (int)(t / ((long long)tiles * tiles_y));
    const int y0 = tY * kStemTileH, x0 = tX * kStemTile;
    __syncthreads();   // the previous tile's patch is no longer read (first pass: sw / sb are complete)
    for (int i = threadIdx.x; i < kStemInH * kStemIn * 3; i += blockDim.x) {
      const int c = i % 3, px = (i / 3) % kStemIn, py = i / (3 * kStemIn);
      const int yy = 2 * y0 - 3 + py, xx = 2 * x0 - 3 + px;
      const bool ok = yy >= 0 && yy < S && xx >= 0 && xx < S;
      // zero padding applies to the NORMALISED image; the reference's operation order: scale, subtract, divide
      sin_[py * kStemPitch + px * 3 + c] =
          ok ? ((float)__ldg(img + (((long long)b * S + yy) * S + xx) * 3 + c) * (1.f / 255.f) - nm[c]) / ns[c] : 0.f;
    }
    __syncthreads();
    unsigned long long acc[2][16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[0][j] = acc[1][j] = stem_pk2(sb[32 * half + 2 * j], sb[32 * half + 2 * j + 1]);
#pragma unroll 1
    for (int ky = 0; ky < 7; ++ky) {
      // 21 contiguous values per pixel: (kx, ci); the second pixel sits 8 output rows = 16 input rows further down
      const float* irow = sin_ + (2 * ty + ky) * kStemPitch + 6 * tx;
      const float4* wk = reinterpret_cast<const float4*>(sw + ky * 21 * 64 + 32 * half);
#pragma unroll
      for (int k = 0; k < 21; ++k) {
        const float v0 = irow[k], v1 = irow[16 * kStemPitch + k];
        const unsigned long long vv0 = stem_pk2(v0, v0), vv1 = stem_pk2(v1, v1);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 wv = wk[k * 16 + q];
          const unsigned long long w01 = stem_pk2(wv.x, wv.y), w23 = stem_pk2(wv.z, wv.w);
          acc[0][2 * q] = stem_fma2(vv0, w01, acc[0][2 * q]);
          acc[0][2 * q + 1] = stem_fma2(vv0, w23, acc[0][2 * q + 1]);
          acc[1][2 * q] = stem_fma2(vv1, w01, acc[1][2 * q]);
          acc[1][2 * q + 1] = stem_fma2(vv1, w23, acc[1][2 * q + 1]);
        }
      }
    }
#pragma unroll
    for (int pp = 0; pp < 2; ++pp) {
      const long long pix = ((long long)b * So + (y0 + ty + 8 * pp)) * So + (x0 + tx);
      uint4* o = reinterpret_cast<uint4*>(out + pix * 64 + 32 * half);
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        __half2 h[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float lo, hi;
          asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[pp][4 * v + j]));
          h[j] = __floats2half2_rn(fmaxf(lo, 0.f), fmaxf(hi, 0.f));
        }
        o[v] = *reinterpret_cast<uint4*>(h);
      }
    }
  }
}

// ---------------------------------------------------------------- max-pool 3x3 s2 p1, fp16 NHWC (resnet.py:63,75)
__global__ void __launch_bounds__(256) maxpool3s2_kernel(const __half* __restrict__ in, __half* __restrict__ out, int B,
                                                         int Hin, int C) {
  const int Ho = Hin / 2, c8 = C / 8;
  const long long total = (long long)B * Ho * Ho * c8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % c8);
    long long pix = i / c8;
    const int x = (int)(pix % Ho), y = (int)((pix / Ho) % Ho), b = (int)(pix / ((long long)Ho * Ho));
    __half2 m[4];
    bool first = true;
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = 2 * y + ky - 1;
      if (yy < 0 || yy >= Hin) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = 2 * x + kx - 1;
        if (xx < 0 || xx >= Hin) continue;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + (((long long)b * Hin + yy) * Hin + xx) * C) + cg);
        const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int k = 0; k < 4; ++k) m[k] = first ? h[k] : __hmax2(m[k], h[k]);
        first = false;
      }
    }
    *(reinterpret_cast<uint4*>(out + pix * C) + cg) = *reinterpret_cast<uint4*>(m);
  }
}

// ---------------------------------------------------------------- out[b,y,x,:] = in[b,2y,2x,:]  (stride-2 phase pick)
__global__ void __launch_bounds__(256) subsample2_kernel(const __half* __restrict__ in, __half* __restrict__ out, int B,
                                                         int Hin, int C) {
  const int Ho = Hin / 2, c8 = C / 8;
  const long long total = (long long)B * Ho * Ho * c8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % c8);
    long long pix = i / c8;
    const int x = (int)(pix % Ho), y = (int)((pix / Ho) % Ho), b = (int)(pix / ((long long)Ho * Ho));
    *(reinterpret_cast<uint4*>(out + pix * C) + cg) =
        __ldg(reinterpret_cast<const uint4*>(in + (((long long)b * Hin + 2 * y) * Hin + 2 * x) * C) + cg);
  }
}

// ---------------------------------------------------------------- global average pool: fp16 NHWC [B,HW,C] -> fp32 [B,S,C]
// grid = (B, S slabs of pixels), block = 256; thread t owns channel pair (t % (C/2)) and strides over its slab's pixels;
// fp32 partials via shared memory.  out[b][s][:] holds slab s's share of the mean (already divided by HW); the consumer
// (vec_dense_kernel) adds the S partial vectors, so the pool of a 64 x 64 x 256 map is read by 16 CTAs per image.
__global__ void __launch_bounds__(256) avgpool_kernel(const __half* __restrict__ in, float* __restrict__ out, int HW,
                                                      int C) {
  __shared__ float2 part[256];
  const int b = blockIdx.x, c2 = C / 2, lanes = 256 / c2;
  const int slab = blockIdx.y, nslab = gridDim.y;
  const int per = (HW + nslab - 1) / nslab, q0 = slab * per, q1 = min(HW, q0 + per);
  const int cp = threadIdx.x % c2, pl = threadIdx.x / c2;
  float2 acc = make_float2(0.f, 0.f);
  if (pl < lanes) {
    const __half2* p = reinterpret_cast<const __half2*>(in + (long long)b * HW * C) + cp;
    for (int q = q0 + pl; q < q1; q += lanes) {
      const float2 v = __half22float2(p[(long long)q * c2]);
      acc.x += v.x; acc.y += v.y;
    }
  }
  part[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < c2) {
    float2 s = make_float2(0.f, 0.f);
    for (int l = 0; l < lanes; ++l) { s.x += part[l * c2 + threadIdx.x].x; s.y += part[l * c2 + threadIdx.x].y; }
    float* o = out + ((long long)b * nslab + slab) * C;
    o[2 * threadIdx.x] = s.x / (float)HW;
    o[2 * threadIdx.x + 1] = s.y / (float)HW;
  }
}

// ---------------------------------------------------------------- per-image dense layer y = act(W x + bias), fp32
// (the 1x1 convs on pooled vectors: conv_avg, conv_atten, ffm.conv1/conv2).  act: 0 none, 1 relu, 2 sigmoid.
__global__ void __launch_bounds__(256) vec_dense_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bias, float* __restrict__ y, int Cin,
                                                        int Cout, int act, int nparts) {
  extern __shared__ float sx[];
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < Cin; i += blockDim.x) {   // x[b] = sum of nparts partial vectors (avgpool slabs)
    float v = 0.f;
    for (int p = 0; p < nparts; ++p) v += x[((long long)b * nparts + p) * Cin + i];
    sx[i] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int o = warp; o < Cout; o += blockDim.x >> 5) {
    const float* wr = w + (long long)o * Cin;
    float a = 0.f;
    for (int i = lane; i < Cin; i += 32) a = fmaf(wr[i], sx[i], a);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s);
    if (lane == 0) {
      a += bias ? bias[o] : 0.f;
      if (act == 1) a = fmaxf(a, 0.f);
      if (act == 2) a = 1.f / (1.f + __expf(-a));
      y[(long long)b * Cout + o] = a;
    }
  }
}

// ---------------------------------------------------------------- out = up_u( in * (scale[b,c] + sbias) + addv[b,c] + addt )
// in / addt fp16 NHWC [B,H,H,C], out fp16 NHWC [B,uH,uH,C] (u = 1 or 2, nearest: F.interpolate of model.py:134,141).
__global__ void __launch_bounds__(256) scale_add_up_kernel(const __half* __restrict__ in, const float* __restrict__ scale,
                                                           float sbias, const float* __restrict__ addv,
                                                           const __half* __restrict__ addt, __half* __restrict__ out,
                                                           int B, int H, int C, int up) {
  const int Ho = H * up, c8 = C / 8;
  const long long total = (long long)B * Ho * Ho * c8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % c8);
    long long pix = i / c8;
    const int x = (int)(pix % Ho), y = (int)((pix / Ho) % Ho), b = (int)(pix / ((long long)Ho * Ho));
    const long long src = (((long long)b * H + y / up) * H + x / up) * C;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + src) + cg);
    uint4 t = make_uint4(0u, 0u, 0u, 0u);
    if (addt) t = __ldg(reinterpret_cast<const uint4*>(addt + src) + cg);
    const __half2* hv = reinterpret_cast<const __half2*>(&v);
    const __half2* ht = reinterpret_cast<const __half2*>(&t);
    __half2 o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = cg * 8 + 2 * k;
      const float2 f = __half22float2(hv[k]);
      const float2 a = __half22float2(ht[k]);
      const float s0 = scale[(long long)b * C + c] + sbias, s1 = scale[(long long)b * C + c + 1] + sbias;
      const float v0 = addv ? addv[(long long)b * C + c] : 0.f, v1 = addv ? addv[(long long)b * C + c + 1] : 0.f;
      o[k] = __floats2half2_rn(fmaf(f.x, s0, v0 + a.x), fmaf(f.y, s1, v1 + a.y));
    }
    *(reinterpret_cast<uint4*>(out + pix * C) + cg) = *reinterpret_cast<uint4*>(o);
  }
}

// ---------------------------------------------------------------- tail: bilinear (align_corners) + argmax + label swap
// logits fp32 NHWC [B,h,h,32] (19 valid) at 1/8 resolution.  For output pixel (Y,X) of an So x So label map that samples
// the S x S logit map at (Y*step, X*step) (step = S/So: cv2.resize INTER_NEAREST of hair_editor.py:334 takes source index
// floor(dst * S/So); step = 1 gives the full parsing map of my_parsing_util.py:45-46):
//   F.interpolate(out, (S,S), 'bilinear', align_corners=True)  (model.py:270): src = dst * (h-1)/(S-1)
//   argmax over the 19 classes, first maximum (numpy argmax), then lut[label]  (my_parsing_util.py:49-54).
__global__ void __launch_bounds__(256) logits_to_mask_kernel(const float* __restrict__ logits,
                                                             const uint8_t* __restrict__ lut, uint8_t* __restrict__ out,
                                                             int B, int h, int S, int So, int ncls) {
  const int step = S / So;
  const float sc = (float)(h - 1) / (float)(S - 1);
  const long long total = (long long)B * So * So;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(i % So), Y = (int)((i / So) % So), b = (int)(i / ((long long)So * So));
    // ATen upsample_bilinear2d (align_corners): real index = scale * dst; i0 = (int) idx; lambda = idx - i0
    const float fy = sc * (float)(Y * step), fx = sc * (float)(X * step);
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < h - 1 ? 1 : 0);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float hy = 1.f - ly, hx = 1.f - lx;
    const float* p00 = logits + (((long long)b * h + y0) * h + x0) * 32;
    const float* p01 = logits + (((long long)b * h + y0) * h + x1) * 32;
    const float* p10 = logits + (((long long)b * h + y1) * h + x0) * 32;
    const float* p11 = logits + (((long long)b * h + y1) * h + x1) * 32;
    float best = -3.4e38f;
    int arg = 0;
    for (int c = 0; c < ncls; ++c) {
      // same association as ATen: h0 * (w0 * v00 + w1 * v01) + h1 * (w0 * v10 + w1 * v11)
      const float v = hy * (hx * __ldg(p00 + c) + lx * __ldg(p01 + c)) + ly * (hx * __ldg(p10 + c) + lx * __ldg(p11 + c));
      if (v > best) { best = v; arg = c; }
    }
    out[i] = lut ? lut[arg] : (uint8_t)arg;
  }
}

// ---------------------------------------------------------------- Pillow's BILINEAR resize, 8 bits per channel, bit exact
// my_parsing_util.py:35 resizes every image to 512x512 with PIL before the network.  Pillow resamples separably, first
// along x then along y, with a triangle filter of support max(scale, 1), coefficients normalised in double precision and
// rounded to 22-bit fixed point, an int32 accumulator started at 1 << 21 and the intermediate image rounded to uint8
// between the passes (libImaging/Resample.c).  The host computes the coefficient tables with the same double arithmetic
// (ctrlhair_b200/bisenet.py: pil_bilinear_tables); the passes below are pure integer work.
//   one pass: out[.., o, ..] = clip8((2^21 + sum_{k < n[o]} in[.., lo[o] + k, ..] * coef[o][k]) >> 22)
__global__ void __launch_bounds__(256) pil_resample_pass_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                                const int* __restrict__ bounds /*[O][2]: lo, n*/,
                                                                const int* __restrict__ coef /*[O][ks]*/, int ks,
                                                                long long outer, int I, int O, int inner) {
  // tensor viewed as [outer][I][inner] -> [outer][O][inner]; x pass: outer = B*H, inner = 3; y pass: outer = B, inner = W*3
  const long long total = outer * O * inner;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % inner);
    const int o = (int)((i / inner) % O);
    const long long q = i / ((long long)inner * O);
    const int lo = __ldg(bounds + 2 * o), n = __ldg(bounds + 2 * o + 1);
    const uint8_t* p = in + (q * I + lo) * inner + c;
    int acc = 1 << 21;
    for (int k = 0; k < n; ++k) acc += (int)__ldg(p + (long long)k * inner) * __ldg(coef + o * ks + k);
    acc >>= 22;
    out[i] = (uint8_t)(acc < 0 ? 0 : (acc > 255 ? 255 : acc));
  }
}

// The same pass, four consecutive `inner` elements per thread (inner % 4 == 0 and 4-byte aligned rows: the y pass, whose
// inner extent is a whole image row): one 32-bit load per tap instead of four byte loads, one 32-bit store.
__global__ void __launch_bounds__(256) pil_resample_pass4_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                                 const int* __restrict__ bounds, const int* __restrict__ coef,
                                                                 int ks, long long outer, int I, int O, int inner) {
  const int inner4 = inner >> 2;
  const long long total = outer * O * inner4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % inner4);
    const int o = (int)((i / inner4) % O);
    const long long q = i / ((long long)inner4 * O);
    const int lo = __ldg(bounds + 2 * o), n = __ldg(bounds + 2 * o + 1);
    const uint8_t* p = in + (q * I + lo) * inner + 4 * c4;
    int a0 = 1 << 21, a1 = 1 << 21, a2 = 1 << 21, a3 = 1 << 21;
    for (int k = 0; k < n; ++k) {
      const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(p + (long long)k * inner));
      const int w = __ldg(coef + o * ks + k);
      a0 += (int)(v & 0xffu) * w;
      a1 += (int)((v >> 8) & 0xffu) * w;
      a2 += (int)((v >> 16) & 0xffu) * w;
      a3 += (int)(v >> 24) * w;
    }
    auto clip8 = [](int a) { a >>= 22; return (uint32_t)(a < 0 ? 0 : (a > 255 ? 255 : a)); };
    reinterpret_cast<uint32_t*>(out)[(q * O + o) * inner4 + c4] = clip8(a0) | (clip8(a1) << 8) | (clip8(a2) << 16) | (clip8(a3) << 24);
  }
}

struct BTensor {
  std::string name;
  int64_t offset, nbytes;
  int dtype;
};

struct BConv {   // one tcgen05 conv of the schedule
  std::string name;
  int H;            // output (= computed) resolution
  int nseg;
  struct Seg { int64_t a_off; int C; int taps; int t_w; } seg[3];
  int N, Nrows, t_b, act;
  int64_t out_off;
  int out_dtype;
};

}  // namespace chb

using namespace chb;

struct chb_bisenet {
  chb_bisenet_config cfg;
  std::vector<BTensor> tensors;
  int64_t blob_bytes = 0, ws_bytes = 0;
  const uint8_t* blob = nullptr;
  uint8_t* ws = nullptr;
  std::map<std::string, int> tid;
  std::map<std::string, int64_t> buf;       // workspace buffers by name
  std::vector<BConv> convs;
  std::map<std::string, int> cid;           // conv index by name
  std::map<int, std::vector<ConvPlan>> plans;
};

namespace chb {

static int badd(chb_bisenet* n, const std::string& name, int64_t nbytes, int dtype) {
  BTensor t;
  t.name = name; t.offset = n->blob_bytes; t.nbytes = nbytes; t.dtype = dtype;
  n->tensors.push_back(t);
  n->blob_bytes += (nbytes + 255) / 256 * 256;
  n->tid[name] = (int)n->tensors.size() - 1;
  return (int)n->tensors.size() - 1;
}
static int64_t bws(chb_bisenet* n, const std::string& name, int64_t nbytes) {
  const int64_t off = n->ws_bytes;
  n->ws_bytes += (nbytes + 1023) / 1024 * 1024;
  n->buf[name] = off;
  return off;
}

// conv with up to 3 K-segments; weights "<name>.w<k>" [Nrows][taps*C] fp16, bias "<name>.b" [Nrows] fp32 (BN folded)
static void add_conv(chb_bisenet* n, const std::string& name, int H, int N, int act, const std::string& out, int out_dtype,
                     std::initializer_list<std::tuple<std::string, int, int>> segs) {
  BConv c;
  c.name = name; c.H = H; c.N = N; c.Nrows = (N + 31) / 32 * 32; c.act = act; c.nseg = 0;
  for (auto& s : segs) {
    BConv::Seg& g = c.seg[c.nseg];
    g.a_off = n->buf.at(std::get<0>(s));
    g.C = std::get<1>(s);
    g.taps = std::get<2>(s);
    g.t_w = badd(n, name + ".w" + std::to_string(c.nseg), (int64_t)c.Nrows * g.taps * g.C * 2, CHB_F16);
    ++c.nseg;
  }
  c.t_b = badd(n, name + ".b", (int64_t)c.Nrows * 4, CHB_F32);
  c.out_off = n->buf.at(out);
  c.out_dtype = out_dtype;
  n->cid[name] = (int)n->convs.size();
  n->convs.push_back(c);
}

static void build_bisenet_layout(chb_bisenet* n) {
  const int64_t B = n->cfg.max_batch, S = n->cfg.size;
  const int R = (int)S / 4;
  auto act16 = [&](const std::string& nm, int64_t H, int64_t C) { bws(n, nm, B * H * H * C * 2); };
  bws(n, "img", B * S * S * 3);
  bws(n, "mask", B * S * S);
  act16("stem", S / 2, 64);
  act16("p0", R, 64);
  badd(n, "stem.w", 64 * 148 * 4, CHB_F32);
  badd(n, "stem.b", 64 * 4, CHB_F32);
  // ---- ResNet-18 layers (resnet.py:54-81).  Block input `x`, scratch `a` (conv1 output), output `y`.
  const int ch[4] = {64, 128, 256, 512};
  std::string x = "p0";
  int H = R, Cin = 64;
  for (int li = 0; li < 4; ++li) {
    const int C = ch[li];
    for (int bi = 0; bi < 2; ++bi) {
      const std::string p = "layer" + std::to_string(li + 1) + "." + std::to_string(bi);
      const bool down = li > 0 && bi == 0;
      if (down) {
        // conv1 3x3/s2 at the input resolution, then the even phase; the 1x1/s2 shortcut reads the subsampled input
        act16(p + ".full", H, C);
        add_conv(n, p + ".conv1", H, C, CHB_ACT_RELU, p + ".full", CHB_F16, {{x, Cin, 9}});
        act16(p + ".a", H / 2, C);
        act16(p + ".xs", H / 2, Cin);
        act16(p + ".y", H / 2, C);
        H /= 2;
        add_conv(n, p + ".conv2", H, C, CHB_ACT_RELU, p + ".y", CHB_F16, {{p + ".a", C, 9}, {p + ".xs", Cin, 1}});
      } else {
        act16(p + ".a", H, C);
        act16(p + ".y", H, C);
        add_conv(n, p + ".conv1", H, C, CHB_ACT_RELU, p + ".a", CHB_F16, {{x, C, 9}});
        add_conv(n, p + ".conv2", H, C, CHB_ACT_RELU, p + ".y", CHB_F16, {{p + ".a", C, 9}, {x, C, 1}});
      }
      x = p + ".y";
      Cin = C;
    }
  }
  const int H8 = R / 2, H16 = R / 4, H32 = R / 8;
  // ---- context path (model.py:127-146)
  bws(n, "v_avg32", B * 512 * 4 * kPoolSlabs);
  bws(n, "v_avg", B * 128 * 4);
  badd(n, "conv_avg.w", 128 * 512 * 4, CHB_F32);
  badd(n, "conv_avg.b", 128 * 4, CHB_F32);
  for (const char* arm : {"arm32", "arm16"}) {
    const bool is32 = std::string(arm) == "arm32";
    const int Hh = is32 ? H32 : H16, Ci = is32 ? 512 : 256;
    act16(std::string(arm) + ".feat", Hh, 128);
    add_conv(n, std::string(arm) + ".conv", Hh, 128, CHB_ACT_RELU, std::string(arm) + ".feat", CHB_F16,
             {{is32 ? "layer4.1.y" : "layer3.1.y", Ci, 9}});
    bws(n, std::string(arm) + ".pool", B * 128 * 4 * kPoolSlabs);
    bws(n, std::string(arm) + ".att", B * 128 * 4);
    badd(n, std::string(arm) + ".att.w", 128 * 128 * 4, CHB_F32);
    badd(n, std::string(arm) + ".att.b", 128 * 4, CHB_F32);
    act16(std::string(arm) + ".sum_up", 2 * Hh, 128);
    act16(std::string(arm) + ".head", 2 * Hh, 128);
    add_conv(n, is32 ? "conv_head32" : "conv_head16", 2 * Hh, 128, CHB_ACT_RELU, std::string(arm) + ".head", CHB_F16,
             {{std::string(arm) + ".sum_up", 128, 9}});
  }
  // ---- feature fusion (model.py:218-228): cat[feat_res8, feat_cp8] as two K-segments
  act16("ffm.feat", H8, 256);
  add_conv(n, "ffm.convblk", H8, 256, CHB_ACT_RELU, "ffm.feat", CHB_F16, {{"layer2.1.y", 128, 1}, {"arm16.head", 128, 1}});
  bws(n, "ffm.pool", B * 256 * 4 * kPoolSlabs);
  bws(n, "ffm.mid", B * 64 * 4);
  bws(n, "ffm.att", B * 256 * 4);
  badd(n, "ffm.conv1.w", 64 * 256 * 4, CHB_F32);
  badd(n, "ffm.conv2.w", 256 * 64 * 4, CHB_F32);
  act16("ffm.out", H8, 256);
  // ---- output head (model.py:47-51)
  act16("out.feat", H8, 256);
  add_conv(n, "conv_out.conv", H8, 256, CHB_ACT_RELU, "out.feat", CHB_F16, {{"ffm.out", 256, 9}});
  bws(n, "logits", B * H8 * H8 * 32 * 4);
  add_conv(n, "conv_out.conv_out", H8, n->cfg.n_classes, CHB_ACT_NONE, "logits", CHB_F32, {{"out.feat", 256, 1}});
  badd(n, "label_lut", 32, CHB_F32);   // 19 bytes used (uint8 LUT; dtype tag only sizes it)
}

static int bisenet_plans(chb_bisenet* n, int B, std::vector<ConvPlan>& plans) {
  plans.resize(n->convs.size());
  for (size_t i = 0; i < n->convs.size(); ++i) {
    const BConv& c = n->convs[i];
    chb_conv_desc d;
    memset(&d, 0, sizeof d);
    d.B = B; d.H = c.H; d.W = c.H;
    d.TW = c.H < 8 ? c.H : 8;
    d.TH = c.H < 16 ? c.H : 16;
    d.TB = 1;
    d.nseg = c.nseg;
    for (int s = 0; s < c.nseg; ++s) {
      chb_conv_seg& g = d.seg[s];
      g.a = n->ws + c.seg[s].a_off;
      g.Ca = c.seg[s].C; g.C = c.seg[s].C; g.ch_off = 0; g.taps = c.seg[s].taps;
      g.a_sx = g.Ca; g.a_sy = (int64_t)c.H * g.Ca; g.a_sb = (int64_t)c.H * c.H * g.Ca;
      g.w = n->blob + n->tensors[c.seg[s].t_w].offset;
    }
    d.N = c.N; d.Nrows = c.Nrows;
    d.BN = c.Nrows < 256 ? c.Nrows : 256;
    // few pixel tiles at 16x16 / 32x32: narrow N tiles so that more CTAs share the weight stream
    const long long m_tiles = (long long)B * ((c.H + d.TW - 1) / d.TW) * ((c.H + d.TH - 1) / d.TH);
    while (d.BN > 64 && d.BN % 128 == 0 && m_tiles * (c.Nrows / d.BN) < device_sm_count()) d.BN /= 2;
    d.epi = CHB_EPI_PLAIN; d.act = c.act;
    d.bias = reinterpret_cast<const float*>(n->blob + n->tensors[c.t_b].offset);
    d.out = n->ws + c.out_off; d.out_dtype = c.out_dtype;
    d.o_sn = 1; d.o_sx = c.Nrows; d.o_sy = (int64_t)c.H * c.Nrows; d.o_sb = (int64_t)c.H * c.H * c.Nrows;
    int rc = build_conv_plan(d, &plans[i]);
    if (rc != CHB_OK) return rc;
  }
  return CHB_OK;
}

}  // namespace chb

extern "C" {

int chb_bisenet_create(const chb_bisenet_config* cfg, chb_bisenet** out) {
  if (!cfg || !out || cfg->size < 128 || cfg->size % 128 != 0 || cfg->n_classes <= 0 || cfg->n_classes > 32 ||
      cfg->max_batch <= 0) {
    set_error("chb_bisenet_create: need size a multiple of 128 (512 in the reference), n_classes in 1..32, max_batch > 0");
    return CHB_ERR_ARG;
  }
  chb_bisenet* n = new chb_bisenet();
  n->cfg = *cfg;
  build_bisenet_layout(n);
  *out = n;
  return CHB_OK;
}

void chb_bisenet_destroy(chb_bisenet* n) { delete n; }
int chb_bisenet_num_tensors(const chb_bisenet* n) { return n ? (int)n->tensors.size() : 0; }
int chb_bisenet_tensor_info(const chb_bisenet* n, int i, char* name, int cap, int64_t* offset, int64_t* nbytes,
                            int* dtype) {
  if (!n || i < 0 || i >= (int)n->tensors.size()) {
    set_error("chb_bisenet_tensor_info: index out of range");
    return CHB_ERR_ARG;
  }
  const BTensor& t = n->tensors[i];
  if (name && cap > 0) snprintf(name, cap, "%s", t.name.c_str());
  if (offset) *offset = t.offset;
  if (nbytes) *nbytes = t.nbytes;
  if (dtype) *dtype = t.dtype;
  return CHB_OK;
}
int64_t chb_bisenet_blob_bytes(const chb_bisenet* n) { return n ? n->blob_bytes : 0; }
int64_t chb_bisenet_workspace_bytes(const chb_bisenet* n) { return n ? n->ws_bytes : 0; }
int chb_bisenet_launches(const chb_bisenet* n) {
  // stem, pool, convs, 3 subsample pairs, 4 avg pools, 5 dense, 3 scale/add, tail
  return n ? 2 + (int)n->convs.size() + 6 + 4 + 5 + 3 + 1 : 0;
}

int chb_bisenet_bind(chb_bisenet* n, const void* blob, void* workspace) {
  if (!n || !blob || !workspace || (reinterpret_cast<uintptr_t>(blob) & 255) ||
      (reinterpret_cast<uintptr_t>(workspace) & 1023)) {
    set_error("chb_bisenet_bind: NULL or misaligned pointers (blob 256 B, workspace 1024 B)");
    return CHB_ERR_ARG;
  }
  int rc = chb_check_device();
  if (rc != CHB_OK) return rc;
  n->blob = reinterpret_cast<const uint8_t*>(blob);
  n->ws = reinterpret_cast<uint8_t*>(workspace);
  n->plans.clear();
  return CHB_OK;
}

// img u8 [B,S,S,3] RGB on the device -> label map u8 [B,So,So] (So = out_size, S % So == 0), labels already swapped to
// the CelebAMask-HQ order when the blob's LUT says so (my_parsing_util.py:49-54).  logits_out (optional) receives the
// 1/8-resolution logits fp32 [B,S/8,S/8,32] for checks.
int chb_bisenet_forward(chb_bisenet* n, const uint8_t* img, uint8_t* mask, int out_size, float* logits_out, int B,
                        void* stream_) {
  if (!n || !img || !mask || !n->ws || !n->blob) {
    set_error("chb_bisenet_forward: NULL argument or unbound network");
    return CHB_ERR_ARG;
  }
  const int S = n->cfg.size;
  if (B <= 0 || B > n->cfg.max_batch || out_size <= 0 || out_size > S || S % out_size != 0) {
    set_error("chb_bisenet_forward: need 0 < B <= max_batch and an out_size that divides the network size");
    return CHB_ERR_ARG;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  auto it = n->plans.find(B);
  if (it == n->plans.end()) {
    std::vector<ConvPlan> pl;
    int rc = bisenet_plans(n, B, pl);
    if (rc != CHB_OK) return rc;
    it = n->plans.emplace(B, std::move(pl)).first;
  }
  const std::vector<ConvPlan>& pl = it->second;
  uint8_t* ws = n->ws;
  auto Hp = [&](const std::string& nm) { return reinterpret_cast<__half*>(ws + n->buf.at(nm)); };
  auto Fp = [&](const std::string& nm) { return reinterpret_cast<float*>(ws + n->buf.at(nm)); };
  auto Wf = [&](const std::string& nm) { return reinterpret_cast<const float*>(n->blob + n->tensors[n->tid.at(nm)].offset); };
  int rc = CHB_OK;
  auto conv = [&](const std::string& nm) {
    if (rc == CHB_OK) rc = launch_conv_plan(pl[n->cid.at(nm)], CHB_IMPL_TCGEN05, st);
  };
  auto dense = [&](const float* x, const std::string& w, const char* b, float* y, int Cin, int Cout, int act,
                   int nparts) {
    vec_dense_kernel<<<B, 256, Cin * sizeof(float), st>>>(x, Wf(w), b ? Wf(b) : nullptr, y, Cin, Cout, act, nparts);
  };
  auto pool = [&](const __half* x, float* y, int HW, int C) {   // slabs of >= 64 pixels
    int slabs = HW / 64;
    slabs = slabs < 1 ? 1 : (slabs > kPoolSlabs ? kPoolSlabs : slabs);
    avgpool_kernel<<<dim3(B, slabs), 256, 0, st>>>(x, y, HW, C);
    return slabs;
  };
  const int R = S / 4, H8 = R / 2, H16 = R / 4, H32 = R / 8;
  bisenet_stem_kernel<<<bgrid((long long)B * (S / 16) * (S / 32), 1), 128, 0, st>>>(img, Wf("stem.w"), Wf("stem.b"),
                                                                                  Hp("stem"), B, S);
  maxpool3s2_kernel<<<bgrid((long long)B * R * R * 8, 256), 256, 0, st>>>(Hp("stem"), Hp("p0"), B, S / 2, 64);
  const int ch[4] = {64, 128, 256, 512};
  std::string x = "p0";
  int H = R, Cin = 64;
  for (int li = 0; li < 4; ++li) {
    const int C = ch[li];
    for (int bi = 0; bi < 2; ++bi) {
      const std::string p = "layer" + std::to_string(li + 1) + "." + std::to_string(bi);
      if (li > 0 && bi == 0) {
        conv(p + ".conv1");
        subsample2_kernel<<<bgrid((long long)B * (H / 2) * (H / 2) * (C / 8), 256), 256, 0, st>>>(Hp(p + ".full"),
                                                                                                Hp(p + ".a"), B, H, C);
        subsample2_kernel<<<bgrid((long long)B * (H / 2) * (H / 2) * (Cin / 8), 256), 256, 0, st>>>(Hp(x), Hp(p + ".xs"),
                                                                                                  B, H, Cin);
        H /= 2;
        conv(p + ".conv2");
      } else {
        conv(p + ".conv1");
        conv(p + ".conv2");
      }
      x = p + ".y";
      Cin = C;
    }
  }
  // context path
  int np = pool(Hp("layer4.1.y"), Fp("v_avg32"), H32 * H32, 512);
  dense(Fp("v_avg32"), "conv_avg.w", "conv_avg.b", Fp("v_avg"), 512, 128, 1, np);
  conv("arm32.conv");
  np = pool(Hp("arm32.feat"), Fp("arm32.pool"), H32 * H32, 128);
  dense(Fp("arm32.pool"), "arm32.att.w", "arm32.att.b", Fp("arm32.att"), 128, 128, 2, np);
  scale_add_up_kernel<<<bgrid((long long)B * H16 * H16 * 16, 256), 256, 0, st>>>(Hp("arm32.feat"), Fp("arm32.att"), 0.f,
                                                                               Fp("v_avg"), nullptr, Hp("arm32.sum_up"),
                                                                               B, H32, 128, 2);
  conv("conv_head32");
  conv("arm16.conv");
  np = pool(Hp("arm16.feat"), Fp("arm16.pool"), H16 * H16, 128);
  dense(Fp("arm16.pool"), "arm16.att.w", "arm16.att.b", Fp("arm16.att"), 128, 128, 2, np);
  scale_add_up_kernel<<<bgrid((long long)B * H8 * H8 * 16, 256), 256, 0, st>>>(Hp("arm16.feat"), Fp("arm16.att"), 0.f,
                                                                             nullptr, Hp("arm32.head"),
                                                                             Hp("arm16.sum_up"), B, H16, 128, 2);
  conv("conv_head16");
  // feature fusion
  conv("ffm.convblk");
  np = pool(Hp("ffm.feat"), Fp("ffm.pool"), H8 * H8, 256);
  dense(Fp("ffm.pool"), "ffm.conv1.w", nullptr, Fp("ffm.mid"), 256, 64, 1, np);
  dense(Fp("ffm.mid"), "ffm.conv2.w", nullptr, Fp("ffm.att"), 64, 256, 2, 1);
  scale_add_up_kernel<<<bgrid((long long)B * H8 * H8 * 32, 256), 256, 0, st>>>(Hp("ffm.feat"), Fp("ffm.att"), 1.f, nullptr,
                                                                             nullptr, Hp("ffm.out"), B, H8, 256, 1);
  conv("conv_out.conv");
  conv("conv_out.conv_out");
  if (rc != CHB_OK) return rc;
  const uint8_t* lut = n->blob + n->tensors[n->tid.at("label_lut")].offset;
  logits_to_mask_kernel<<<bgrid((long long)B * out_size * out_size, 256), 256, 0, st>>>(Fp("logits"), lut, mask, B, H8, S,
                                                                                      out_size, n->cfg.n_classes);
  cudaError_t err = cudaSuccess;
  if (logits_out)
    err = cudaMemcpyAsync(logits_out, Fp("logits"), (size_t)B * H8 * H8 * 32 * 4, cudaMemcpyDeviceToDevice, st);
  if (err == cudaSuccess) err = cudaGetLastError();
  if (err != cudaSuccess) {
    set_error(std::string("bisenet launch failed: ") + cudaGetErrorString(err));
    return CHB_ERR_CUDA;
  }
  return CHB_OK;
}

// Pillow-exact bilinear resize of uint8 [B,H,W,C] images on the device (see pil_resample_pass_kernel).  tmp holds the
// x-resampled intermediate [B,H,OW,C]; the tables come from the host (double arithmetic identical to Pillow's).
int chb_pil_resize_bilinear(const uint8_t* in, uint8_t* tmp, uint8_t* out, int B, int H, int W, int C, int OH, int OW,
                            const int* xbounds, const int* xcoef, int xks, const int* ybounds, const int* ycoef, int yks,
                            void* stream_) {
  if (!in || !tmp || !out || !xbounds || !xcoef || !ybounds || !ycoef || B <= 0 || H <= 0 || W <= 0 || C <= 0 || OH <= 0 ||
      OW <= 0 || xks <= 0 || yks <= 0) {
    set_error("chb_pil_resize_bilinear: bad arguments");
    return CHB_ERR_ARG;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  pil_resample_pass_kernel<<<bgrid((long long)B * H * OW * C, 256), 256, 0, st>>>(in, tmp, xbounds, xcoef, xks,
                                                                                (long long)B * H, W, OW, C);
  const int row = OW * C;
  if (row % 4 == 0 && ((reinterpret_cast<uintptr_t>(tmp) | reinterpret_cast<uintptr_t>(out)) & 3) == 0)
    pil_resample_pass4_kernel<<<bgrid((long long)B * OH * (row / 4), 256), 256, 0, st>>>(tmp, out, ybounds, ycoef, yks, B, H,
                                                                                       OH, row);
  else
    pil_resample_pass_kernel<<<bgrid((long long)B * OH * row, 256), 256, 0, st>>>(tmp, out, ybounds, ycoef, yks, B, H, OH,
                                                                                 row);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    set_error(std::string("pil_resize launch failed: ") + cudaGetErrorString(err));
    return CHB_ERR_CUDA;
  }
  return CHB_OK;
}

// Host buffers in and out: img u8 [B,S,S,3], mask u8 [B,So,So]; copies on `stream`, returns when the mask has landed.
int chb_bisenet_forward_host(chb_bisenet* n, const uint8_t* img_host, uint8_t* mask_host, int out_size, int B,
                             void* stream_) {
  if (!n || !img_host || !mask_host || !n->ws) {
    set_error("chb_bisenet_forward_host: NULL argument or unbound network");
    return CHB_ERR_ARG;
  }
  if (B <= 0 || B > n->cfg.max_batch) {
    set_error("chb_bisenet_forward_host: batch exceeds max_batch");
    return CHB_ERR_ARG;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  const size_t S = (size_t)n->cfg.size;
  uint8_t* d_img = n->ws + n->buf.at("img");
  uint8_t* d_mask = n->ws + n->buf.at("mask");
  cudaError_t err = cudaMemcpyAsync(d_img, img_host, (size_t)B * S * S * 3, cudaMemcpyHostToDevice, st);
  if (err != cudaSuccess) {
    set_error(std::string("H2D copy failed: ") + cudaGetErrorString(err));
    return CHB_ERR_CUDA;
  }
  int rc = chb_bisenet_forward(n, d_img, d_mask, out_size, nullptr, B, stream_);
  if (rc != CHB_OK) return rc;
  err = cudaMemcpyAsync(mask_host, d_mask, (size_t)B * out_size * out_size, cudaMemcpyDeviceToHost, st);
  if (err == cudaSuccess) err = cudaStreamSynchronize(st);
  if (err != cudaSuccess) {
    set_error(std::string("D2H copy / sync failed: ") + cudaGetErrorString(err));
    return CHB_ERR_CUDA;
  }
  return CHB_OK;
}

}  // extern "C"
