// Small memory-bound helpers around the conv operator: label -> one-hot pyramid, noise planes, dtype casts.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <string>

#include "../../include/ctrlhair_b200.h"
#include "conv_igemm.cuh"

namespace chb {

constexpr int kMaxLevels = 8;
struct PyramidParams {
  const uint8_t* labels;
  int B, S, nlevels, nclass;
  int ones_ch0, ones_n;  // channels [ones_ch0, ones_ch0 + ones_n) are written as 1.0 at every pixel (bias carriers)
  int shift[kMaxLevels];
  __half* out[kMaxLevels];
  long long first_pix[kMaxLevels + 1];  // prefix sum of B*r*r per level
};

// One thread per (level, image, pixel): writes the 32-channel fp16 one-hot row (64 B) of that pixel.
// Follows pix2pix_model.py:136-141 (zeros().scatter_(1, label, 1.0)) and the nearest resize
// F.interpolate(seg, (r, r)) of normalization.py:115 / generator.py:75: src index = dst << shift.
__global__ void onehot_pyramid_kernel(const PyramidParams p) {
  const long long total = p.first_pix[p.nlevels];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int l = 0;
    while (i >= p.first_pix[l + 1]) ++l;
    long long pix = i - p.first_pix[l];
    const int sh = p.shift[l];
    const int r = p.S >> sh;
    const int x = (int)(pix % r);
    const int y = (int)((pix / r) % r);
    const int b = (int)(pix / ((long long)r * r));
    const int lab = p.labels[((long long)b * p.S + ((long long)y << sh)) * p.S + ((long long)x << sh)];
    uint32_t w[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) w[k] = 0u;
    if (lab < p.nclass) {
      const uint32_t one = 0x3C00u;  // fp16 1.0
#pragma unroll
      for (int k = 0; k < 16; ++k)
        if (k == (lab >> 1)) w[k] = (lab & 1) ? (one << 16) : one;
    }
    for (int c = p.ones_ch0; c < p.ones_ch0 + p.ones_n; ++c) {
      const uint32_t one = 0x3C00u;
#pragma unroll
      for (int k = 0; k < 16; ++k)
        if (k == (c >> 1)) w[k] |= (c & 1) ? (one << 16) : one;
    }
    uint4* o = reinterpret_cast<uint4*>(p.out[l] + pix * 32);
    o[0] = make_uint4(w[0], w[1], w[2], w[3]);
    o[1] = make_uint4(w[4], w[5], w[6], w[7]);
    o[2] = make_uint4(w[8], w[9], w[10], w[11]);
    o[3] = make_uint4(w[12], w[13], w[14], w[15]);
  }
}

// Philox4x32-10 counter-based generator (Salmon et al. 2011), 4 normals per counter via Box-Muller.
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

__global__ void noise_fill_kernel(float* out, long long n, uint64_t seed, uint64_t offset) {
  const long long quads = (n + 3) / 4;  // (`seed` is a plain kernel parameter: a captured CUDA graph re-seeds this node
                                        //  with cudaGraphExecKernelNodeSetParams, see chb_generator_forward_graph)
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < quads;
       i += (long long)gridDim.x * blockDim.x) {
    const uint64_t ctr = (uint64_t)i + offset;
    uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0x243F6A88u, 0x85A308D3u};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      philox_round(c, k0, k1);
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    float z[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float u1 = ((float)c[2 * h] + 1.0f) * 2.3283064365386963e-10f;  // (0, 1]
      const float u2 = (float)c[2 * h + 1] * 2.3283064365386963e-10f;
      const float rad = sqrtf(-2.0f * __logf(u1));
      float sn, cs;
      __sincosf(6.283185307179586f * u2, &sn, &cs);
      z[2 * h] = rad * cs;
      z[2 * h + 1] = rad * sn;
    }
    const long long base = i * 4;
    if (base + 4 <= n) {
      *reinterpret_cast<float4*>(out + base) = make_float4(z[0], z[1], z[2], z[3]);
    } else {
      for (int k = 0; k < 4 && base + k < n; ++k) out[base + k] = z[k];
    }
  }
}

__global__ void f32_to_f16_kernel(const float* in, __half* out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = __float2half_rn(in[i]);
}

// codes fp32 [B][NC][L] -> fp16 [NC][B][L] (class-major, so that "image == class" for the grouped fc_mu GEMM)
__global__ void codes_cast_kernel(const float* in, __half* out, int B, int NC, int L) {
  const long long n = (long long)B * NC * L;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int l = (int)(i % L);
    const int j = (int)((i / L) % NC);
    const int b = (int)(i / ((long long)L * NC));
    out[((long long)j * B + b) * L + l] = __float2half_rn(in[i]);
  }
}

static int check_launch(const char* what) {
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    set_error(std::string(what) + " launch failed: " + cudaGetErrorString(err));
    return CHB_ERR_CUDA;
  }
  return CHB_OK;
}

static int grid_for(long long n, int threads) {
  long long g = (n + threads - 1) / threads;
  const long long cap = (long long)device_sm_count() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// image = tanh(bias + sum over the 3x3 neighbourhood of the per-tap partial sums)   (generator.py:107-108)
//   y   fp32 [B,S,S,32]: y[p][tap*3 + co] = sum_ci x[p][ci] * W[co][ci][tap]  (the 1x1 GEMM "conv_img.taps")
//   out fp32 [B,3,S,S]:  out[co][p] = tanh(bias[co] + sum_tap y[p + d(tap)][tap*3 + co]), zero outside the image
// HBM-bound: 128 B read + 12 B written per pixel.  One CTA = 32 x 8 pixels; the (8+2) x (32+2) halo of y rows is staged
// in shared memory with a 33-float row pitch so that the 27 reads per pixel are bank-conflict free.
constexpr int kGatherTX = 32, kGatherTY = 8, kGatherPitch = 33;
__global__ void __launch_bounds__(kGatherTX * kGatherTY) img_from_taps_kernel(const float* __restrict__ y,
                                                                              const float* __restrict__ bias,
                                                                              float* __restrict__ out, int S) {
  __shared__ float t[(kGatherTY + 2) * (kGatherTX + 2) * kGatherPitch];
  const int tiles_x = S / kGatherTX, tiles_y = S / kGatherTY;
  int tile = blockIdx.x;
  const int x0 = (tile % tiles_x) * kGatherTX;
  tile /= tiles_x;
  const int y0 = (tile % tiles_y) * kGatherTY;
  const int b = tile / tiles_y;
  const float* yb = y + (size_t)b * S * S * 32;
  constexpr int HW = kGatherTX + 2, HP = (kGatherTY + 2) * HW;
  for (int i = threadIdx.x; i < HP * 8; i += kGatherTX * kGatherTY) {
    const int pix = i >> 3, q = i & 7;
    const int py = pix / HW, px = pix - py * HW;
    const int yy = y0 - 1 + py, xx = x0 - 1 + px;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (yy >= 0 && yy < S && xx >= 0 && xx < S)
      v = __ldg(reinterpret_cast<const float4*>(yb + ((size_t)yy * S + xx) * 32) + q);
    float* d = t + pix * kGatherPitch + q * 4;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  __syncthreads();
  const int lx = threadIdx.x & (kGatherTX - 1), ly = threadIdx.x / kGatherTX;
  float a0 = __ldg(bias), a1 = __ldg(bias + 1), a2 = __ldg(bias + 2);
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      // output pixel (ly, lx) reads the neighbour at offset (ky-1, kx-1), whose partial sum for THIS tap is its
      // contribution to (ly, lx): conv zero padding == the zero rows staged for out-of-image neighbours
      const float* r = t + ((ly + ky) * HW + lx + kx) * kGatherPitch + (ky * 3 + kx) * 3;
      a0 += r[0]; a1 += r[1]; a2 += r[2];
    }
  const size_t plane = (size_t)S * S;
  float* o = out + (size_t)b * 3 * plane + (size_t)(y0 + ly) * S + x0 + lx;
  o[0] = tanhf(a0);
  o[plane] = tanhf(a1);
  o[2 * plane] = tanhf(a2);
}

int img_from_taps(const float* y, const float* bias, float* out, int B, int S, cudaStream_t stream) {
  if (!y || !bias || !out || B <= 0 || S < kGatherTX || S % kGatherTX != 0) {
    set_error("img_from_taps: bad arguments (S must be a multiple of 32)");
    return CHB_ERR_ARG;
  }
  img_from_taps_kernel<<<B * (S / kGatherTX) * (S / kGatherTY), kGatherTX * kGatherTY, 0, stream>>>(y, bias, out, S);
  return check_launch("img_from_taps");
}

int codes_cast_transpose(const float* in, void* out, int B, int NC, int L, cudaStream_t stream) {
  codes_cast_kernel<<<grid_for((long long)B * NC * L, 256), 256, 0, stream>>>(in, reinterpret_cast<__half*>(out), B,
                                                                             NC, L);
  return check_launch("codes_cast");
}

}  // namespace chb

namespace chb {
// Same, with channels [ones_ch0, ones_ch0 + ones_n) set to 1.0 at every pixel.  The generator uses two such channels
// to carry conv biases through the GEMM: sum_j one_hot[p][j] = 1, and the centre tap of a zero-padded 3x3 conv never
// leaves the image, so weight[n][centre tap][ones channel] = bias[n] (split hi + lo in fp16) adds the bias exactly.
int onehot_pyramid_ones(const uint8_t* labels, int B, int S, int nlevels, const int* shifts, void* const* outs,
                             int nclass, int ones_ch0, int ones_n, void* stream) {
  if (ones_n < 0 || (ones_n > 0 && (ones_ch0 < nclass || ones_ch0 + ones_n > 32))) {
    set_error("onehot_pyramid: constant channels must lie in [nclass, 32)");
    return CHB_ERR_ARG;
  }
  if (!labels || !shifts || !outs || B <= 0 || S <= 0 || nlevels <= 0 || nlevels > kMaxLevels || nclass <= 0 ||
      nclass > 32) {
    set_error("chb_onehot_pyramid: bad arguments (need 1..8 levels, 1..32 classes)");
    return CHB_ERR_ARG;
  }
  PyramidParams p;
  p.labels = labels; p.B = B; p.S = S; p.nlevels = nlevels; p.nclass = nclass;
  p.ones_ch0 = ones_ch0; p.ones_n = ones_n;
  p.first_pix[0] = 0;
  for (int l = 0; l < nlevels; ++l) {
    if (shifts[l] < 0 || (S >> shifts[l]) <= 0 || ((S >> shifts[l]) << shifts[l]) != S || !outs[l]) {
      set_error("chb_onehot_pyramid: level resolution must divide S");
      return CHB_ERR_ARG;
    }
    p.shift[l] = shifts[l];
    p.out[l] = reinterpret_cast<__half*>(outs[l]);
    const long long r = S >> shifts[l];
    p.first_pix[l + 1] = p.first_pix[l] + (long long)B * r * r;
  }
  onehot_pyramid_kernel<<<grid_for(p.first_pix[nlevels], 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("onehot_pyramid");
}

}  // namespace chb

extern "C" {

int chb_onehot_pyramid(const uint8_t* labels, int B, int S, int nlevels, const int* shifts, void* const* outs,
                       int nclass, void* stream) {
  return chb::onehot_pyramid_ones(labels, B, S, nlevels, shifts, outs, nclass, 0, 0, stream);
}

const void* chb_noise_fill_kernel_address(void) { return reinterpret_cast<const void*>(&chb::noise_fill_kernel); }

int chb_noise_fill(float* out, int64_t n, uint64_t seed, uint64_t offset, void* stream) {
  using namespace chb;
  if (!out || n <= 0 || (reinterpret_cast<uintptr_t>(out) & 15)) {
    set_error("chb_noise_fill: need a 16-byte aligned buffer and n > 0");
    return CHB_ERR_ARG;
  }
  noise_fill_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(out, n, seed,
                                                                                                   offset);
  return check_launch("noise_fill");
}

int chb_f32_to_f16(const float* in, void* out, int64_t n, void* stream) {
  using namespace chb;
  if (!in || !out || n <= 0) {
    set_error("chb_f32_to_f16: bad arguments");
    return CHB_ERR_ARG;
  }
  f32_to_f16_kernel<<<grid_for(n, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      in, reinterpret_cast<__half*>(out), n);
  return check_launch("f32_to_f16");
}

}  // extern "C"
