// Fused small-MLP forward for the colour/texture branch (all three nets are chains of <= 6 dense layers of width
// <= 512: color_texture_branch/model_eigengan.py:34-83 EigenGenerator, model.py:86-127 Discriminator used as the
// code encoder, predictor/predictor_model.py:14-41 Predictor).  One CTA per sample keeps the activations in
// shared memory for the whole chain; weights are stored transposed ([in][out]) so that a warp reads 128
// contiguous bytes per k.  fp32 throughout: these nets are launch-latency bound, not FLOP bound.
#include <cuda_runtime.h>

#include <string>

#include "../../include/ctrlhair_b200.h"
#include "conv_igemm.cuh"

namespace chb {

constexpr int kMlpMaxLayers = 8;
constexpr int kMlpMaxDim = 1024;
constexpr int kMlpThreads = 512;
constexpr int kMlpWarps = kMlpThreads / 32;

struct MlpParams {
  chb_mlp_layer layer[kMlpMaxLayers];
  int nlayers;
  const float* x;
  const float* z;
  int zdim;
  float* out;
  int B;
};

__device__ __forceinline__ float mlp_act(float v, int act) {
  if (act == CHB_ACT_RELU) return fmaxf(v, 0.f);
  if (act == CHB_ACT_LRELU) return v > 0.f ? v : 0.2f * v;
  if (act == CHB_ACT_TANH) return tanhf(v);
  return v;
}

__global__ void __launch_bounds__(kMlpThreads) mlp_chain_kernel(const MlpParams p) {
  __shared__ float buf[2][kMlpMaxDim];
  __shared__ float part[kMlpWarps * 32];   // k-sliced partial sums of a narrow layer (slices x out_dim <= 512)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
    float* cur = buf[0];
    float* nxt = buf[1];
    const int in0 = p.layer[0].in_dim;
    for (int i = threadIdx.x; i < in0; i += kMlpThreads) cur[i] = p.x[(long long)b * in0 + i];
    __syncthreads();
    for (int l = 0; l < p.nlayers; ++l) {
      const chb_mlp_layer& L = p.layer[l];
      // input-side ops: subspace injection  x += (L * z) @ U + mu  (model_eigengan.py:24), then the activation that
      // precedes the Linear inside nn.Sequential(LeakyReLU, Linear) (model_eigengan.py:50-53)
      if (L.inj_nb > 0 || L.pre_act != CHB_ACT_NONE) {
        for (int i = threadIdx.x; i < L.in_dim; i += kMlpThreads) {
          float v = cur[i];
          if (L.inj_nb > 0) {
            float s = L.inj_mu[i];
            for (int k = 0; k < L.inj_nb; ++k)
              s = fmaf(L.inj_l[k] * p.z[(long long)b * p.zdim + L.inj_zoff + k], L.inj_u[k * L.in_dim + i], s);
            v += s;
          }
          cur[i] = mlp_act(v, L.pre_act);
        }
        __syncthreads();
      }
      // A warp takes 32 consecutive outputs (one 128-byte weight row per k); layers narrower than 16 x 32 outputs also
      // split K across the spare warps (the chain over k, not the FMAs, is what a small dense layer costs).  Four
      // independent accumulators and sixteen weight loads in flight per thread.
      const int O = L.out_dim, I = L.in_dim;
      const int nb = (O + 31) >> 5;
      int ks = 1;
      while (ks * 2 * nb <= kMlpWarps) ks *= 2;
      const bool last = l + 1 == p.nlayers;
      auto finish = [&](int j, float acc) {
        acc = mlp_act(acc, L.post_act);
        if (last) p.out[(long long)b * O + j] = acc;
        else nxt[j] = acc;
      };
      for (int it = warp; it < nb * ks; it += kMlpWarps) {
        const int blk = it % nb, sl = it / nb;
        const int j = blk * 32 + lane;
        if (j >= O) continue;
        const int k0 = (int)((long long)I * sl / ks), k1 = (int)((long long)I * (sl + 1) / ks);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        const float* w = L.wt + j;
        const long long od = O;
        int k = k0;
        for (; k + 16 <= k1; k += 16) {
          float wv[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) wv[u] = __ldg(w + (k + u) * od);
#pragma unroll
          for (int u = 0; u < 16; u += 4) {
            a0 = fmaf(cur[k + u], wv[u], a0);
            a1 = fmaf(cur[k + u + 1], wv[u + 1], a1);
            a2 = fmaf(cur[k + u + 2], wv[u + 2], a2);
            a3 = fmaf(cur[k + u + 3], wv[u + 3], a3);
          }
        }
        for (; k < k1; ++k) a0 = fmaf(cur[k], __ldg(w + k * od), a0);
        const float acc = (a0 + a1) + (a2 + a3);
        if (ks == 1) finish(j, acc + (L.bias ? L.bias[j] : 0.f));
        else part[sl * O + j] = acc;
      }
      if (ks > 1) {
        __syncthreads();
        for (int j = threadIdx.x; j < O; j += kMlpThreads) {
          float acc = L.bias ? L.bias[j] : 0.f;
          for (int sl = 0; sl < ks; ++sl) acc += part[sl * O + j];   // fixed order
          finish(j, acc);
        }
      }
      __syncthreads();
      float* t = cur;
      cur = nxt;
      nxt = t;
    }
  }
}

}  // namespace chb

extern "C" int chb_mlp_forward(const chb_mlp_layer* layers, int nlayers, const float* x, const float* z, int zdim,
                               float* out, int B, void* stream) {
  using namespace chb;
  if (!layers || nlayers <= 0 || nlayers > kMlpMaxLayers || !x || !out || B <= 0) {
    set_error("chb_mlp_forward: bad arguments (1..8 layers, non-NULL buffers, B > 0)");
    return CHB_ERR_ARG;
  }
  int rc = chb_check_device();
  if (rc != CHB_OK) return rc;
  MlpParams p;
  p.nlayers = nlayers; p.x = x; p.z = z; p.zdim = zdim; p.out = out; p.B = B;
  for (int l = 0; l < nlayers; ++l) {
    const chb_mlp_layer& L = layers[l];
    if (L.in_dim <= 0 || L.out_dim <= 0 || L.in_dim > kMlpMaxDim || L.out_dim > kMlpMaxDim || !L.wt ||
        (l > 0 && L.in_dim != layers[l - 1].out_dim) ||
        (L.inj_nb > 0 && (!z || !L.inj_u || !L.inj_l || !L.inj_mu || L.inj_zoff + L.inj_nb > zdim))) {
      set_error("chb_mlp_forward: inconsistent layer description");
      return CHB_ERR_ARG;
    }
    p.layer[l] = L;
  }
  const int grid = B < 4 * device_sm_count() ? B : 4 * device_sm_count();
  mlp_chain_kernel<<<grid, kMlpThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    set_error(std::string("mlp launch failed: ") + cudaGetErrorString(err));
    return CHB_ERR_CUDA;
  }
  return CHB_OK;
}
