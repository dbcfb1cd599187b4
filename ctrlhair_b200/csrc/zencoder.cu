// SEAN style encoder (Zencoder) forward: sean_codes/models/networks/architecture.py:154-207.
//
//   L1  ReflectionPad(1) + conv3x3(3->32)               -> direct SIMT conv (K = 27), fp32 NHWC
//   L2  conv3x3 s2 (32->64), L3 conv3x3 s2 (64->128)    -> tcgen05 conv at the input resolution, the stride-2 phase is
//                                                          picked by the InstanceNorm kernels that follow
//   L4  ConvTranspose3x3 s2 p1 op1 (128->256)           -> zero-insertion + tcgen05 conv3x3 with the flipped kernel
//   L5  ReflectionPad(1) + conv3x3(256->512) + tanh     -> tcgen05 conv over the explicitly padded tensor (a_pad = 1)
//   InstanceNorm2d(affine=False, eps 1e-5) + LeakyReLU(0.2) after L1..L4: stats kernel (double accumulation) + apply
//   kernel that also produces the layout the next conv wants (fp16 NHWC; plain / zero-inserted / reflection-padded)
//   region pooling (architecture.py:193-207): per (image, class) mean of the 512-channel code over the region,
//   zero for absent classes.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/ctrlhair_b200.h"
#include "conv_igemm.cuh"

namespace chb {

// ---------------------------------------------------------------- L1: 3 -> 32, reflection pad, fp32
__global__ void __launch_bounds__(128) zenc_conv1_kernel(const float* __restrict__ img,
                                                         const float* __restrict__ w /*[32][27]: co, (ky,kx,ci)*/,
                                                         const float* __restrict__ bias, float* __restrict__ out, int B,
                                                         int S) {
  // weights padded to 28 per output channel so that four taps come with one 16-byte shared-memory load: the kernel was
  // bound by one broadcast LDS per FMA (864 per pixel), now 224 LDS.128 per pixel
  __shared__ __align__(16) float sw[32 * 28];
  __shared__ float sb[32];
  for (int i = threadIdx.x; i < 32 * 28; i += blockDim.x) {
    const int co = i / 28, k = i - co * 28;
    sw[i] = k < 27 ? w[co * 27 + k] : 0.f;
  }
  if (threadIdx.x < 32) sb[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const long long total = (long long)B * S * S;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % S);
    const int y = (int)((i / S) % S);
    const int b = (int)(i / ((long long)S * S));
    float in[28];
    in[27] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      int yy = y + ky - 1;
      yy = yy < 0 ? -yy : (yy >= S ? 2 * S - 2 - yy : yy);  // ReflectionPad2d(1)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        int xx = x + kx - 1;
        xx = xx < 0 ? -xx : (xx >= S ? 2 * S - 2 - xx : xx);
#pragma unroll
        for (int c = 0; c < 3; ++c) in[(ky * 3 + kx) * 3 + c] = __ldg(img + (((long long)b * 3 + c) * S + yy) * S + xx);
      }
    }
    float4* o = reinterpret_cast<float4*>(out + i * 32);
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int co = g * 4 + j;
        const float4* wr = reinterpret_cast<const float4*>(sw + co * 28);
        float acc = sb[co];
#pragma unroll
        for (int q = 0; q < 7; ++q) {   // same accumulation order as before: k = 0 .. 26 (the 28th term is 0 * 0)
          const float4 wv = wr[q];
          acc = fmaf(in[4 * q], wv.x, acc);
          acc = fmaf(in[4 * q + 1], wv.y, acc);
          acc = fmaf(in[4 * q + 2], wv.z, acc);
          acc = fmaf(in[4 * q + 3], wv.w, acc);
        }
        v[j] = acc;
      }
      o[g] = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}

// ---------------------------------------------------------------- InstanceNorm statistics
// x fp32 NHWC [B, Hs, Ws, C]; the normalised map is the sub-grid (y*s, x*s), y < H, x < W.  Per (b, c) sum and sum of
// squares in double.  grid = (slabs, B), block = 256: a thread owns 4 consecutive channels (one 16-byte load per
// pixel), 256 / (C/4) pixels are in flight per block step; the block combines its lanes in shared memory and issues
// ONE atomic per (channel, moment) — HBM-bound: algorithmic bytes = 4 B per element of the sub-grid.
__global__ void __launch_bounds__(256) in_stats_kernel(const float* __restrict__ x, double* __restrict__ sums, int H,
                                                       int W, int C, int s, long long row_stride,
                                                       long long img_stride) {
  __shared__ double part[256 * 8];   // [lane][channel][2], lanes * C = 1024 entries
  const int b = blockIdx.y;
  const int c4 = C >> 2;
  const int lanes = 256 / c4;
  const int cg = threadIdx.x % c4, pl = threadIdx.x / c4;
  const int npix = H * W;
  const float* xb = x + (long long)b * img_stride + cg * 4;
  double d[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int cnt = 0;
  const int step = gridDim.x * lanes;
  for (int p = blockIdx.x * lanes + pl; p < npix; p += 4 * step) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int q = p + u * step;
      if (q < npix) {
        const int yy = q / W, xx = q - yy * W;
        v[u] = __ldg(reinterpret_cast<const float4*>(xb + (long long)(yy * s) * row_stride + (long long)(xx * s) * C));
      } else {
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      f[0] += v[u].x; f[1] = fmaf(v[u].x, v[u].x, f[1]);
      f[2] += v[u].y; f[3] = fmaf(v[u].y, v[u].y, f[3]);
      f[4] += v[u].z; f[5] = fmaf(v[u].z, v[u].z, f[5]);
      f[6] += v[u].w; f[7] = fmaf(v[u].w, v[u].w, f[7]);
    }
    if (++cnt == 16) {  // flush to double regularly: fp32 running sums stay short (64 pixels)
#pragma unroll
      for (int k = 0; k < 8; ++k) { d[k] += f[k]; f[k] = 0.f; }
      cnt = 0;
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) part[(pl * c4 + cg) * 8 + k] = d[k] + (double)f[k];
  __syncthreads();
  // thread t < 2C: (channel, moment) = (t / 2, t % 2); part index of lane l = (l * c4 + ch / 4) * 8 + (ch % 4) * 2 + m
  for (int t = threadIdx.x; t < 2 * C; t += 256) {
    const int ch = t >> 1, m = t & 1;
    double acc = 0.0;
    for (int l = 0; l < lanes; ++l) acc += part[(l * c4 + (ch >> 2)) * 8 + (ch & 3) * 2 + m];
    atomicAdd(&sums[((long long)b * C + ch) * 2 + m], acc);
  }
}

// (b, c): {rstd, -mean * rstd} in fp32 from the double moments (biased variance, eps 1e-5, as nn.InstanceNorm2d)
__global__ void in_finalize_kernel(const double* __restrict__ sums, float2* __restrict__ ss, int n, double inv_n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double mean = sums[2 * i] * inv_n;
  const double var = sums[2 * i + 1] * inv_n - mean * mean;
  const double r = 1.0 / sqrt(var + 1e-5);
  ss[i] = make_float2((float)r, (float)(-mean * r));
}

// ---------------------------------------------------------------- InstanceNorm apply + LeakyReLU + relayout (fp16 out)
// mode 0: out[b, y, x]       = f(in[y, x])                        out extent H x W
// mode 1: zero insertion     out[b, 2y, 2x] = f(in[y, x]), else 0  out extent 2H x 2W   (ConvTranspose s2 as a conv)
// mode 2: reflection pad 1   out[b, y, x] = f(in[refl(y-1), refl(x-1)])  out extent (H+2) x (W+2)
// One thread per 8 channels of an output pixel: two 16-byte loads, one 16-byte store, 8 FMAs (HBM-bound: 4 B read +
// 2 B written per element).
__global__ void __launch_bounds__(256) in_apply_kernel(const float* __restrict__ x, const float2* __restrict__ ss,
                                                       __half* __restrict__ out, int B, int H, int W, int C, int s,
                                                       long long row_stride, long long img_stride, int mode) {
  const int OH = mode == 1 ? 2 * H : (mode == 2 ? H + 2 : H);
  const int OW = mode == 1 ? 2 * W : (mode == 2 ? W + 2 : W);
  const int c8 = C / 8;
  const unsigned per_img = (unsigned)OH * OW * c8;
  const long long total = (long long)B * per_img;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / per_img);
    unsigned p = (unsigned)(i - (long long)b * per_img);
    const int cg = (int)(p % c8);
    p /= c8;
    const int ox = (int)(p % OW), oy = (int)(p / OW);
    int iy = oy, ix = ox;
    bool zero = false;
    if (mode == 1) {
      zero = (oy & 1) || (ox & 1);
      iy = oy >> 1; ix = ox >> 1;
    } else if (mode == 2) {
      iy = oy - 1; ix = ox - 1;
      iy = iy < 0 ? -iy : (iy >= H ? 2 * H - 2 - iy : iy);
      ix = ix < 0 ? -ix : (ix >= W ? 2 * W - 2 - ix : ix);
    }
    uint32_t pk[4] = {0u, 0u, 0u, 0u};
    if (!zero) {
      const float* src = x + (long long)b * img_stride + (long long)(iy * s) * row_stride + (long long)(ix * s) * C + cg * 8;
      const float4 a = __ldg(reinterpret_cast<const float4*>(src)), bq = __ldg(reinterpret_cast<const float4*>(src + 4));
      const float v[8] = {a.x, a.y, a.z, a.w, bq.x, bq.y, bq.z, bq.w};
      const float4* sp = reinterpret_cast<const float4*>(ss + (long long)b * C + cg * 8);
      float o[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 q = __ldg(sp + k);   // {rstd, shift} of channels 2k, 2k + 1
        const float t0 = fmaf(v[2 * k], q.x, q.y), t1 = fmaf(v[2 * k + 1], q.z, q.w);
        o[2 * k] = t0 > 0.f ? t0 : 0.2f * t0;
        o[2 * k + 1] = t1 > 0.f ? t1 : 0.2f * t1;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        __half2 h = __floats2half2_rn(o[2 * k], o[2 * k + 1]);
        pk[k] = *reinterpret_cast<uint32_t*>(&h);
      }
    }
    *reinterpret_cast<uint4*>(out + (((long long)b * OH + oy) * OW + ox) * C + cg * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

// ---------------------------------------------------------------- region pooling
// codes fp32 NHWC [B, R, R, C]; labels u8 [B, S, S], class of code pixel (y,x) = labels[y << sh][x << sh]
// (F.interpolate nearest, architecture.py:181).  grid = (slabs, B), block = C / 4 threads (one float4 of channels each).
// A block walks its pixels in order with 8 loads in flight and keeps the running sum of the current label run in
// registers; a run is flushed into the block's [NC][C] shared accumulator when the label changes (label maps are
// piecewise constant along a row), and the accumulator goes out with one atomic per touched entry.  HBM-bound: 4 B per
// code element.
__global__ void __launch_bounds__(128) region_pool_kernel(const float* __restrict__ codes,
                                                          const uint8_t* __restrict__ labels,
                                                          float* __restrict__ sums /*[B][NC][C]*/,
                                                          int* __restrict__ counts /*[B][NC]*/, int R, int C, int S,
                                                          int sh, int NC, int pix_per_block) {
  extern __shared__ float acc[];  // [NC][C]
  __shared__ int cnt[32];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < NC * C; i += blockDim.x) acc[i] = 0.f;
  if (threadIdx.x < 32) cnt[threadIdx.x] = 0;
  __syncthreads();
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(R * R, p0 + pix_per_block);
  const int c = threadIdx.x * 4;
  const uint8_t* lb = labels + (long long)b * S * S;
  const float* cb = codes + (long long)b * R * R * C + c;
  int cur = -1, run = 0;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  auto flush = [&]() {
    if (cur >= 0 && cur < NC) {   // thread-private columns of the accumulator: no conflicts
      float* d = acc + cur * C + c;
      d[0] += a.x; d[1] += a.y; d[2] += a.z; d[3] += a.w;
      if (threadIdx.x == 0) cnt[cur] += run;
    }
  };
  for (int p = p0; p < p1; p += 8) {
    float4 v[8];
    int lab[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int q = p + u;
      if (q < p1) {
        const int y = q / R, x = q - y * R;
        lab[u] = lb[((long long)y << sh) * S + ((long long)x << sh)];
        v[u] = __ldg(reinterpret_cast<const float4*>(cb + (long long)q * C));
      } else {
        lab[u] = -2;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (lab[u] == -2) continue;
      if (lab[u] != cur) {
        flush();
        cur = lab[u]; run = 0;
        a = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      a.x += v[u].x; a.y += v[u].y; a.z += v[u].z; a.w += v[u].w;
      ++run;
    }
  }
  flush();
  __syncthreads();
  for (int i = threadIdx.x; i < NC * C; i += blockDim.x) {
    const float v = acc[i];
    if (v != 0.f) atomicAdd(&sums[(long long)b * NC * C + i], v);
  }
  if (threadIdx.x < NC && cnt[threadIdx.x]) atomicAdd(&counts[b * NC + threadIdx.x], cnt[threadIdx.x]);
}

__global__ void region_finalize_kernel(const float* __restrict__ sums, const int* __restrict__ counts,
                                       float* __restrict__ out, long long n, int C) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int area = counts[i / C];
    out[i] = area > 0 ? sums[i] / (float)area : 0.f;
  }
}

static int zgrid(long long n, int threads) {
  long long g = (n + threads - 1) / threads;
  const long long cap = (long long)device_sm_count() * 16;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

struct ZTensor {
  std::string name;
  int64_t offset, nbytes;
  int dtype;
};

}  // namespace chb

using namespace chb;

struct chb_zencoder {
  chb_zenc_config cfg;
  std::vector<ZTensor> tensors;
  int64_t blob_bytes = 0, ws_bytes = 0;
  int t_w1, t_b1, t_w[4], t_b[4];
  int64_t ws_img, ws_labels, ws_out, ws_c1, ws_a1, ws_c2, ws_a2, ws_c3, ws_z3, ws_c4, ws_a4, ws_codes, ws_sums, ws_psum,
      ws_pcnt;
  const uint8_t* blob = nullptr;
  uint8_t* ws = nullptr;
  std::map<int, std::vector<ConvPlan>> plans;
};

namespace chb {
static int zadd(chb_zencoder* z, const char* name, int64_t nbytes, int dtype) {
  ZTensor t{name, z->blob_bytes, nbytes, dtype};
  z->tensors.push_back(t);
  z->blob_bytes += (nbytes + 255) / 256 * 256;
  return (int)z->tensors.size() - 1;
}
static int64_t zws(chb_zencoder* z, int64_t n) {
  const int64_t o = z->ws_bytes;
  z->ws_bytes += (n + 1023) / 1024 * 1024;
  return o;
}
static const int kZC[5] = {32, 64, 128, 256, 512};  // channels after L1..L5

static int zencoder_plans(chb_zencoder* z, int B, std::vector<ConvPlan>& plans) {
  const int S = z->cfg.crop, S2 = S / 2;
  uint8_t* ws = z->ws;
  const uint8_t* blob = z->blob;
  auto mk = [&](const void* a, int Hin, int Win, int C, const void* w, const float* bias, int N, int H, int W,
                int a_pad, void* out, int act, ConvPlan* plan) -> int {
    chb_conv_desc d;
    memset(&d, 0, sizeof d);
    d.B = B; d.H = H; d.W = W;
    d.TW = 8; d.TH = 16; d.TB = 1;
    d.nseg = 1;
    chb_conv_seg& s = d.seg[0];
    s.a = a; s.Ca = C; s.C = C; s.taps = 9; s.w = w; s.a_pad = a_pad;
    s.a_sx = C; s.a_sy = (int64_t)Win * C; s.a_sb = (int64_t)Hin * Win * C;
    d.N = d.Nrows = N; d.BN = N < 256 ? N : 256;
    d.epi = CHB_EPI_PLAIN; d.act = act; d.bias = bias;
    d.out = out; d.out_dtype = CHB_F32;
    d.o_sn = 1; d.o_sx = N; d.o_sy = (int64_t)W * N; d.o_sb = (int64_t)H * W * N;
    return build_conv_plan(d, plan);
  };
  plans.resize(4);
  int rc;
  // L2 at full resolution S x S (stride-2 phase taken later), L3 at S/2, L4 (zero-inserted) at S/2, L5 over the padded map
  if ((rc = mk(ws + z->ws_a1, S, S, 32, blob + z->tensors[z->t_w[0]].offset,
               reinterpret_cast<const float*>(blob + z->tensors[z->t_b[0]].offset), 64, S, S, 0, ws + z->ws_c2,
               CHB_ACT_NONE, &plans[0])) != CHB_OK) return rc;
  if ((rc = mk(ws + z->ws_a2, S2, S2, 64, blob + z->tensors[z->t_w[1]].offset,
               reinterpret_cast<const float*>(blob + z->tensors[z->t_b[1]].offset), 128, S2, S2, 0, ws + z->ws_c3,
               CHB_ACT_NONE, &plans[1])) != CHB_OK) return rc;
  if ((rc = mk(ws + z->ws_z3, S2, S2, 128, blob + z->tensors[z->t_w[2]].offset,
               reinterpret_cast<const float*>(blob + z->tensors[z->t_b[2]].offset), 256, S2, S2, 0, ws + z->ws_c4,
               CHB_ACT_NONE, &plans[2])) != CHB_OK) return rc;
  if ((rc = mk(ws + z->ws_a4, S2 + 2, S2 + 2, 256, blob + z->tensors[z->t_w[3]].offset,
               reinterpret_cast<const float*>(blob + z->tensors[z->t_b[3]].offset), 512, S2, S2, 1, ws + z->ws_codes,
               CHB_ACT_TANH, &plans[3])) != CHB_OK) return rc;
  return CHB_OK;
}
}  // namespace chb

extern "C" {

int chb_zencoder_create(const chb_zenc_config* cfg, chb_zencoder** out) {
  if (!cfg || !out || cfg->crop < 64 || (cfg->crop & (cfg->crop - 1)) || cfg->label_nc <= 0 || cfg->label_nc > 32 ||
      cfg->max_batch <= 0) {
    set_error("chb_zencoder_create: need crop a power of two >= 64, label_nc in 1..32, max_batch > 0");
    return CHB_ERR_ARG;
  }
  chb_zencoder* z = new chb_zencoder();
  z->cfg = *cfg;
  const int64_t B = cfg->max_batch, S = cfg->crop, S2 = S / 2, S4 = S / 4, NC = cfg->label_nc;
  z->t_w1 = zadd(z, "l1.w", 32 * 27 * 4, CHB_F32);
  z->t_b1 = zadd(z, "l1.b", 32 * 4, CHB_F32);
  const int cin[4] = {32, 64, 128, 256};
  const char* nm[4] = {"l2", "l3", "l4", "l5"};
  for (int i = 0; i < 4; ++i) {
    z->t_w[i] = zadd(z, (std::string(nm[i]) + ".w").c_str(), (int64_t)kZC[i + 1] * 9 * cin[i] * 2, CHB_F16);
    z->t_b[i] = zadd(z, (std::string(nm[i]) + ".b").c_str(), (int64_t)kZC[i + 1] * 4, CHB_F32);
  }
  z->ws_img = zws(z, B * 3 * S * S * 4);
  z->ws_labels = zws(z, B * S * S);
  z->ws_out = zws(z, B * NC * 512 * 4);
  z->ws_c1 = zws(z, B * S * S * 32 * 4);
  z->ws_a1 = zws(z, B * S * S * 32 * 2);
  z->ws_c2 = zws(z, B * S * S * 64 * 4);       // L2 computed at full resolution
  z->ws_a2 = zws(z, B * S2 * S2 * 64 * 2);
  z->ws_c3 = zws(z, B * S2 * S2 * 128 * 4);     // L3 computed at S/2
  z->ws_z3 = zws(z, B * S2 * S2 * 128 * 2);     // zero-inserted S/4 -> S/2
  z->ws_c4 = zws(z, B * S2 * S2 * 256 * 4);
  z->ws_a4 = zws(z, B * (S2 + 2) * (S2 + 2) * 256 * 2);
  z->ws_codes = zws(z, B * S2 * S2 * 512 * 4);
  z->ws_sums = zws(z, B * 256 * 2 * 8 + B * 256 * 8);   // double moments [B][256][2] + float2 {rstd, shift} [B][256]
  z->ws_psum = zws(z, B * NC * 512 * 4);
  z->ws_pcnt = zws(z, B * NC * 4);
  (void)S4;
  *out = z;
  return CHB_OK;
}

void chb_zencoder_destroy(chb_zencoder* z) { delete z; }
int chb_zencoder_num_tensors(const chb_zencoder* z) { return z ? (int)z->tensors.size() : 0; }
int chb_zencoder_tensor_info(const chb_zencoder* z, int i, char* name, int cap, int64_t* offset, int64_t* nbytes,
                             int* dtype) {
  if (!z || i < 0 || i >= (int)z->tensors.size()) {
    set_error("chb_zencoder_tensor_info: index out of range");
    return CHB_ERR_ARG;
  }
  const ZTensor& t = z->tensors[i];
  if (name && cap > 0) snprintf(name, cap, "%s", t.name.c_str());
  if (offset) *offset = t.offset;
  if (nbytes) *nbytes = t.nbytes;
  if (dtype) *dtype = t.dtype;
  return CHB_OK;
}
int64_t chb_zencoder_blob_bytes(const chb_zencoder* z) { return z ? z->blob_bytes : 0; }
int64_t chb_zencoder_workspace_bytes(const chb_zencoder* z) { return z ? z->ws_bytes : 0; }

int chb_zencoder_bind(chb_zencoder* z, const void* blob, void* workspace) {
  if (!z || !blob || !workspace || (reinterpret_cast<uintptr_t>(blob) & 255) ||
      (reinterpret_cast<uintptr_t>(workspace) & 1023)) {
    set_error("chb_zencoder_bind: NULL or misaligned pointers (blob 256 B, workspace 1024 B)");
    return CHB_ERR_ARG;
  }
  int rc = chb_check_device();
  if (rc != CHB_OK) return rc;
  z->blob = reinterpret_cast<const uint8_t*>(blob);
  z->ws = reinterpret_cast<uint8_t*>(workspace);
  z->plans.clear();
  return CHB_OK;
}

int chb_zencoder_forward(chb_zencoder* z, const float* img, const uint8_t* labels, float* out, int B, void* stream_) {
  if (!z || !img || !labels || !out || !z->ws || !z->blob) {
    set_error("chb_zencoder_forward: NULL argument or unbound encoder");
    return CHB_ERR_ARG;
  }
  if (B <= 0 || B > z->cfg.max_batch) {
    set_error("chb_zencoder_forward: batch exceeds max_batch");
    return CHB_ERR_ARG;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  auto it = z->plans.find(B);
  if (it == z->plans.end()) {
    std::vector<ConvPlan> pl;
    int rc = zencoder_plans(z, B, pl);
    if (rc != CHB_OK) return rc;
    it = z->plans.emplace(B, std::move(pl)).first;
  }
  const std::vector<ConvPlan>& pl = it->second;
  const int S = z->cfg.crop, S2 = S / 2, S4 = S / 4, NC = z->cfg.label_nc;
  uint8_t* ws = z->ws;
  double* sums = reinterpret_cast<double*>(ws + z->ws_sums);
  auto F = [&](int64_t o) { return reinterpret_cast<float*>(ws + o); };
  auto Hh = [&](int64_t o) { return reinterpret_cast<__half*>(ws + o); };
  auto norm = [&](const float* x, int H, int W, int C, int s, long long row_stride, long long img_stride, __half* y,
                  int mode) {
    cudaMemsetAsync(sums, 0, (size_t)B * C * 2 * sizeof(double), st);
    const int lanes = 256 / (C / 4);
    // enough blocks to fill the machine (B * slabs >= ~8 per SM), each with >= 64 pixels per lane where possible
    long long slabs = ((long long)H * W + lanes * 64 - 1) / (lanes * 64);
    const long long want = ((long long)device_sm_count() * 8 + B - 1) / B;
    if (slabs > want) slabs = want;
    if (slabs < 1) slabs = 1;
    in_stats_kernel<<<dim3((unsigned)slabs, (unsigned)B), 256, 0, st>>>(x, sums, H, W, C, s, row_stride, img_stride);
    float2* ss = reinterpret_cast<float2*>(sums + (size_t)B * 256 * 2);   // second half of the statistics buffer
    in_finalize_kernel<<<(B * C + 255) / 256, 256, 0, st>>>(sums, ss, B * C, 1.0 / ((double)H * W));
    const int OH = mode == 1 ? 2 * H : (mode == 2 ? H + 2 : H), OW = mode == 1 ? 2 * W : (mode == 2 ? W + 2 : W);
    in_apply_kernel<<<zgrid((long long)B * OH * OW * (C / 8), 256), 256, 0, st>>>(x, ss, y, B, H, W, C, s, row_stride,
                                                                                img_stride, mode);
  };
  // L1
  zenc_conv1_kernel<<<zgrid((long long)B * S * S, 128), 128, 0, st>>>(
      img, reinterpret_cast<const float*>(z->blob + z->tensors[z->t_w1].offset),
      reinterpret_cast<const float*>(z->blob + z->tensors[z->t_b1].offset), F(z->ws_c1), B, S);
  norm(F(z->ws_c1), S, S, 32, 1, (long long)S * 32, (long long)S * S * 32, Hh(z->ws_a1), 0);
  // L2 (stride 2: conv at S x S, keep the even phase)
  int rc = launch_conv_plan(pl[0], CHB_IMPL_TCGEN05, st);
  if (rc != CHB_OK) return rc;
  norm(F(z->ws_c2), S2, S2, 64, 2, (long long)S * 64, (long long)S * S * 64, Hh(z->ws_a2), 0);
  // L3 (stride 2 at S/2 -> S/4), output zero-inserted back to S/2 for the transposed conv
  if ((rc = launch_conv_plan(pl[1], CHB_IMPL_TCGEN05, st)) != CHB_OK) return rc;
  norm(F(z->ws_c3), S4, S4, 128, 2, (long long)S2 * 128, (long long)S2 * S2 * 128, Hh(z->ws_z3), 1);
  // L4 (ConvTranspose as conv over the zero-inserted map), output reflection-padded for L5
  if ((rc = launch_conv_plan(pl[2], CHB_IMPL_TCGEN05, st)) != CHB_OK) return rc;
  norm(F(z->ws_c4), S2, S2, 256, 1, (long long)S2 * 256, (long long)S2 * S2 * 256, Hh(z->ws_a4), 2);
  // L5 + tanh
  if ((rc = launch_conv_plan(pl[3], CHB_IMPL_TCGEN05, st)) != CHB_OK) return rc;
  // region pooling
  float* psum = F(z->ws_psum);
  int* pcnt = reinterpret_cast<int*>(ws + z->ws_pcnt);
  cudaMemsetAsync(psum, 0, (size_t)B * NC * 512 * 4, st);
  cudaMemsetAsync(pcnt, 0, (size_t)B * NC * 4, st);
  int sh = 0;
  while ((S2 << sh) < S) ++sh;
  const int ppb = 256;
  const int slabs = (S2 * S2 + ppb - 1) / ppb;
  const size_t smem = (size_t)NC * 512 * 4;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(region_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 512 * 4);
    attr_set = true;
  }
  region_pool_kernel<<<dim3((unsigned)slabs, (unsigned)B), 128, smem, st>>>(F(z->ws_codes), labels, psum, pcnt, S2, 512,
                                                                           S, sh, NC, ppb);
  region_finalize_kernel<<<zgrid((long long)B * NC * 512, 256), 256, 0, st>>>(psum, pcnt, out, (long long)B * NC * 512,
                                                                             512);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    set_error(std::string("zencoder launch failed: ") + cudaGetErrorString(err));
    return CHB_ERR_CUDA;
  }
  return CHB_OK;
}

int chb_zencoder_forward_host(chb_zencoder* z, const float* img_host, const uint8_t* labels_host, float* out_host, int B,
                              void* stream_) {
  if (!z || !img_host || !labels_host || !out_host || !z->ws) {
    set_error("chb_zencoder_forward_host: NULL argument or unbound encoder");
    return CHB_ERR_ARG;
  }
  if (B <= 0 || B > z->cfg.max_batch) {
    set_error("chb_zencoder_forward_host: batch exceeds max_batch");
    return CHB_ERR_ARG;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  const size_t S2 = (size_t)z->cfg.crop * z->cfg.crop;
  cudaError_t err = cudaMemcpyAsync(z->ws + z->ws_img, img_host, (size_t)B * 3 * S2 * 4, cudaMemcpyHostToDevice, st);
  if (err == cudaSuccess)
    err = cudaMemcpyAsync(z->ws + z->ws_labels, labels_host, (size_t)B * S2, cudaMemcpyHostToDevice, st);
  if (err != cudaSuccess) {
    set_error(std::string("H2D copy failed: ") + cudaGetErrorString(err));
    return CHB_ERR_CUDA;
  }
  float* dout = reinterpret_cast<float*>(z->ws + z->ws_out);
  int rc = chb_zencoder_forward(z, reinterpret_cast<const float*>(z->ws + z->ws_img), z->ws + z->ws_labels, dout, B,
                                stream_);
  if (rc != CHB_OK) return rc;
  err = cudaMemcpyAsync(out_host, dout, (size_t)B * z->cfg.label_nc * 512 * 4, cudaMemcpyDeviceToHost, st);
  if (err == cudaSuccess) err = cudaStreamSynchronize(st);
  if (err != cudaSuccess) {
    set_error(std::string("D2H copy / sync failed: ") + cudaGetErrorString(err));
    return CHB_ERR_CUDA;
  }
  return CHB_OK;
}

}  // extern "C"
