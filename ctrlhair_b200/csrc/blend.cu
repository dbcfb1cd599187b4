// Post-processing either side of the generator, device resident:
//   * Poisson blending of the generated image into the input face (poisson_blending.py:29-87, hair_editor.py:257-308):
//     blend mask (hair union + elliptical dilations), the sparse system solved by conjugate gradients in fp64 with the
//     whole state of one (image, channel) system resident in the registers and shared memory of an 8- or 16-CTA
//     thread-block cluster (halo rows and the dot products of an iteration travel through distributed shared memory,
//     no HBM traffic inside the loop).  Two kernels: poisson_cg2_kernel for the reference's 256-column images,
//     poisson_cg_kernel (first generation, classic CG, 1024 threads) for every other size;
//   * 8-bit RGB <-> HSV (cv2.cvtColor at ui/backend.py:98-101,108-125) and label map <-> one-hot
//     (shape_branch/shape_util.py:6-20), which remove the host round trips of Backend.parse_img / output.
#include <cooperative_groups.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <mutex>
#include <string>

#include "../../include/ctrlhair_b200.h"
#include "conv_igemm.cuh"
#include "ptx_sm100.cuh"

namespace cg = cooperative_groups;

namespace chb {

namespace {

int blend_check(const char* what) {
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    set_error(std::string(what) + ": " + cudaGetErrorString(err));
    return CHB_ERR_CUDA;
  }
  return CHB_OK;
}

int blend_grid(long long n, int threads) {
  long long g = (n + threads - 1) / threads;
  const long long cap = (long long)device_sm_count() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

constexpr int kHair = 13;        // global_value_utils.py:49-52
constexpr int kBackground = 0;   // PARSING_LABEL_LIST.index('background')

// ndarray.astype('uint8') of an in-range float: truncation toward zero (x86 numpy wraps out-of-range values modulo 256)
__device__ __forceinline__ uint8_t f32_to_u8_trunc(float v) { return (uint8_t)(((int)truncf(v)) & 0xFF); }

}  // namespace

// ------------------------------------------------------------------------------------------------------------------
// hair_editor.py:273-288: generator output [B,3,H,W] in [-1,1] -> cv2 layout uint8 [B,H,W,3], (x*127.5+127.5).astype(uint8)
__global__ void image_to_u8_kernel(const float* __restrict__ img, uint8_t* __restrict__ out, int B, int H, int W) {
  const long long hw = (long long)H * W, total = (long long)B * hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / hw, pix = i - b * hw;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = img[(b * 3 + c) * hw + pix];
      out[i * 3 + c] = f32_to_u8_trunc(__fadd_rn(__fmul_rn(v, 127.5f), 127.5f));  // numpy: two roundings, no FMA
    }
  }
}

// hair_editor.py:297-306.  cv2.getStructuringElement(MORPH_ELLIPSE): half-width of each row of the 13x13 / 5x5 element.
__constant__ int kEll13[13] = {0, 3, 4, 5, 6, 6, 6, 6, 6, 5, 4, 3, 0};
__constant__ int kEll5[5] = {0, 2, 2, 2, 0};

__global__ void blend_mask_kernel(const uint8_t* __restrict__ target_parsing, const uint8_t* __restrict__ face_parsing,
                                  uint8_t* __restrict__ res_mask_dilated, uint8_t* __restrict__ solve_mask, int B, int H,
                                  int W) {
  const long long hw = (long long)H * W, total = (long long)B * hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / hw;
    const int pix = (int)(i - b * hw), y = pix / W, x = pix - y * W;
    const uint8_t* tp = target_parsing + b * hw;
    const uint8_t* fp = face_parsing + b * hw;
    const bool bg = tp[pix] == kBackground;
    const int k = bg ? 5 : 13, r = k >> 1;
    int hit = 0;
    for (int dy = -r; dy <= r && !hit; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= H) continue;  // pixels outside the image never win a dilation
      const int hwid = bg ? kEll5[dy + r] : kEll13[dy + r];
      const int x0 = max(0, x - hwid), x1 = min(W - 1, x + hwid);
      for (int xx = x0; xx <= x1; ++xx) {
        if (tp[yy * W + xx] == kHair || fp[yy * W + xx] == kHair) { hit = 1; break; }
      }
    }
    if (res_mask_dilated) res_mask_dilated[i] = (uint8_t)hit;
    if (solve_mask) solve_mask[i] = (uint8_t)(1 - hit);  // the `1 - res_mask_dilated` handed to poisson_blending
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Poisson solve.  The reference assembles (poisson_blending.py:45-76), per channel, the H*W x H*W system
//   row k = identity             for interior pixels with mask == 0            (f = target)
//   row k = 5-point Laplacian    for every other pixel (mask != 0, or on the image border), truncated at the border
//   b[k]  = Laplacian(source)[k] where mask != 0, target[k] where mask == 0
// and calls spsolve.  Eliminating the identity rows leaves a symmetric positive definite system on the set U of
// Laplacian-row pixels:  4 f_i - sum_{j in N(i), j in U} f_j = b_i + sum_{j in N(i), j not in U} target_j,
// which is solved here by conjugate gradients in fp64.
//
// First-generation kernel (general sizes: W <= 512, ceil(H/8) * W <= 8192): classic CG, 8 pixels per thread addressed
// through per-pixel flags, x / p in shared memory, three cluster barriers and two reductions per iteration.
constexpr int kPoiCluster = 8;     // CTAs per system (portable cluster size)
constexpr int kPoiThreads = 1024;
constexpr int kPoiPx = 8;          // pixels per thread -> up to 8 * 1024 * 8 = 65536 pixels per system

struct PoissonParams {
  const uint8_t* source;   // [B,H,W,3]
  const uint8_t* target;   // [B,H,W,3]
  const uint8_t* mask;     // [B,H,W], non-zero = solve
  uint8_t* out;            // [B,H,W,3]
  float* stats;            // [B*3][2] = iterations, final relative residual; may be null
  const double* lut_fwd;   // [256] v ** (1/2.2) as the caller's host computes it, or null (device pow)
  const uint8_t* lut_known;// [256] uint8((v ** (1/2.2)) ** 2.2) on the caller's host, or null
  const __half* coarse_inv;// preconditioned kernel: [B][256][256] inverse coarse operators (poisson_coarse_inverse_kernel)
  int B, H, W, R;          // R = rows per CTA
  int with_gamma, max_iter;
  double tol2;             // squared relative residual target
};

struct PoissonSmem {
  double lut[256];                 // v ** (1/gamma)
  double warp_part[32];
  double slots[2][kPoiCluster];    // per-CTA partial sums of the two alternating cluster reductions
};

__device__ __forceinline__ double poisson_cluster_sum(cg::cluster_group& cluster, double v, PoissonSmem* sm, int set,
                                                      unsigned rank) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if (lane == 0) sm->warp_part[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double s = sm->warp_part[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    s = __shfl_sync(0xffffffffu, s, 0);
    if (lane < kPoiCluster) {
      double* remote = cluster.map_shared_rank(&sm->slots[set][0], lane);   // CTA `lane` of the cluster
      remote[rank] = s;
    }
  }
  cluster.sync();  // release/acquire across the cluster: every CTA now holds all eight partials
  double tot = 0.0;
#pragma unroll
  for (int k = 0; k < kPoiCluster; ++k) tot += sm->slots[set][k];  // same order in every CTA: bitwise identical sums
  return tot;
}

__global__ void __cluster_dims__(kPoiCluster, 1, 1) __launch_bounds__(kPoiThreads, 1)
    poisson_cg_kernel(const PoissonParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  const int sys = blockIdx.x / kPoiCluster;     // (image, channel)
  const int b = sys / 3, ch = sys - b * 3;
  const int H = p.H, W = p.W, R = p.R;
  const int row0 = (int)rank * R;
  const int n_own = max(0, min(R, H - row0));   // rows of this CTA
  const int n_next = max(0, min(R, H - (row0 + R)));

  PoissonSmem* sm = reinterpret_cast<PoissonSmem*>(smem_raw);
  double* pbuf = reinterpret_cast<double*>(smem_raw + sizeof(PoissonSmem));  // [(R + 2) rows][W]: search direction + halos
  double* xbuf = pbuf + (size_t)(R + 2) * W;                                  // [R rows][W]: iterate

  const int tid = threadIdx.x;
  for (int i = tid; i < (R + 2) * W; i += kPoiThreads) pbuf[i] = 0.0;
  if (tid < 256)
    sm->lut[tid] = p.with_gamma ? (p.lut_fwd ? p.lut_fwd[tid] : pow((double)tid, 1.0 / 2.2)) : (double)tid;
  if (tid < 2 * kPoiCluster) sm->slots[tid / kPoiCluster][tid % kPoiCluster] = 0.0;
  cluster.sync();  // nobody pushes a halo row into a buffer that is still being cleared

  const long long img_off = (long long)b * H * W;
  const uint8_t* src = p.source + img_off * 3 + ch;
  const uint8_t* tgt = p.target + img_off * 3 + ch;
  const uint8_t* msk = p.mask + img_off;

  // per-pixel flags: bit 0 = in U, bit 1 = has a left neighbour, bit 2 = has a right neighbour
  unsigned flags = 0;
  double r[kPoiPx], q[kPoiPx];
  double bb_local = 0.0;
#pragma unroll
  for (int j = 0; j < kPoiPx; ++j) {
    r[j] = 0.0;
    const int li = tid + j * kPoiThreads;
    if (li >= n_own * W) continue;
    const int lr = li / W, x = li - lr * W, y = row0 + lr;
    const int g = y * W + x;
    const bool m = msk[g] != 0;
    const bool border = (y == 0) | (y == H - 1) | (x == 0) | (x == W - 1);
    const bool inU = m | border;
    const double s = sm->lut[src[3 * g]], t = sm->lut[tgt[3 * g]];
    double rhs = m ? 4.0 * s : t;
    // neighbours inside the image: source Laplacian where mask != 0; known targets move to the right-hand side
    const int ny[4] = {y - 1, y + 1, y, y};
    const int nx[4] = {x, x, x - 1, x + 1};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int yy = ny[k], xx = nx[k];
      if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
      const int gn = yy * W + xx;
      if (m) rhs -= sm->lut[src[3 * gn]];
      const bool nborder = (yy == 0) | (yy == H - 1) | (xx == 0) | (xx == W - 1);
      if (!(nborder || msk[gn] != 0)) rhs += sm->lut[tgt[3 * gn]];
    }
    unsigned f = 0;
    if (inU) f |= 1u;
    if (x > 0) f |= 2u;
    if (x < W - 1) f |= 4u;
    flags |= f << (3 * j);
    if (inU) {
      r[j] = rhs;
      bb_local += rhs * rhs;
      const double x0 = m ? s : t;  // initial guess: the source where its gradients are kept, the target elsewhere
      pbuf[(lr + 1) * W + x] = x0;
      xbuf[li] = x0;
    } else {
      xbuf[li] = t;                 // known pixel: f = target (pbuf stays 0 there)
    }
  }
  __syncthreads();

  auto push_halos = [&]() {
    // first owned row -> bottom halo (row R + 1) of the CTA above; last owned row -> top halo (row 0) of the CTA below
    if (n_own > 0) {
      if (rank > 0 && tid < W) {
        double* remote = cluster.map_shared_rank(pbuf, rank - 1);
        remote[(R + 1) * W + tid] = pbuf[W + tid];
      }
      if (n_next > 0 && tid >= 512 && tid - 512 < W) {
        double* remote = cluster.map_shared_rank(pbuf, rank + 1);
        remote[tid - 512] = pbuf[n_own * W + (tid - 512)];
      }
    }
  };
  auto apply = [&](int j) -> double {  // (A p)_i for pixel j of this thread (p == 0 outside U and outside the image)
    const unsigned f = (flags >> (3 * j)) & 7u;
    if (!(f & 1u)) return 0.0;
    const int li = tid + j * kPoiThreads;
    const int idx = li + W;  // (lr + 1) * W + x
    double v = 4.0 * pbuf[idx] - pbuf[idx - W] - pbuf[idx + W];
    if (f & 2u) v -= pbuf[idx - 1];
    if (f & 4u) v -= pbuf[idx + 1];
    return v;
  };

  push_halos();
  const double bb = poisson_cluster_sum(cluster, bb_local, sm, 0, rank);  // its cluster.sync also publishes the halos
  // r = rhs - A x0
  double rr_local = 0.0;
#pragma unroll
  for (int j = 0; j < kPoiPx; ++j) {
    r[j] -= apply(j);
    rr_local += r[j] * r[j];
  }
  cluster.sync();  // every CTA has finished reading x0 (own rows and halos) before p = r overwrites it
#pragma unroll
  for (int j = 0; j < kPoiPx; ++j)
    if ((flags >> (3 * j)) & 1u) pbuf[tid + j * kPoiThreads + W] = r[j];
  __syncthreads();
  push_halos();
  double rr = poisson_cluster_sum(cluster, rr_local, sm, 1, rank);
  const double thresh = p.tol2 * bb;

  int it = 0;
  while (it < p.max_iter && rr > thresh) {
    double pq = 0.0;
#pragma unroll
    for (int j = 0; j < kPoiPx; ++j) {
      q[j] = apply(j);
      if ((flags >> (3 * j)) & 1u) pq += q[j] * pbuf[tid + j * kPoiThreads + W];
    }
    const double pAp = poisson_cluster_sum(cluster, pq, sm, 0, rank);
    const double alpha = rr / pAp;
    double rn_local = 0.0;
#pragma unroll
    for (int j = 0; j < kPoiPx; ++j) {
      if ((flags >> (3 * j)) & 1u) {
        const int li = tid + j * kPoiThreads;
        xbuf[li] += alpha * pbuf[li + W];
        r[j] -= alpha * q[j];
        rn_local += r[j] * r[j];
      }
    }
    const double rn = poisson_cluster_sum(cluster, rn_local, sm, 1, rank);
    ++it;
    const double beta = rn / rr;
    rr = rn;
    if (!(rr > thresh) || it >= p.max_iter) break;  // identical in every CTA of the cluster (same sums, same order)
#pragma unroll
    for (int j = 0; j < kPoiPx; ++j) {
      if ((flags >> (3 * j)) & 1u) {
        const int idx = tid + j * kPoiThreads + W;
        pbuf[idx] = r[j] + beta * pbuf[idx];
      }
    }
    __syncthreads();
    push_halos();
    cluster.sync();
  }

  // poisson_blending.py:81-86: back through the gamma curve, clip to [0, 255], truncate to uint8
  uint8_t* out = p.out + img_off * 3 + ch;
#pragma unroll
  for (int j = 0; j < kPoiPx; ++j) {
    const int li = tid + j * kPoiThreads;
    if (li >= n_own * W) continue;
    const int g = row0 * W + li;
    if (p.with_gamma && p.lut_known && !((flags >> (3 * j)) & 1u)) {
      // f == target here: pow(pow(v, 1/2.2), 2.2) lands within an ulp of the integer v and truncates to v or v - 1
      // depending on the host's pow; the caller's table reproduces its own host exactly
      out[3 * g] = p.lut_known[tgt[3 * g]];
      continue;
    }
    double v = xbuf[li];
    if (p.with_gamma) v = v > 0.0 ? pow(v, 2.2) : 0.0;  // numpy: negative ** 2.2 = nan -> 0 after astype
    v = v > 255.0 ? 255.0 : (v < 0.0 ? 0.0 : v);
    out[3 * g] = (uint8_t)(int)v;
  }
  if (p.stats && rank == 0 && tid == 0) {
    p.stats[2 * sys] = (float)it;
    p.stats[2 * sys + 1] = bb > 0.0 ? (float)sqrt(rr / bb) : 0.0f;
  }
  cluster.sync();  // no CTA leaves while a peer may still address its shared memory
}

// ------------------------------------------------------------------------------------------------------------------
// Second generation of the solver (the default when the image fits it): Chronopoulos-Gear conjugate gradients, whose two
// dot products (r.r and r.Ar) come out of ONE cluster reduction per iteration, with the per-pixel state in registers.
//   * 512 threads (256 in the 16-CTA shape); a thread owns a 16-row strip of one column, so the vertical neighbours of the 5-point stencil are
//     its own registers and only left / right (and the two strip ends) are read from shared memory;
//   * r, p = search direction and s = A p live in registers (96 of the 128 available), x and w = A r in shared memory
//     (touched once per iteration each), r additionally in a haloed shared buffer for the neighbours;
//   * per iteration: one halo exchange (split cluster barrier, overlapped with the interior rows) + one two-value
//     reduction signalled through transaction barriers (st.async into the peers' shared memory, no cluster
//     rendezvous), one block reduction instead of 2, and 7 shared-memory accesses per pixel instead of 11.
// x_{i+1} = x_i + a_i p_i, r_{i+1} = r_i - a_i s_i, w = A r_{i+1}, g = (r,r), d = (r,w), b = g/g_old,
// a = g / (d - b g / a_old), p = r + b p, s = w + b s.
constexpr int kPoi2MaxCluster = 16;
constexpr int kPoi2Rows = 16;   // rows per thread

constexpr int kPoi2Agg = 16;    // preconditioner: 16 x 16-pixel aggregates (a thread's 16-row strip lies in one)
constexpr int kPoi2Coarse = 256;   // coarse unknowns of a 256 x 256 image
constexpr float kPoi2InvScale = 256.f;   // the fp16 copy of A_c^-1 (entries <~ 1) is stored times 256: keeps the small
                                         // far-field entries out of the subnormal range

struct Poisson2Smem {
  double lut[256];
  double warp_part[3][16];
  double slots[2][3][kPoi2MaxCluster];   // [parity][value][CTA]
  uint64_t red_bar[2];                   // one transaction barrier per parity: the bytes of one exchange per phase
};

// Preconditioned variant only (placed behind the three pixel buffers).
struct Poisson2Coarse {
  double gathered[2][kPoi2Coarse];   // [parity][aggregate]: P^T w of the whole image, every CTA receives all of it
  double rc[kPoi2Coarse];            // P^T r, kept by recurrence in every CTA (fp64: it decays with r by 1e-11)
  double sc[kPoi2Coarse];            // P^T s
  float rc32[kPoi2Coarse];           // rc once more for the fp32 matvec
  float ec[32];                      // (A_c^-1 rc) of this CTA's aggregates
};

// The address of `saddr` (a shared::cta address of this CTA) in CTA `cta` of the cluster.
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(cta));
  return r;
}
// 8-byte store into a peer CTA's shared memory that signals 8 bytes on that CTA's transaction barrier when it lands.
__device__ __forceinline__ void st_async_f64(uint32_t remote_addr, double v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(remote_addr),
               "l"(__double_as_longlong(v)), "r"(remote_bar)
               : "memory");
}

__device__ __forceinline__ void cluster_arrive_release() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acquire() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Cluster-wide sum of two values per iteration without a cluster barrier: every CTA sends its two partials to every
// CTA of the cluster with st.async, each store completing 8 bytes on the RECEIVER's transaction barrier; a CTA waits on
// its own barrier only (CL = CTAs per cluster, NW = warps per CTA <= 16).
// No cluster-wide rendezvous and no fence: the latency is one DSMEM store.  Two barriers / slot sets alternate; a CTA can
// only be one reduction ahead of any other (it needs everyone's partials to finish the current one), so a set is
// never overwritten while somebody still reads it.
template <int CL, int NW>
__device__ __forceinline__ void poisson_cluster_sum2_async(double& a, double& b, Poisson2Smem* sm, int set,
                                                           uint32_t phase, unsigned rank) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_down_sync(0xffffffffu, a, o);
    b += __shfl_down_sync(0xffffffffu, b, o);
  }
  if (lane == 0) { sm->warp_part[0][warp] = a; sm->warp_part[1][warp] = b; }
  __syncthreads();
  if (warp == 0) {
    double v = (lane & 15) < NW ? sm->warp_part[lane >> 4][lane & 15] : 0.0;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o, 16);
    const double v0 = __shfl_sync(0xffffffffu, v, 0), v1 = __shfl_sync(0xffffffffu, v, 16);
    if (lane == 0) mbar_arrive_expect_tx(&sm->red_bar[set], 16u * CL);
    if (lane < 2 * CL) {
      const int which = lane / CL, dst = lane % CL;
      st_async_f64(mapa_u32(smem_u32(&sm->slots[set][which][rank]), (uint32_t)dst), which ? v1 : v0,
                   mapa_u32(smem_u32(&sm->red_bar[set]), (uint32_t)dst));
    }
  }
  mbar_wait(&sm->red_bar[set], phase);
  double ta = 0.0, tb = 0.0;
#pragma unroll
  for (int k = 0; k < CL; ++k) { ta += sm->slots[set][0][k]; tb += sm->slots[set][1][k]; }
  a = ta; b = tb;
}

// The exchange of a preconditioned iteration: three cluster-wide sums AND the all-gather of the coarse restriction of
// one vector (`wsum` = this thread's 16-row column sum; the 16 lanes that share an aggregate add up, then send the
// aggregate's sum to every CTA) on the same transaction barrier: 8 * (3 CL + 256) bytes land per phase.
template <int CL, int NW>
__device__ __forceinline__ void poisson_cluster_exchange(double& a, double& b, double& c, double wsum, int agg,
                                                         Poisson2Smem* sm, Poisson2Coarse* co, int set, uint32_t phase,
                                                         unsigned rank) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t bar = smem_u32(&sm->red_bar[set]);
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
  if ((lane & 15) < CL)
    st_async_f64(mapa_u32(smem_u32(&co->gathered[set][agg]), (uint32_t)(lane & 15)), wsum,
                 mapa_u32(bar, (uint32_t)(lane & 15)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_down_sync(0xffffffffu, a, o);
    b += __shfl_down_sync(0xffffffffu, b, o);
    c += __shfl_down_sync(0xffffffffu, c, o);
  }
  if (lane == 0) { sm->warp_part[0][warp] = a; sm->warp_part[1][warp] = b; sm->warp_part[2][warp] = c; }
  __syncthreads();
  if (warp == 0) {
    double v = (lane & 15) < NW ? sm->warp_part[lane >> 4][lane & 15] : 0.0;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o, 16);
    const double v0 = __shfl_sync(0xffffffffu, v, 0), v1 = __shfl_sync(0xffffffffu, v, 16);
    if (lane == 0) mbar_arrive_expect_tx(&sm->red_bar[set], 8u * (3u * CL + kPoi2Coarse));
    if (lane < 2 * CL) {
      const int which = lane / CL, dst = lane % CL;
      st_async_f64(mapa_u32(smem_u32(&sm->slots[set][which][rank]), (uint32_t)dst), which ? v1 : v0,
                   mapa_u32(bar, (uint32_t)dst));
    }
  } else if (warp == 1) {
    double v = lane < NW ? sm->warp_part[2][lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    v = __shfl_sync(0xffffffffu, v, 0);
    if (lane < CL)
      st_async_f64(mapa_u32(smem_u32(&sm->slots[set][2][rank]), (uint32_t)lane), v, mapa_u32(bar, (uint32_t)lane));
  }
  mbar_wait(&sm->red_bar[set], phase);
  double ta = 0.0, tb = 0.0, tc = 0.0;
#pragma unroll
  for (int k = 0; k < CL; ++k) { ta += sm->slots[set][0][k]; tb += sm->slots[set][1][k]; tc += sm->slots[set][2][k]; }
  a = ta; b = tb; c = tc;
}

// PRE: two-level additive preconditioner M^-1 = D^-1 + P A_c^-1 P^T (Jacobi + a Galerkin coarse correction on 16 x 16
// pixel aggregates; P = piecewise constant on the aggregate's pixels of U).  A_c = P^T A P (256 x 256, one per IMAGE:
// the three channels share U) is inverted once by poisson_coarse_inverse_kernel; a CTA keeps the fp16 rows of its own
// aggregates in shared memory.  ~650 -> ~150 iterations at the same stopping rule; the extra per-iteration work is a
// 16-MAC-per-thread matvec, and P^T r is maintained by recurrence from P^T w, whose all-gather rides on the dot-product
// exchange, so an iteration still has one halo barrier and one exchange.
//   u = M^-1 r, w = A u, g = (r,u), d = (w,u), n = (r,r); b = g/g_old, a = g/(d - b g/a_old);
//   p = u + b p, s = w + b s, x += a p, r -= a s;   P^T s = P^T w + b P^T s, P^T r -= a P^T s.
template <int W, int R, int CL, bool PRE>   // columns, rows per CTA, CTAs per cluster (shape from the launch attribute)
__global__ void __launch_bounds__((R / kPoi2Rows) * W, PRE ? 1 : 512 / ((R / kPoi2Rows) * W))
    poisson_cg2_kernel(const PoissonParams p) {   // (PRE: one CTA per SM by shared memory, so 256 threads get 255 registers)
  constexpr int kPoi2Threads = (R / kPoi2Rows) * W;
  // compile-time geometry (the reference's 256 x 256 images: 8 CTAs x 32 rows): every shared-memory offset below is
  // an immediate, which is what lets r, p and s stay in registers.  The inner loop is branch free: pixels outside U
  // (and rows beyond the image in a partial CTA) keep r = p = s = 0 because their stencil value is selected to 0,
  // and the haloed buffer has a zero column on either side so that column 0 / W-1 need no special case.
  static_assert(R % kPoi2Rows == 0 && kPoi2Threads <= 512 && CL <= kPoi2MaxCluster, "one thread per (16-row strip, column)");
  constexpr int P = W + 2;   // pitch of the haloed residual buffer
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  const int sys = blockIdx.x / CL;
  const int b = sys / 3, ch = sys - b * 3;
  const int H = p.H;
  const int row0 = (int)rank * R;
  const int n_own = max(0, min(R, H - row0));
  const int n_next = max(0, min(R, H - (row0 + R)));

  Poisson2Smem* sm = reinterpret_cast<Poisson2Smem*>(smem_raw);
  double* rbuf = reinterpret_cast<double*>(smem_raw + sizeof(Poisson2Smem));  // [(R + 2)][P] residual + halo ring
  double* xbuf = rbuf + (R + 2) * P;                                           // [R][W] iterate
  double* wbuf = xbuf + R * W;                                                 // [R][W] A r  (PRE: A u)
  Poisson2Coarse* co = reinterpret_cast<Poisson2Coarse*>(wbuf + R * W);        // PRE only
  __half* ainv = reinterpret_cast<__half*>(co + 1);                            // PRE only: [R][256] rows of A_c^-1

  const int tid = threadIdx.x;
  for (int i = tid; i < (R + 2) * P; i += kPoi2Threads) rbuf[i] = 0.0;
  if (tid < 256)
    sm->lut[tid] = p.with_gamma ? (p.lut_fwd ? p.lut_fwd[tid] : pow((double)tid, 1.0 / 2.2)) : (double)tid;
  if (tid < 6 * kPoi2MaxCluster) (&sm->slots[0][0][0])[tid] = 0.0;
  if constexpr (PRE) {
    // rows [rank * R, rank * R + R) of this image's coarse inverse (fp16, 512 B per row, written by the launch before)
    const uint4* g = reinterpret_cast<const uint4*>(p.coarse_inv + ((size_t)(blockIdx.x / CL / 3) * kPoi2Coarse +
                                                                     (size_t)rank * R) * kPoi2Coarse);
    for (int i = tid; i < R * kPoi2Coarse / 8; i += kPoi2Threads) reinterpret_cast<uint4*>(ainv)[i] = g[i];
  }
  if (tid == 0) {
    mbar_init(&sm->red_bar[0], 1);
    mbar_init(&sm->red_bar[1], 1);
    fence_mbar_init();
  }
  cluster.sync();

  const long long img_off = (long long)b * H * W;
  const uint8_t* src = p.source + img_off * 3 + ch;
  const uint8_t* tgt = p.target + img_off * 3 + ch;
  const uint8_t* msk = p.mask + img_off;

  // this thread: column `col`, local rows lr0 .. lr0 + 15
  const int strip = tid / W, col = tid - strip * W;
  const int lr0 = strip * kPoi2Rows;
  unsigned umask = 0, vmask = 0;              // bit j: pixel j is in U / exists
  double bb_local = 0.0;
  // assembly of the right-hand side (parked in wbuf), the initial guess (rbuf, xbuf) and the masks: a rolled loop
#pragma unroll 1
  for (int j = 0; j < kPoi2Rows; ++j) {
    const int lr = lr0 + j;
    double x0 = 0.0, rhs = 0.0;
    if (lr < n_own) {
      vmask |= 1u << j;
      const int x = col, y = row0 + lr;
      const int g = y * W + x;
      const bool m = msk[g] != 0;
      const bool border = (y == 0) | (y == H - 1) | (x == 0) | (x == W - 1);
      const bool inU = m | border;
      const double sv = sm->lut[src[3 * g]], tv = sm->lut[tgt[3 * g]];
      rhs = m ? 4.0 * sv : tv;
#pragma unroll 1
      for (int k = 0; k < 4; ++k) {
        const int yy = y + (k == 0 ? -1 : (k == 1 ? 1 : 0)), xx = x + (k == 2 ? -1 : (k == 3 ? 1 : 0));
        if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
        const int gn = yy * W + xx;
        if (m) rhs -= sm->lut[src[3 * gn]];
        const bool nborder = (yy == 0) | (yy == H - 1) | (xx == 0) | (xx == W - 1);
        if (!(nborder || msk[gn] != 0)) rhs += sm->lut[tgt[3 * gn]];
      }
      x0 = tv;
      if (inU) {
        umask |= 1u << j;
        bb_local += rhs * rhs;
        x0 = m ? sv : tv;
        rbuf[(lr + 1) * P + 1 + col] = x0;
      } else {
        rhs = 0.0;
      }
    }
    xbuf[lr * W + col] = x0;
    wbuf[lr * W + col] = rhs;
  }

  double* rc = rbuf + (lr0 + 1) * P + 1 + col;   // this thread's first pixel in the haloed buffer
  double* xc = xbuf + lr0 * W + col;
  double* wc = wbuf + lr0 * W + col;
  // boundary rows travel straight from the owning thread into the neighbour CTA's halo row
  double* halo_up = (strip == 0 && rank > 0 && n_own > 0)
                        ? cluster.map_shared_rank(rbuf, rank - 1) + (R + 1) * P + 1 + col : nullptr;
  double* halo_dn = (strip == R / kPoi2Rows - 1 && n_next > 0)
                        ? cluster.map_shared_rank(rbuf, rank + 1) + 1 + col : nullptr;

  if (halo_up) *halo_up = rc[0];
  if (halo_dn) *halo_dn = rc[(kPoi2Rows - 1) * P];
  cluster.sync();
  // r = rhs - A x0 (all five points from shared memory: rbuf holds x0 here)
  double r[kPoi2Rows], pd[kPoi2Rows], s[kPoi2Rows];
#pragma unroll
  for (int j = 0; j < kPoi2Rows; ++j) {
    pd[j] = 0.0; s[j] = 0.0;
    const double v = 4.0 * rc[j * P] - rc[(j - 1) * P] - rc[(j + 1) * P] - rc[j * P - 1] - rc[j * P + 1];
    r[j] = ((umask >> j) & 1u) ? wc[j * W] - v : 0.0;
  }
  double zero = 0.0;
  uint32_t red_phases = 0u;   // bit `set`: the phase parity of that set's transaction barrier
  const int agg = (int)rank * R + strip * kPoi2Agg + (col >> 4);   // this thread's aggregate (image-wide index)
  if constexpr (PRE) {
    double rsum = 0.0, zero2 = 0.0;
#pragma unroll
    for (int j = 0; j < kPoi2Rows; ++j) rsum += r[j];
    poisson_cluster_exchange<CL, kPoi2Threads / 32>(bb_local, zero, zero2, rsum, agg, sm, co, 0, red_phases & 1u, rank);
    if (tid < kPoi2Coarse) {
      const double v = co->gathered[0][tid];
      co->rc[tid] = v; co->sc[tid] = 0.0; co->rc32[tid] = (float)v;
    }
  } else {
    poisson_cluster_sum2_async<CL, kPoi2Threads / 32>(bb_local, zero, sm, 0, red_phases & 1u, rank);
  }
  red_phases ^= 1u;
  cluster.sync();   // every CTA is done reading x0 from rbuf (own rows and halos) before r overwrites it
  const double bb = bb_local;
  const double thresh = p.tol2 * bb;

  int it = 0, parity = 1;
  double gamma = 1.0, alpha = 1.0;
  bool first = true;
  if constexpr (PRE) {
    double rr = bb;   // (r, r) of the last exchange, for the statistics
    while (true) {
      // e = (A_c^-1 P^T r) of this thread's aggregate: 16 threads per row of the CTA's slice, 16 columns each
      {
        const uint4* arow = reinterpret_cast<const uint4*>(ainv + (tid >> 4) * kPoi2Coarse + (tid & 15) * 16);
        const uint4 q0 = arow[0], q1 = arow[1];
        const float* rv = co->rc32 + (tid & 15);
        float acc0 = 0.f, acc1 = 0.f;
        auto mac2 = [&](uint32_t packed, int k) {   // two fp16 entries: columns (tid & 15) + 16 * (2k), + 16 * (2k + 1)
          acc0 = fmaf(__half2float(__ushort_as_half((unsigned short)(packed & 0xffffu))), rv[16 * (2 * k)], acc0);
          acc1 = fmaf(__half2float(__ushort_as_half((unsigned short)(packed >> 16))), rv[16 * (2 * k + 1)], acc1);
        };
        mac2(q0.x, 0); mac2(q0.y, 1); mac2(q0.z, 2); mac2(q0.w, 3);
        mac2(q1.x, 4); mac2(q1.y, 5); mac2(q1.z, 6); mac2(q1.w, 7);
        float acc = (acc0 + acc1) * (1.f / kPoi2InvScale);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ((tid & 15) == 0) co->ec[tid >> 4] = acc;
      }
      __syncthreads();
      const double e = (double)co->ec[strip * kPoi2Agg + (col >> 4)];
      auto uval = [&](int j) { return ((umask >> j) & 1u) ? fma(0.25, r[j], e) : 0.0; };
      // publish u: own pixels, and the strip-end rows into the neighbours' halos
#pragma unroll
      for (int j = 0; j < kPoi2Rows; ++j) rc[j * P] = uval(j);
      if (halo_up) *halo_up = uval(0);
      if (halo_dn) *halo_dn = uval(kPoi2Rows - 1);
      __syncthreads();
      cluster_arrive_release();
      double g_l = 0.0, d_l = 0.0, n_l = 0.0, w_l = 0.0;
      auto row = [&](int j, double up, double dn) {
        const double c = uval(j);
        double v = 4.0 * c - up - dn - rc[j * P - 1] - rc[j * P + 1];
        v = ((umask >> j) & 1u) ? v : 0.0;
        wc[j * W] = v;
        g_l += r[j] * c;
        d_l += c * v;
        n_l += r[j] * r[j];
        w_l += v;
      };
#pragma unroll
      for (int j = 1; j < kPoi2Rows - 1; ++j) row(j, uval(j - 1), uval(j + 1));
      cluster_wait_acquire();
      row(0, rc[-P], uval(1));
      row(kPoi2Rows - 1, uval(kPoi2Rows - 2), rc[kPoi2Rows * P]);
      poisson_cluster_exchange<CL, kPoi2Threads / 32>(g_l, d_l, n_l, w_l, agg, sm, co, parity, (red_phases >> parity) & 1u, rank);
      red_phases ^= 1u << parity;
      const double gamma_new = g_l, delta = d_l;
      rr = n_l;
      if (!(rr > thresh) || it >= p.max_iter) break;
      double beta;
      if (first) { beta = 0.0; alpha = gamma_new / delta; first = false; }
      else {
        beta = gamma_new / gamma;
        const double ga = gamma * alpha;
        alpha = gamma_new * ga / (delta * ga - gamma_new * gamma_new);
      }
      gamma = gamma_new;
      if (tid < kPoi2Coarse) {
        const double sc = co->gathered[parity][tid] + beta * co->sc[tid];
        const double rcv = co->rc[tid] - alpha * sc;
        co->sc[tid] = sc; co->rc[tid] = rcv; co->rc32[tid] = (float)rcv;
      }
      parity ^= 1;
#pragma unroll
      for (int j = 0; j < kPoi2Rows; ++j) {
        pd[j] = uval(j) + beta * pd[j];
        s[j] = wc[j * W] + beta * s[j];
        xc[j * W] += alpha * pd[j];
        r[j] -= alpha * s[j];
      }
      ++it;
      __syncthreads();   // rc32 of the next iteration is complete (and everybody is done with ec)
    }
    gamma = rr;
  } else
  while (true) {
    // publish r: own pixels, and the strip-end rows into the neighbours' halos
#pragma unroll
    for (int j = 0; j < kPoi2Rows; ++j) rc[j * P] = r[j];
    if (halo_up) *halo_up = r[0];
    if (halo_dn) *halo_dn = r[kPoi2Rows - 1];
    __syncthreads();            // the CTA's own rows are visible: rows 1..14 of every strip can start
    cluster_arrive_release();   // ... while the halo rows of the neighbour CTAs are still in flight
    // w = A r, g = (r, r), d = (r, w); vertical neighbours from registers, the rest from shared memory
    double g_l = 0.0, d_l = 0.0;
    auto row = [&](int j, double up, double dn) {
      const double c = r[j];
      double v = 4.0 * c - up - dn - rc[j * P - 1] - rc[j * P + 1];
      v = ((umask >> j) & 1u) ? v : 0.0;
      wc[j * W] = v;
      g_l += c * c;
      d_l += c * v;
    };
#pragma unroll
    for (int j = 1; j < kPoi2Rows - 1; ++j) row(j, r[j - 1], r[j + 1]);
    cluster_wait_acquire();
    row(0, rc[-P], r[1]);
    row(kPoi2Rows - 1, r[kPoi2Rows - 2], rc[kPoi2Rows * P]);
    poisson_cluster_sum2_async<CL, kPoi2Threads / 32>(g_l, d_l, sm, parity, (red_phases >> parity) & 1u, rank);
    red_phases ^= 1u << parity;
    parity ^= 1;
    const double gamma_new = g_l, delta = d_l;
    if (!(gamma_new > thresh) || it >= p.max_iter) { gamma = gamma_new; break; }
    double beta;
    if (first) { beta = 0.0; alpha = gamma_new / delta; first = false; }
    else {
      // a = g / (d - b g / a_old) with b = g / g_old, as one division: a = g g_old a_old / (d g_old a_old - g g)
      beta = gamma_new / gamma;
      const double ga = gamma * alpha;
      alpha = gamma_new * ga / (delta * ga - gamma_new * gamma_new);
    }
    gamma = gamma_new;
#pragma unroll
    for (int j = 0; j < kPoi2Rows; ++j) {
      pd[j] = r[j] + beta * pd[j];
      s[j] = wc[j * W] + beta * s[j];
      xc[j * W] += alpha * pd[j];
      r[j] -= alpha * s[j];
    }
    ++it;
  }

  uint8_t* out = p.out + img_off * 3 + ch;
#pragma unroll 1
  for (int j = 0; j < kPoi2Rows; ++j) {
    if (!((vmask >> j) & 1u)) continue;
    const int lr = lr0 + j;
    const int g = (row0 + lr) * W + col;
    if (p.with_gamma && p.lut_known && !((umask >> j) & 1u)) {
      out[3 * g] = p.lut_known[tgt[3 * g]];
      continue;
    }
    double v = xbuf[lr * W + col];
    if (p.with_gamma) v = v > 0.0 ? pow(v, 2.2) : 0.0;
    v = v > 255.0 ? 255.0 : (v < 0.0 ? 0.0 : v);
    out[3 * g] = (uint8_t)(int)v;
  }
  if (p.stats && rank == 0 && tid == 0) {
    p.stats[2 * sys] = (float)it;
    p.stats[2 * sys + 1] = bb > 0.0 ? (float)sqrt(gamma / bb) : 0.0f;
  }
  cluster.sync();
}

// The coarse operator of the preconditioner and its inverse, one 256-thread CTA per image.
//   A_c[a][a] = 4 |a & U| - 2 #(edges of U inside a),   A_c[a][a'] = -#(edges of U between a and a'),
// (1 on the diagonal of an aggregate without unknowns: its row and column are otherwise zero and never used).
// A_c is symmetric positive definite and banded (the 5-point pattern of the 16 x 16 aggregate grid: half bandwidth
// 16), so: (1) banded Cholesky A_c = L L^T, right-looking, the 16 x 16 trailing window updated by one thread per
// element (256 steps, one barrier each); (2) thread b solves L y = e_b and L^T x = y for column b of the inverse with
// the last 16 values in registers (the band rows are broadcast reads), keeping the lower triangle (rows >= b) in a
// packed triangular shared-memory array; (3) the triangle is written out mirrored, so the fp16 copy is exactly
// symmetric.  fp32 throughout: a preconditioner needs no more.
// out layout: [image][row a][(a' & 15) * 16 + (a' >> 4)] — the order poisson_cg2_kernel<PRE>'s matvec reads.
constexpr int kCoarseBandPitch = 20;   // floats per band row: [0] = 1 / L_ii, [d] = the d-th off-diagonal, d = 1..16
static size_t poisson_coarse_smem_bytes() {
  return sizeof(float) * (3 * (size_t)kPoi2Coarse * kCoarseBandPitch + (size_t)kPoi2Coarse * (kPoi2Coarse + 1) / 2);
}
__global__ void __launch_bounds__(256, 1)
    poisson_coarse_inverse_kernel(const uint8_t* __restrict__ mask, int H, __half* __restrict__ out) {
  constexpr int W = 256, N = kPoi2Coarse, LP = kCoarseBandPitch;
  extern __shared__ __align__(16) float csm[];
  float* Ab = csm;             // working band: Ab[i][d] = A[i][i - d]
  float* Lf = Ab + N * LP;     // rows of L:    Lf[i][d] = L[i][i - d], Lf[i][0] = 1 / L[i][i]
  float* Uf = Lf + N * LP;     // columns of L: Uf[i][d] = L[i + d][i], Uf[i][0] = 1 / L[i][i]
  float* Z = Uf + N * LP;      // packed lower triangle, row i at i (i + 1) / 2
  __shared__ int cnt[4][N];    // unknowns, inner edges, edges to the right / lower aggregate
  const int tid = threadIdx.x;
  const uint8_t* m = mask + (size_t)blockIdx.x * H * W;
  for (int i = tid; i < 4 * N; i += 256) (&cnt[0][0])[i] = 0;
  for (int i = tid; i < 3 * N * LP; i += 256) csm[i] = 0.f;
  __syncthreads();
  auto in_u = [&](int y, int x) {
    return y < H && x < W && (m[y * W + x] != 0 || y == 0 || y == H - 1 || x == 0 || x == W - 1);
  };
  // a thread per 16-pixel row segment of an aggregate
  for (int seg = tid; seg < 16 * N; seg += 256) {
    const int a = seg >> 4, y = (a >> 4) * 16 + (seg & 15), x0 = (a & 15) * 16;
    if (y >= H) continue;
    int n = 0, inner = 0, right = 0, down = 0;
    const bool last_row = (seg & 15) == 15;
#pragma unroll 4
    for (int k = 0; k < 16; ++k) {
      const int x = x0 + k;
      if (!in_u(y, x)) continue;
      ++n;
      if (in_u(y, x + 1)) { if (k == 15) ++right; else ++inner; }
      if (in_u(y + 1, x)) { if (last_row) ++down; else ++inner; }
    }
    if (n) atomicAdd(&cnt[0][a], n);
    if (inner) atomicAdd(&cnt[1][a], inner);
    if (right) atomicAdd(&cnt[2][a], right);
    if (down) atomicAdd(&cnt[3][a], down);
  }
  __syncthreads();
  {
    const int i = tid;
    Ab[i * LP] = cnt[0][i] ? (float)(4 * cnt[0][i] - 2 * cnt[1][i]) : 1.f;
    if (i >= 1 && ((i - 1) & 15) != 15) Ab[i * LP + 1] = -(float)cnt[2][i - 1];
    if (i >= 16) Ab[i * LP + 16] = -(float)cnt[3][i - 16];
  }
  __syncthreads();
  // (1) Cholesky.  Step k reads column k of the working band (rows k + 1 .. k + 16) and the pivot; it writes only
  // elements (i, j) with j > k, so one barrier per step orders everything.
  {
    const int ti = tid >> 4, tj = tid & 15;
#pragma unroll 1
    for (int k = 0; k < N; ++k) {
      const float inv = rsqrtf(fmaxf(Ab[k * LP], 1e-6f));   // (pivots are >= lambda_min(A_c) ~ 0.1: the guard never binds)
      const int i = k + 1 + ti, j = k + 1 + tj;
      if (i < N) {
        const float li = Ab[i * LP + ti + 1] * inv;
        if (tj <= ti) Ab[i * LP + (ti - tj)] -= li * (Ab[j * LP + tj + 1] * inv);
        if (tj == 0) { Lf[i * LP + ti + 1] = li; Uf[k * LP + ti + 1] = li; }
      }
      if (tid == 0) { Lf[k * LP] = inv; Uf[k * LP] = inv; }
      __syncthreads();
    }
  }
  // (2) column b of the inverse
  {
    const int b = tid, first_blk = (b & ~31) >> 4;
    float win[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) win[q] = 0.f;
#pragma unroll 1
    for (int ib = first_blk; ib < 16; ++ib) {
#pragma unroll
      for (int ii = 0; ii < 16; ++ii) {
        const int i = 16 * ib + ii;
        const float4* lr = reinterpret_cast<const float4*>(Lf + i * LP);
        const float4 v0 = lr[0], v1 = lr[1], v2 = lr[2], v3 = lr[3], v4 = lr[4];
        const float l[17] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w,
                             v3.x, v3.y, v3.z, v3.w, v4.x};
        float s0 = i == b ? 1.f : 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int d = 1; d <= 16; d += 4) {
          s0 = fmaf(-l[d], win[(ii - d) & 15], s0);
          s1 = fmaf(-l[d + 1], win[(ii - d - 1) & 15], s1);
          s2 = fmaf(-l[d + 2], win[(ii - d - 2) & 15], s2);
          s3 = fmaf(-l[d + 3], win[(ii - d - 3) & 15], s3);
        }
        const float y = ((s0 + s1) + (s2 + s3)) * l[0];
        win[ii] = y;
        if (i >= b) Z[i * (i + 1) / 2 + b] = y;
      }
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) win[q] = 0.f;
#pragma unroll 1
    for (int ib = 15; ib >= first_blk; --ib) {
#pragma unroll
      for (int ii = 15; ii >= 0; --ii) {
        const int i = 16 * ib + ii;
        const float4* ur = reinterpret_cast<const float4*>(Uf + i * LP);
        const float4 v0 = ur[0], v1 = ur[1], v2 = ur[2], v3 = ur[3], v4 = ur[4];
        const float u[17] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w,
                             v3.x, v3.y, v3.z, v3.w, v4.x};
        float s0 = i >= b ? Z[i * (i + 1) / 2 + b] : 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int d = 1; d <= 16; d += 4) {
          s0 = fmaf(-u[d], win[(ii + d) & 15], s0);
          s1 = fmaf(-u[d + 1], win[(ii + d + 1) & 15], s1);
          s2 = fmaf(-u[d + 2], win[(ii + d + 2) & 15], s2);
          s3 = fmaf(-u[d + 3], win[(ii + d + 3) & 15], s3);
        }
        const float x = ((s0 + s1) + (s2 + s3)) * u[0];
        win[ii] = x;
        if (i >= b) Z[i * (i + 1) / 2 + b] = x;
      }
    }
  }
  __syncthreads();
  // (3) mirrored, permuted fp16 copy: thread = position within an output row
  {
    const int ap = (tid >> 4) + 16 * (tid & 15);   // position tid holds column ap
    __half* o = out + (size_t)blockIdx.x * N * N;
#pragma unroll 4
    for (int a = 0; a < N; ++a) {
      const int hi = max(a, ap), lo = min(a, ap);
      o[a * N + tid] = __float2half_rn(Z[hi * (hi + 1) / 2 + lo] * kPoi2InvScale);
    }
  }
}

constexpr int kPoi2W = 256, kPoi2R = 32;   // the instantiated geometry: 256 columns, H in 249..256
// ... and the same image on 16-CTA clusters (non-portable size) of 256-thread CTAs, two CTAs per SM
constexpr int kPoi2R16 = 16;
static size_t poisson2_smem_bytes(int R, int W, bool pre = false) {
  return sizeof(Poisson2Smem) + (size_t)(R + 2) * (W + 2) * sizeof(double) + 2 * (size_t)R * W * sizeof(double) +
         (pre ? sizeof(Poisson2Coarse) + (size_t)R * kPoi2Coarse * sizeof(__half) : 0);
}

// Stream-ordered scratch for the coarse inverses (128 KB per image) from a private pool per device that keeps its
// memory between calls (the default pool would hand it back to the driver at every synchronisation).
static cudaMemPool_t poisson_pool() {
  static std::mutex mu;
  static cudaMemPool_t pools[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  if (!pools[dev]) {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    if (cudaMemPoolCreate(&pools[dev], &props) != cudaSuccess) { pools[dev] = nullptr; cudaGetLastError(); return nullptr; }
    uint64_t keep = ~0ull;
    cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &keep);
  }
  return pools[dev];
}
static bool poisson2_fits(int R, int W) { return W == kPoi2W && R == kPoi2R; }

static size_t poisson_smem_bytes(int R, int W) {
  return sizeof(PoissonSmem) + (size_t)(R + 2) * W * sizeof(double) + (size_t)R * W * sizeof(double);
}

// Launches the coarse inversion into `dst`, or into stream-ordered scratch returned through *scratch (the caller frees it
// with cudaFreeAsync on the same stream) when dst is null.
static cudaError_t poisson_coarse_launch(const uint8_t* mask, int B, int H, __half* dst, __half** scratch,
                                         cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(poisson_coarse_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)poisson_coarse_smem_bytes());
  if (e != cudaSuccess) return e;
  if (!dst) {
    cudaMemPool_t pool = poisson_pool();
    if (!pool) return cudaErrorMemoryAllocation;
    e = cudaMallocFromPoolAsync(reinterpret_cast<void**>(scratch),
                                (size_t)B * kPoi2Coarse * kPoi2Coarse * sizeof(__half), pool, stream);
    if (e != cudaSuccess) return e;
    dst = *scratch;
  }
  poisson_coarse_inverse_kernel<<<dim3((unsigned)B), 256, poisson_coarse_smem_bytes(), stream>>>(mask, H, dst);
  return cudaGetLastError();
}

static int poisson_launch(const uint8_t* source, const uint8_t* target, const uint8_t* mask, uint8_t* out, int B, int H,
                          int W, int with_gamma, double tol, int max_iter, float* stats, const double* lut_fwd,
                          const uint8_t* lut_known, cudaStream_t stream) {
  if (!source || !target || !mask || !out || B <= 0 || H < 3 || W < 3) {
    set_error("chb_poisson_blend: bad arguments (need B > 0, H, W >= 3)");
    return CHB_ERR_ARG;
  }
  const int R = (H + kPoiCluster - 1) / kPoiCluster;
  if ((long long)R * W > (long long)kPoiPx * kPoiThreads || W > 512) {
    set_error("chb_poisson_blend: image too large for the cluster-resident solver (ceil(H/8)*W <= 8192, W <= 512)");
    return CHB_ERR_ARG;
  }
  if (!(tol > 0.0) || max_iter < 0) {
    set_error("chb_poisson_blend: need tol > 0 and max_iter >= 0");
    return CHB_ERR_ARG;
  }
  int rc = chb_check_device();
  if (rc != CHB_OK) return rc;
#ifdef CHB_TUNING_ENV  // A/B switch, tuning builds only (tools/build_variant.py with VARIANT_FLAGS=-DCHB_TUNING_ENV)
  static const bool force_v1 = [] { const char* v = getenv("CHB_POISSON_V1"); return v && atoi(v) != 0; }();
#else
  const bool force_v1 = false;
#endif
  // second generation for the reference's image size; first generation for any other W <= 512, ceil(H/8)*W <= 8192
  const bool v2 = !force_v1 && poisson2_fits(R, W);
  size_t smem = v2 ? poisson2_smem_bytes(R, W) : poisson_smem_bytes(R, W);
  // 16-CTA clusters (non-portable size, 256-thread CTAs) halve the rows per CTA: 20 % lower latency while at most 8
  // clusters are in flight (interactive B <= 2, the reference's own use), same throughput as 8-CTA clusters beyond
  // that (measured, profiles/r1_s_poisson_solver.txt).  CHB_POISSON_CL16 = 0 / 1 overrides the choice.
#ifdef CHB_TUNING_ENV
  static const int force_cl16 = [] { const char* v = getenv("CHB_POISSON_CL16"); return v ? (atoi(v) != 0 ? 1 : 0) : -1; }();
#else
  const int force_cl16 = -1;
#endif
  // The preconditioned solver (8-CTA clusters) is the product path for 256-column images; the plain second-generation
  // kernels stay for A/B runs of tuning builds (CHB_POISSON_PRE=0).
#ifdef CHB_TUNING_ENV
  static const bool no_pre = [] { const char* v = getenv("CHB_POISSON_PRE"); return v && atoi(v) == 0; }();
#else
  const bool no_pre = false;
#endif
  const bool pre = v2 && !no_pre;
  // 16-CTA clusters for interactive batches (see above), 8-CTA clusters otherwise; with or without the preconditioner
  const bool cl16 = v2 && (force_cl16 >= 0 ? force_cl16 == 1 : B * 3 <= 8);
  typedef void (*Poi2Kernel)(const PoissonParams);
  const Poi2Kernel k2 = pre ? (cl16 ? poisson_cg2_kernel<kPoi2W, kPoi2R16, 16, true>
                                    : poisson_cg2_kernel<kPoi2W, kPoi2R, kPoiCluster, true>)
                            : (cl16 ? poisson_cg2_kernel<kPoi2W, kPoi2R16, 16, false>
                                    : poisson_cg2_kernel<kPoi2W, kPoi2R, kPoiCluster, false>);
  if (v2) smem = poisson2_smem_bytes(cl16 ? kPoi2R16 : R, W, pre);
  // (set on every call: the attribute is per device, and a process may drive several)
  cudaError_t e;
  if (v2) {
    e = cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess && cl16) e = cudaFuncSetAttribute(k2, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  } else {
    e = cudaFuncSetAttribute(poisson_cg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  if (e != cudaSuccess) {
    set_error(std::string("poisson kernel attribute: ") + cudaGetErrorString(e));
    return CHB_ERR_CUDA;
  }
  PoissonParams p;
  p.source = source; p.target = target; p.mask = mask; p.out = out; p.stats = stats;
  p.lut_fwd = lut_fwd; p.lut_known = lut_known;
  p.B = B; p.H = H; p.W = W; p.R = R;
  p.with_gamma = with_gamma ? 1 : 0; p.max_iter = max_iter; p.tol2 = tol * tol;
  p.coarse_inv = nullptr;
  __half* coarse = nullptr;
  if (pre) {
    e = poisson_coarse_launch(mask, B, H, nullptr, &coarse, stream);
    if (e != cudaSuccess) {
      set_error(std::string("poisson coarse inverse: ") + cudaGetErrorString(e));
      cudaGetLastError();
      return CHB_ERR_CUDA;
    }
    p.coarse_inv = coarse;
  }
  if (v2) {
    const int cl = cl16 ? 16 : kPoiCluster;
    const int threads = (cl16 ? kPoi2R16 : kPoi2R) / kPoi2Rows * kPoi2W;
    if (cl16) p.R = kPoi2R16;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * 3 * cl));
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, k2, p);
    if (coarse) cudaFreeAsync(coarse, stream);
    if (e != cudaSuccess) {
      set_error(std::string("poisson_cg2 launch: ") + cudaGetErrorString(e));
      return CHB_ERR_CUDA;
    }
  } else {
    poisson_cg_kernel<<<dim3((unsigned)(B * 3 * kPoiCluster)), kPoiThreads, smem, stream>>>(p);
  }
  return blend_check("poisson_cg");
}

// ------------------------------------------------------------------------------------------------------------------
// 8-bit colour space, one thread per triple.
// RGB -> HSV: OpenCV's RGB2HSV_b (fixed point, hsv_shift 12, H range 180), as cv2.cvtColor(uint8, COLOR_RGB2HSV).
__device__ __forceinline__ int round_half_even_div(long long num, long long den) {  // cvRound(num / den), num, den > 0
  long long qd = num / den, rem = num - qd * den;
  if (2 * rem > den || (2 * rem == den && (qd & 1))) ++qd;
  return (int)qd;
}

__global__ void rgb_to_hsv_kernel(const float* __restrict__ rgb_f, const uint8_t* __restrict__ rgb_u8,
                                  uint8_t* __restrict__ hsv, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int r, g, b;
    if (rgb_f) {  // ui/backend.py:98-99: the float colour is cast with astype('uint8') first
      r = f32_to_u8_trunc(rgb_f[3 * i]); g = f32_to_u8_trunc(rgb_f[3 * i + 1]); b = f32_to_u8_trunc(rgb_f[3 * i + 2]);
    } else {
      r = rgb_u8[3 * i]; g = rgb_u8[3 * i + 1]; b = rgb_u8[3 * i + 2];
    }
    const int v = max(r, max(g, b)), vmin = min(r, min(g, b)), diff = v - vmin;
    const int sdiv = v ? round_half_even_div(255LL << 12, v) : 0;
    const int hdiv = diff ? round_half_even_div(180LL << 12, 6LL * diff) : 0;
    const int s = (diff * sdiv + (1 << 11)) >> 12;
    int h = (v == r) ? (g - b) : ((v == g) ? (b - r + 2 * diff) : (r - g + 4 * diff));
    h = (h * hdiv + (1 << 11)) >> 12;   // arithmetic shift of a possibly negative value, as in OpenCV
    if (h < 0) h += 180;
    hsv[3 * i] = (uint8_t)h; hsv[3 * i + 1] = (uint8_t)s; hsv[3 * i + 2] = (uint8_t)v;
  }
}

// HSV -> RGB: OpenCV's scalar HSV2RGB_b path (what a one-pixel cv2.cvtColor call runs): float32 HSV2RGB_native with
// hscale 6/180 on (h, s/255, v/255); the shipped binary fuses `1 - s*h` and `1 - s*(1-h)` (pinned exhaustively).
__global__ void hsv_to_rgb_kernel(const uint8_t* __restrict__ hsv, uint8_t* __restrict__ rgb, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float h = (float)hsv[3 * i];
    const float s = __fmul_rn((float)hsv[3 * i + 1], 1.f / 255.f), v = __fmul_rn((float)hsv[3 * i + 2], 1.f / 255.f);
    float bb, gg, rr;
    if (hsv[3 * i + 1] == 0) {
      bb = gg = rr = v;
    } else {
      h = __fmul_rn(h, 6.f / 180.f);
      h = fmodf(h, 6.f);
      int sector = (int)floorf(h);
      h = __fsub_rn(h, (float)sector);
      if ((unsigned)sector >= 6u) { sector = 0; h = 0.f; }
      float tab[4];
      tab[0] = v;
      tab[1] = __fmul_rn(v, __fsub_rn(1.f, s));
      tab[2] = __fmul_rn(v, __fmaf_rn(-s, h, 1.f));
      tab[3] = __fmul_rn(v, __fmaf_rn(-s, __fsub_rn(1.f, h), 1.f));
      const int sd[6][3] = {{1, 3, 0}, {1, 0, 2}, {3, 0, 1}, {0, 2, 1}, {0, 1, 3}, {2, 1, 0}};
      bb = tab[sd[sector][0]]; gg = tab[sd[sector][1]]; rr = tab[sd[sector][2]];
    }
    const float o[3] = {rr, gg, bb};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      int q = __float2int_rn(__fmul_rn(o[c], 255.f));  // saturate_cast<uchar>: round half to even, clamp
      rgb[3 * i + c] = (uint8_t)min(255, max(0, q));
    }
  }
}

// shape_util.py:17-20: label = argmax over channels (first maximum wins), 255 where every channel is 0.
__global__ void onehot_to_label_kernel(const float* __restrict__ one_hot, uint8_t* __restrict__ labels, int B, int C,
                                       long long hw) {
  const long long total = (long long)B * hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / hw, pix = i - b * hw;
    const float* src = one_hot + b * C * hw + pix;
    float best = src[0];
    int arg = 0;
    for (int c = 1; c < C; ++c) {
      const float v = src[c * hw];
      if (v > best) { best = v; arg = c; }
    }
    labels[i] = (best == 0.f) ? (uint8_t)255 : (uint8_t)arg;
  }
}

// shape_util.py:6-14: uint8 labels (255 = none) -> float32 one-hot [B,C,H,W]; out-of-range labels give an all-zero pixel.
__global__ void label_to_onehot_kernel(const uint8_t* __restrict__ labels, float* __restrict__ out, int B, int C,
                                       long long hw) {
  const long long total = (long long)B * C * hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i % hw, bc = i / hw;
    const int c = (int)(bc % C);
    const long long b = bc / C;
    out[i] = labels[b * hw + pix] == c ? 1.f : 0.f;
  }
}

}  // namespace chb

extern "C" {

int chb_image_to_u8(const float* img, uint8_t* out, int B, int H, int W, void* stream) {
  using namespace chb;
  if (!img || !out || B <= 0 || H <= 0 || W <= 0) {
    set_error("chb_image_to_u8: bad arguments");
    return CHB_ERR_ARG;
  }
  image_to_u8_kernel<<<blend_grid((long long)B * H * W, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      img, out, B, H, W);
  return blend_check("image_to_u8");
}

int chb_blend_mask(const uint8_t* target_parsing, const uint8_t* face_parsing, uint8_t* res_mask_dilated,
                   uint8_t* solve_mask, int B, int H, int W, void* stream) {
  using namespace chb;
  if (!target_parsing || !face_parsing || (!res_mask_dilated && !solve_mask) || B <= 0 || H <= 0 || W <= 0) {
    set_error("chb_blend_mask: bad arguments");
    return CHB_ERR_ARG;
  }
  blend_mask_kernel<<<blend_grid((long long)B * H * W, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      target_parsing, face_parsing, res_mask_dilated, solve_mask, B, H, W);
  return blend_check("blend_mask");
}

int chb_poisson_blend(const uint8_t* source, const uint8_t* target, const uint8_t* mask, uint8_t* out, int B, int H,
                      int W, int with_gamma, double tol, int max_iter, float* stats, const double* lut_fwd,
                      const uint8_t* lut_known, void* stream) {
  return chb::poisson_launch(source, target, mask, out, B, H, W, with_gamma, tol, max_iter, stats, lut_fwd, lut_known,
                             reinterpret_cast<cudaStream_t>(stream));
}

int chb_poisson_coarse_inverse(const uint8_t* mask, uint16_t* inv_f16, int B, int H, void* stream) {
  using namespace chb;
  if (!mask || !inv_f16 || B <= 0 || H < 3 || H > 256) {
    set_error("chb_poisson_coarse_inverse: bad arguments (256-column masks, 3 <= H <= 256)");
    return CHB_ERR_ARG;
  }
  int rc = chb_check_device();
  if (rc != CHB_OK) return rc;
  cudaError_t e = poisson_coarse_launch(mask, B, H, reinterpret_cast<__half*>(inv_f16), nullptr,
                                        reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) {
    set_error(std::string("chb_poisson_coarse_inverse: ") + cudaGetErrorString(e));
    return CHB_ERR_CUDA;
  }
  return CHB_OK;
}

int64_t chb_postprocess_workspace_bytes(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  return (int64_t)B * H * W * 4;  // uint8 generated image [B,H,W,3] + solve mask [B,H,W]
}

int chb_postprocess_blending(const uint8_t* face_img, const float* res_img, const uint8_t* face_parsing,
                             const uint8_t* target_parsing, uint8_t* out, uint8_t* res_mask_dilated, void* workspace,
                             int B, int H, int W, int blending, double tol, int max_iter, float* stats,
                             const double* lut_fwd, const uint8_t* lut_known, void* stream) {
  using namespace chb;
  if (!res_img || !out || B <= 0 || H <= 0 || W <= 0) {
    set_error("chb_postprocess_blending: bad arguments");
    return CHB_ERR_ARG;
  }
  if (!blending) return chb_image_to_u8(res_img, out, B, H, W, stream);  // hair_editor.py:307-308
  if (!face_img || !face_parsing || !target_parsing || !workspace) {
    set_error("chb_postprocess_blending: blending needs the face image, both parsings and a workspace");
    return CHB_ERR_ARG;
  }
  uint8_t* gen_u8 = reinterpret_cast<uint8_t*>(workspace);
  uint8_t* solve_mask = gen_u8 + (size_t)B * H * W * 3;
  int rc = chb_image_to_u8(res_img, gen_u8, B, H, W, stream);
  if (rc != CHB_OK) return rc;
  rc = chb_blend_mask(target_parsing, face_parsing, res_mask_dilated, solve_mask, B, H, W, stream);
  if (rc != CHB_OK) return rc;
  return chb_poisson_blend(face_img, gen_u8, solve_mask, out, B, H, W, 1, tol, max_iter, stats, lut_fwd, lut_known,
                           stream);
}

int chb_rgb_to_hsv(const float* rgb_f32, const uint8_t* rgb_u8, uint8_t* hsv, int64_t n, void* stream) {
  using namespace chb;
  if ((!rgb_f32 == !rgb_u8) || !hsv || n <= 0) {
    set_error("chb_rgb_to_hsv: pass exactly one of rgb_f32 / rgb_u8, and n > 0");
    return CHB_ERR_ARG;
  }
  rgb_to_hsv_kernel<<<blend_grid(n, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(rgb_f32, rgb_u8, hsv, n);
  return blend_check("rgb_to_hsv");
}

int chb_hsv_to_rgb(const uint8_t* hsv, uint8_t* rgb, int64_t n, void* stream) {
  using namespace chb;
  if (!hsv || !rgb || n <= 0) {
    set_error("chb_hsv_to_rgb: bad arguments");
    return CHB_ERR_ARG;
  }
  hsv_to_rgb_kernel<<<blend_grid(n, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(hsv, rgb, n);
  return blend_check("hsv_to_rgb");
}

int chb_onehot_to_label(const float* one_hot, uint8_t* labels, int B, int C, int64_t hw, void* stream) {
  using namespace chb;
  if (!one_hot || !labels || B <= 0 || C <= 0 || C > 255 || hw <= 0) {
    set_error("chb_onehot_to_label: bad arguments");
    return CHB_ERR_ARG;
  }
  onehot_to_label_kernel<<<blend_grid((long long)B * hw, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      one_hot, labels, B, C, hw);
  return blend_check("onehot_to_label");
}

int chb_label_to_onehot(const uint8_t* labels, float* one_hot, int B, int C, int64_t hw, void* stream) {
  using namespace chb;
  if (!one_hot || !labels || B <= 0 || C <= 0 || C > 255 || hw <= 0) {
    set_error("chb_label_to_onehot: bad arguments");
    return CHB_ERR_ARG;
  }
  label_to_onehot_kernel<<<blend_grid((long long)B * C * hw, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      labels, one_hot, B, C, hw);
  return blend_check("label_to_onehot");
}

}  // extern "C"
