// Thin inline-PTX wrappers for the sm_100a features the conv kernels use:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld / fences).
// Everything here is device-only and header-only.
#pragma once
#include <cstdint>
#include <cuda.h>

namespace chb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (→ launch failure the host can report) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 8000000000ll) __trap();  // ~4 s at boost clock
  }
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* m, void* smem, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* m, void* smem, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, fp16 inputs, fp32 accumulate. One thread issues.
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// KS (2 or 4) MMAs over consecutive 16-element K steps of one smem chunk, issued from a single asm block so that
// the issuing thread spends as few instructions per MMA as possible (it is the only thread feeding the tensor
// core).  Descriptor low words advance by 2 (= 32 bytes >> 4) per K step; `accumulate` only gates the first MMA.
template <int KS>
__device__ __forceinline__ void umma_f16_ss_k(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate);
template <>
__device__ __forceinline__ void umma_f16_ss_k<4>(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, pt;\n\t"
      ".reg .b64 a1, b1, a2, b2, a3, b3;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.eq.b32 pt, %4, %4;\n\t"
      "add.u64 a1, %1, 2;\n\t"
      "add.u64 b1, %2, 2;\n\t"
      "add.u64 a2, %1, 4;\n\t"
      "add.u64 b2, %2, 4;\n\t"
      "add.u64 a3, %1, 6;\n\t"
      "add.u64 b3, %2, 6;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %3, pt;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <>
__device__ __forceinline__ void umma_f16_ss_k<2>(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, pt;\n\t"
      ".reg .b64 a1, b1;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.eq.b32 pt, %4, %4;\n\t"
      "add.u64 a1, %1, 2;\n\t"
      "add.u64 b1, %2, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, pt;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// One kernel row of a 3x3 tap group (taps kx = 0, 1, 2 of one ky) issued from a single asm block: 3 x KS MMAs.
// The A view advances by a_step16 (= row_bytes >> 4: one pixel to the right inside the halo tile) and the weight slab by
// b_step16 per tap, both in descriptor address units (16 bytes).  For narrow N tiles (N <= 128: 32-64 tensor-core
// cycles per MMA) the per-tap loop overhead of the issuing warp, not the tensor core, sets the pace otherwise.
template <int KS>
__device__ __forceinline__ void umma_f16_ss_row3(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t a_step16,
                                                 uint32_t b_step16, uint32_t idesc, uint32_t accumulate);
template <>
__device__ __forceinline__ void umma_f16_ss_row3<4>(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t a_step16,
                                                    uint32_t b_step16, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, pt;\n\t"
      ".reg .b64 a, b, as, bs, ta, tb;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "setp.eq.b32 pt, %6, %6;\n\t"
      "cvt.u64.u32 as, %3;\n\t"
      "cvt.u64.u32 bs, %4;\n\t"
      "mov.b64 a, %1;\n\t"
      "mov.b64 b, %2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, p;\n\t"
      "add.u64 ta, a, 2;\n\t"
      "add.u64 tb, b, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %5, pt;\n\t"
      "add.u64 ta, a, 4;\n\t"
      "add.u64 tb, b, 4;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %5, pt;\n\t"
      "add.u64 ta, a, 6;\n\t"
      "add.u64 tb, b, 6;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %5, pt;\n\t"
      "add.u64 a, a, as;\n\t"
      "add.u64 b, b, bs;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, pt;\n\t"
      "add.u64 ta, a, 2;\n\t"
      "add.u64 tb, b, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %5, pt;\n\t"
      "add.u64 ta, a, 4;\n\t"
      "add.u64 tb, b, 4;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %5, pt;\n\t"
      "add.u64 ta, a, 6;\n\t"
      "add.u64 tb, b, 6;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %5, pt;\n\t"
      "add.u64 a, a, as;\n\t"
      "add.u64 b, b, bs;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, pt;\n\t"
      "add.u64 ta, a, 2;\n\t"
      "add.u64 tb, b, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %5, pt;\n\t"
      "add.u64 ta, a, 4;\n\t"
      "add.u64 tb, b, 4;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %5, pt;\n\t"
      "add.u64 ta, a, 6;\n\t"
      "add.u64 tb, b, 6;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %5, pt;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(a_step16), "r"(b_step16), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <>
__device__ __forceinline__ void umma_f16_ss_row3<2>(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t a_step16,
                                                    uint32_t b_step16, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, pt;\n\t"
      ".reg .b64 a, b, as, bs, ta, tb;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "setp.eq.b32 pt, %6, %6;\n\t"
      "cvt.u64.u32 as, %3;\n\t"
      "cvt.u64.u32 bs, %4;\n\t"
      "mov.b64 a, %1;\n\t"
      "mov.b64 b, %2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, p;\n\t"
      "add.u64 ta, a, 2;\n\t"
      "add.u64 tb, b, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %5, pt;\n\t"
      "add.u64 a, a, as;\n\t"
      "add.u64 b, b, bs;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, pt;\n\t"
      "add.u64 ta, a, 2;\n\t"
      "add.u64 tb, b, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %5, pt;\n\t"
      "add.u64 a, a, as;\n\t"
      "add.u64 b, b, bs;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, pt;\n\t"
      "add.u64 ta, a, 2;\n\t"
      "add.u64 tb, b, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %5, pt;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(a_step16), "r"(b_step16), "r"(idesc), "r"(accumulate)
      : "memory");
}
// All nine taps of a 3x3 segment chunk from one asm block (weight-stationary mode: no per-tap barrier): tap (ky, kx)
// reads the halo view at ky * row_step16 + kx * a_step16 and the weight slab at (ky*3 + kx) * b_step16 (16-byte units).
template <int KS>
__device__ __forceinline__ void umma_f16_ss_tile9(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t a_step16,
                                                  uint32_t row_step16, uint32_t b_step16, uint32_t idesc,
                                                  uint32_t accumulate);
template <>
__device__ __forceinline__ void umma_f16_ss_tile9<4>(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t a_step16,
                                                     uint32_t row_step16, uint32_t b_step16, uint32_t idesc,
                                                     uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, pt;\n\t"
      ".reg .b64 a, b, ar, as, rs, bs, ta, tb;\n\t"
      "setp.ne.b32 p, %7, 0;\n\t"
      "setp.eq.b32 pt, %7, %7;\n\t"
      "cvt.u64.u32 as, %3;\n\t"
      "cvt.u64.u32 rs, %4;\n\t"
      "cvt.u64.u32 bs, %5;\n\t"
      "mov.b64 ar, %1;\n\t"
      "mov.b64 b, %2;\n\t"
      "mov.b64 a, ar;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %6, p;\n\t"
      "add.u64 ta, a, 2;\n\t"
      "add.u64 tb, b, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "add.u64 ta, a, 4;\n\t"
      "add.u64 tb, b, 4;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "add.u64 ta, a, 6;\n\t"
      "add.u64 tb, b, 6;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "add.u64 a, a, as;\n\t"
      "add.u64 b, b, bs;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %6, pt;\n\t"
      "add.u64 ta, a, 2;\n\t"
      "add.u64 tb, b, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "add.u64 ta, a, 4;\n\t"
      "add.u64 tb, b, 4;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "add.u64 ta, a, 6;\n\t"
      "add.u64 tb, b, 6;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "add.u64 a, a, as;\n\t"
      "add.u64 b, b, bs;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %6, pt;\n\t"
      "add.u64 ta, a, 2;\n\t"
      "add.u64 tb, b, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "add.u64 ta, a, 4;\n\t"
      "add.u64 tb, b, 4;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "add.u64 ta, a, 6;\n\t"
      "add.u64 tb, b, 6;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "add.u64 ar, ar, rs;\n\t"
      "mov.b64 a, ar;\n\t"
      "add.u64 b, b, bs;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %6, pt;\n\t"
      "add.u64 ta, a, 2;\n\t"
      "add.u64 tb, b, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "add.u64 ta, a, 4;\n\t"
      "add.u64 tb, b, 4;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "add.u64 ta, a, 6;\n\t"
      "add.u64 tb, b, 6;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "add.u64 a, a, as;\n\t"
      "add.u64 b, b, bs;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %6, pt;\n\t"
      "add.u64 ta, a, 2;\n\t"
      "add.u64 tb, b, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "add.u64 ta, a, 4;\n\t"
      "add.u64 tb, b, 4;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "add.u64 ta, a, 6;\n\t"
      "add.u64 tb, b, 6;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "add.u64 a, a, as;\n\t"
      "add.u64 b, b, bs;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %6, pt;\n\t"
      "add.u64 ta, a, 2;\n\t"
      "add.u64 tb, b, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "add.u64 ta, a, 4;\n\t"
      "add.u64 tb, b, 4;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "add.u64 ta, a, 6;\n\t"
      "add.u64 tb, b, 6;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "add.u64 ar, ar, rs;\n\t"
      "mov.b64 a, ar;\n\t"
      "add.u64 b, b, bs;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %6, pt;\n\t"
      "add.u64 ta, a, 2;\n\t"
      "add.u64 tb, b, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "add.u64 ta, a, 4;\n\t"
      "add.u64 tb, b, 4;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "add.u64 ta, a, 6;\n\t"
      "add.u64 tb, b, 6;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "add.u64 a, a, as;\n\t"
      "add.u64 b, b, bs;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %6, pt;\n\t"
      "add.u64 ta, a, 2;\n\t"
      "add.u64 tb, b, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "add.u64 ta, a, 4;\n\t"
      "add.u64 tb, b, 4;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "add.u64 ta, a, 6;\n\t"
      "add.u64 tb, b, 6;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "add.u64 a, a, as;\n\t"
      "add.u64 b, b, bs;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %6, pt;\n\t"
      "add.u64 ta, a, 2;\n\t"
      "add.u64 tb, b, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "add.u64 ta, a, 4;\n\t"
      "add.u64 tb, b, 4;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "add.u64 ta, a, 6;\n\t"
      "add.u64 tb, b, 6;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(a_step16), "r"(row_step16), "r"(b_step16), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <>
__device__ __forceinline__ void umma_f16_ss_tile9<2>(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t a_step16,
                                                     uint32_t row_step16, uint32_t b_step16, uint32_t idesc,
                                                     uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, pt;\n\t"
      ".reg .b64 a, b, ar, as, rs, bs, ta, tb;\n\t"
      "setp.ne.b32 p, %7, 0;\n\t"
      "setp.eq.b32 pt, %7, %7;\n\t"
      "cvt.u64.u32 as, %3;\n\t"
      "cvt.u64.u32 rs, %4;\n\t"
      "cvt.u64.u32 bs, %5;\n\t"
      "mov.b64 ar, %1;\n\t"
      "mov.b64 b, %2;\n\t"
      "mov.b64 a, ar;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %6, p;\n\t"
      "add.u64 ta, a, 2;\n\t"
      "add.u64 tb, b, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "add.u64 a, a, as;\n\t"
      "add.u64 b, b, bs;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %6, pt;\n\t"
      "add.u64 ta, a, 2;\n\t"
      "add.u64 tb, b, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "add.u64 a, a, as;\n\t"
      "add.u64 b, b, bs;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %6, pt;\n\t"
      "add.u64 ta, a, 2;\n\t"
      "add.u64 tb, b, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "add.u64 ar, ar, rs;\n\t"
      "mov.b64 a, ar;\n\t"
      "add.u64 b, b, bs;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %6, pt;\n\t"
      "add.u64 ta, a, 2;\n\t"
      "add.u64 tb, b, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "add.u64 a, a, as;\n\t"
      "add.u64 b, b, bs;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %6, pt;\n\t"
      "add.u64 ta, a, 2;\n\t"
      "add.u64 tb, b, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "add.u64 a, a, as;\n\t"
      "add.u64 b, b, bs;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %6, pt;\n\t"
      "add.u64 ta, a, 2;\n\t"
      "add.u64 tb, b, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "add.u64 ar, ar, rs;\n\t"
      "mov.b64 a, ar;\n\t"
      "add.u64 b, b, bs;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %6, pt;\n\t"
      "add.u64 ta, a, 2;\n\t"
      "add.u64 tb, b, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "add.u64 a, a, as;\n\t"
      "add.u64 b, b, bs;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %6, pt;\n\t"
      "add.u64 ta, a, 2;\n\t"
      "add.u64 tb, b, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "add.u64 a, a, as;\n\t"
      "add.u64 b, b, bs;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %6, pt;\n\t"
      "add.u64 ta, a, 2;\n\t"
      "add.u64 tb, b, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %6, pt;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(a_step16), "r"(row_step16), "r"(b_step16), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Descriptor from a precomputed high word (stride / version / layout) and a shared-memory address.
__device__ __forceinline__ uint32_t umma_desc_hi(uint32_t row_bytes, uint32_t sbo_bytes) {
  return (sbo_bytes >> 4) | (1u << 14) | ((row_bytes == 128 ? 2u : 4u) << 29);
}
__device__ __forceinline__ uint64_t umma_desc_make(uint32_t hi, uint32_t saddr) {
  const uint32_t lo = ((saddr & 0x3FFFFu) >> 4) | (1u << 16);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}

// Arrives on the mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// tcgen05.ld 32x32b: thread i of the warp receives N consecutive fp32 columns of TMEM lane (base_lane + i).
// The load is asynchronous: values may only be read after tmem_ld_fence(v) (tcgen05.wait::ld).  The fence takes
// the destination registers as in/out operands so the compiler cannot hoist their uses above the wait.
template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[N]);

template <>
__device__ __forceinline__ void tmem_ld<8>(uint32_t taddr, float (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "r"(taddr)
               : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, float (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
        "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
      : "r"(taddr)
      : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld<32>(uint32_t taddr, float (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
        "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]),
        "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]),
        "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
      : "r"(taddr)
      : "memory");
}

template <int N>
__device__ __forceinline__ void tmem_ld_fence(float (&v)[N]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < N; i += 8) {
    asm volatile(""
                 : "+f"(v[i]), "+f"(v[i + 1]), "+f"(v[i + 2]), "+f"(v[i + 3]), "+f"(v[i + 4]), "+f"(v[i + 5]),
                   "+f"(v[i + 6]), "+f"(v[i + 7])
                 :
                 : "memory");
  }
}

// tcgen05.st 32x32b: the inverse of tmem_ld<8> (thread i writes 8 consecutive fp32 columns of TMEM lane base_lane + i).
// Used by the split-K combine, which puts the summed partial accumulators back where the epilogue expects them.
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "f"(v[0]),
               "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Release / acquire fence at GPU scope (lighter than __threadfence(), which is sequentially consistent).
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// Named barrier `id` over `nthreads` threads (a multiple of 32) of the CTA; id 0 is __syncthreads'.
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Shared-memory matrix descriptor, K-major operand, rows of `row_bytes` (128 -> SWIZZLE_128B, 64 -> SWIZZLE_64B).
// 8-row groups are `row_bytes * 8` apart (dense tile as written by a TMA box whose inner extent is row_bytes).
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t row_bytes) {
  const uint64_t layout = (row_bytes == 128) ? 2ull : 4ull;  // SWIZZLE_128B : SWIZZLE_64B
  const uint64_t sbo = (row_bytes * 8u) >> 4;
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);  // start address, 16B units
  d |= 1ull << 16;                                      // leading byte offset (ignored for swizzled K-major)
  d |= sbo << 32;                                       // stride byte offset between 8-row groups
  d |= 1ull << 46;                                      // descriptor version (sm_100)
  d |= layout << 61;
  return d;
}

// General K-major operand: rows of `row_bytes` (128 -> SWIZZLE_128B, 64 -> SWIZZLE_64B), 8-row groups `sbo_bytes`
// apart, any 16-byte aligned start address (the hardware swizzles on absolute address bits; base_offset = 0).
__device__ __forceinline__ uint64_t umma_smem_desc_k(uint32_t saddr, uint32_t row_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= 1ull << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= 1ull << 46;
  d |= (row_bytes == 128 ? 2ull : 4ull) << 61;
  return d;
}

// Same, for a SWIZZLE_128B K-major operand whose 8-row groups are `sbo_bytes` apart and whose start address need
// not be 1024-byte aligned (a tap view into a halo tile): `base_offset` is the descriptor's 3-bit swizzle phase.
__device__ __forceinline__ uint64_t umma_smem_desc_sw128(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_offset) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= 1ull << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= 1ull << 46;
  d |= static_cast<uint64_t>(base_offset & 7u) << 49;
  d |= 2ull << 61;
  return d;
}

// Instruction descriptor: A,B = fp16 K-major, D = fp32, M = 128, N = n.
__device__ __forceinline__ uint32_t umma_idesc_f16(uint32_t n) {
  return (1u << 4) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

}  // namespace chb
