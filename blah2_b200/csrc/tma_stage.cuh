// tma_stage.cuh -- staging of float2 IQ windows into shared memory by TMA bulk copies (sm_100a).
//
// One elected thread issues cp.async.bulk global -> shared (SASS UBLKCP) for the window a CTA will transform
// NEXT while the CTA works on the current one; completion is signalled on an mbarrier (expect_tx / complete_tx),
// the other threads only test its phase bit.  No registers are held for data in flight and no thread waits on a
// global load in front of its butterfly -- which is what the round-1 kernels spent 15-30 % of their stall samples
// on (profiles/r02_summary.md).
//
// Bulk copies need 16-byte aligned source, destination and size, but a float2 window starts on any 8-byte
// boundary: the copy covers the 16-byte aligned INTERIOR [i0, i1) of the window and lands at S[i + par]
// (par = 1 when the window starts on an odd float2, which keeps the destination aligned); the at most one element
// on either side is read with an ordinary load by the thread that owns it.  No byte outside the caller's buffer
// is touched.  Windows that are not one contiguous run (circular wrap of the Wiener-Hopf correlations, the
// reference's index quirk for delayMin > 0) are not staged: `src` stays null and their elements are loaded directly.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace b2 {
namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
  } while (!done);
}

// One staged window.  Element m (m < n) of the window is  S[m + par]  when i0 <= m < i1, else  src[m].
struct Window {
  const float2 *src = nullptr;  // window element 0 in global memory; null: not staged
  int par = 0, i0 = 0, i1 = 0;
  __device__ __forceinline__ uint32_t bytes() const { return (uint32_t)(i1 - i0) * 8u; }
};

// describe the window of n contiguous elements starting at p (n <= capacity of S minus 2)
__device__ __forceinline__ Window make_window(const float2 *p, int n) {
  Window w;
  w.src = p;
  w.par = (int)((reinterpret_cast<uintptr_t>(p) >> 3) & 1);
  w.i0 = w.par;
  w.i1 = n - (int)((reinterpret_cast<uintptr_t>(p + n) >> 3) & 1);
  if (w.i1 < w.i0) w.i1 = w.i0;
  return w;
}

// one thread: start the copy of w's interior into S (16-byte aligned) and arm the barrier with its byte count
__device__ __forceinline__ void issue(const Window &w, float2 *S, uint64_t *bar) {
  const uint32_t b = w.src ? w.bytes() : 0u;
  mbar_expect_tx(bar, b);
  if (b) bulk_g2s(S + w.i0 + w.par, w.src + w.i0, b, bar);
}

// element m of the window (the caller masks m >= n itself)
__device__ __forceinline__ float2 read(const Window &w, const float2 *S, int m) {
  float2 v = S[m + w.par];
  if (m < w.i0 || m >= w.i1) v = w.src[m];
  return v;
}

}  // namespace tma
}  // namespace b2
