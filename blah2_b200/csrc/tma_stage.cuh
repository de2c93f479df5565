// tma_stage.cuh -- staging of float2 IQ windows into shared memory by TMA bulk copies (sm_100a).
//
// One elected thread issues cp.async.bulk global -> shared (SASS UBLKCP) for the window a CTA will transform
// NEXT while the CTA works on the current one; completion is signalled on an mbarrier (expect_tx / complete_tx),
// the other threads only test its phase bit.  No registers are held for data in flight and no thread waits on a
// global load in front of its butterfly -- which is what the round-1 kernels spent 15-30 % of their stall samples
// on (profiles/r02_summary.md).
//
// Bulk copies need 16-byte aligned source, destination and size, but a float2 window starts on any 8-byte
// boundary: the copy is widened to the enclosing 16-byte aligned range -- at most one float2 before and one after
// the window, both INSIDE the caller's array [0, N) -- and lands at S[0]; window element m is then S[m + par]
// (par = 1 when the window starts on an odd float2).  A window whose widening would leave the array (first /
// last element of a misaligned array), or that is not one contiguous run (circular wrap of the Wiener-Hopf
// correlations, the reference's index quirk for delayMin > 0), is not staged: `src` stays null and the caller
// loads its elements directly.  No byte outside the caller's buffer is touched.
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

// One staged window: element m (m < n) of the window is S[m + par].
struct Window {
  const float2 *src = nullptr;  // first element copied (16-byte aligned); null: not staged
  int par = 0;
  uint32_t bytes = 0;
};

// window of n contiguous elements base[start .. start + n) of an array whose elements [lo, hi) may be read
// (n <= capacity of S minus 2)
__device__ __forceinline__ Window make_window(const float2 *base, long long lo, long long hi, long long start, int n) {
  Window w;
  const float2 *p = base + start;
  const int par = (int)((reinterpret_cast<uintptr_t>(p) >> 3) & 1);
  const int tail = (int)((reinterpret_cast<uintptr_t>(p + n) >> 3) & 1);
  if (n <= 0 || (par && start <= lo) || (tail && start + n >= hi)) return w;
  w.src = p - par;
  w.par = par;
  w.bytes = (uint32_t)(n + par + tail) * 8u;
  return w;
}
__device__ __forceinline__ Window make_window(const float2 *base, uint32_t N, uint32_t start, int n) {
  return make_window(base, 0ll, (long long)N, (long long)start, n);
}

// one thread: start the copy into S (16-byte aligned) and arm the barrier with its byte count (0: not staged)
__device__ __forceinline__ void issue(const Window &w, float2 *S, uint64_t *bar) {
  mbar_expect_tx(bar, w.bytes);
  if (w.bytes) bulk_g2s(S, w.src, w.bytes, bar);
}

__device__ __forceinline__ float2 read(const Window &w, const float2 *S, int m) { return S[m + w.par]; }

}  // namespace tma
}  // namespace b2
