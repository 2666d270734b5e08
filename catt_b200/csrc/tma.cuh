// tma.cuh — 1-D bulk asynchronous copy (TMA, cp.async.bulk) global -> shared with an mbarrier, used to stage the
// V/J reference text into shared memory once per CTA.  SASS: UBLKCP + SYNCS.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace catt {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bytes must be a multiple of 16; both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}

// Stage `bytes` (rounded up to 16) from global to shared by one elected thread; every thread of the CTA must call.
// Chunks of <= 32 KiB keep each transaction count well inside the mbarrier tx-count range.
__device__ __forceinline__ void stage_to_smem(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  const uint32_t total = (bytes + 15u) & ~15u;
  if (threadIdx.x == 0) mbar_init(bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, total);
    for (uint32_t o = 0; o < total; o += 32768u) {
      uint32_t n = total - o < 32768u ? total - o : 32768u;
      bulk_g2s((char*)dst_smem + o, (const char*)src_gmem + o, n, bar);
    }
  }
  mbar_wait(bar, 0);
}

}  // namespace catt
