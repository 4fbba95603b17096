// Bulk asynchronous copies (TMA engine, SASS UBLKCP) for contiguous spans: cp.async.bulk between
// global and shared memory, completion through an mbarrier (loads) or a bulk group (stores).
// Used where a warp owns one contiguous span of rows (the [32, 75] SH coefficient slab): one
// instruction moves the whole span, no register staging, and the warp computes while it lands.
// Requirements: 16-byte aligned addresses, size a multiple of 16 bytes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gg {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    // make the initialisation visible to the async proxy before a bulk copy signals the barrier
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}

// one elected thread: arm the barrier with the byte count, then launch the copy
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// one elected thread, after the whole warp's shared-memory writes are done (__syncwarp): publish
// them to the async proxy, store the span, and wait until shared memory may be reused
__device__ __forceinline__ void bulk_store_and_wait(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
}

}  // namespace gg
