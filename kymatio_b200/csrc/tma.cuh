// tma.cuh - thin wrappers over the sm_90+/sm_100a asynchronous bulk-copy (TMA) and mbarrier PTX used by the streaming
// slab kernels (kernels2d_tma.cuh): 1-D bulk copies global <-> shared (SASS UBLKCP), transaction barriers
// (SYNCS.ARRIVE.TRANS64 / SYNCS.PHASECHK) and the generic->async proxy fence.
//
// Usage pattern (one producer thread, N consumer threads, two buffers):
//   producer:  mbar_arrive_expect_tx(full[b], bytes);  bulk_load(buf[b] + ..., src, n, full[b]) ...      (loads)
//   consumers: mbar_wait(full[b], parity);  ... compute in buf[b] ...;  fence_proxy_async();  barrier;
//              one thread: mbar_arrive(done[b])
//   producer:  mbar_wait(done[b], parity);  bulk_store(dst, buf[b] + ..., n) ...;  bulk_commit();
//              bulk_wait_read<0>();   (the stores have finished READING shared memory: buf[b] may be refilled)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sb {
namespace tma {

__device__ __forceinline__ uint32_t saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrive_count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(saddr(bar)), "r"(arrive_count) : "memory");
}
// make the barrier initialisation visible to the async proxy (the TMA unit) before the first copy is issued
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(saddr(bar)) : "memory");
}
// one arrival + `bytes` expected from bulk copies that complete on this barrier
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(saddr(bar)), "r"(bytes) : "memory");
}
// wait until the phase with the given parity has completed (try_wait suspends the thread in hardware up to a time limit)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "SB_MBAR_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n\t"
        "@p bra SB_MBAR_DONE;\n\t"
        "bra SB_MBAR_WAIT;\n\t"
        "SB_MBAR_DONE:\n\t"
        "}" ::"r"(saddr(bar)), "r"(parity)
        : "memory");
}

// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completion is signalled on `bar`
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(saddr(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(saddr(bar))
                 : "memory");
}
// shared -> global (bulk async-group completion)
__device__ __forceinline__ void bulk_store(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(saddr(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of this thread's committed bulk groups are still reading their shared-memory source
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
// ... until at most N groups are still in flight at all (global writes performed)
template <int N> __device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
// order this thread's generic-proxy shared-memory writes before subsequent async-proxy (TMA) reads of them
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// named barrier over a subset of the CTA's warps (id 1..15; `nthreads` a multiple of 32)
template <int ID> __device__ __forceinline__ void named_sync(int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "r"(nthreads) : "memory");
}

}  // namespace tma
}  // namespace sb
