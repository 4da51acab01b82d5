// Asynchronous-copy and cluster helpers (sm_100a PTX) shared by the IRLBA / normalisation kernels:
// mbarrier transactions, 1-D bulk copies in both directions (UBLKCP), proxy fences, cluster barriers and
// distributed-shared-memory addressing.
#pragma once
#include <stdint.h>

__device__ __forceinline__ uint32_t ab_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ab_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ab_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void ab_mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void ab_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ab_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ab_mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ab_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool ab_mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(ab_smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a transaction that never completes traps instead of hanging the GPU
__device__ __forceinline__ void ab_mbar_wait(uint64_t *bar, uint32_t parity) {
    for (uint32_t spins = 0; !ab_mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 28)) __trap();
}
// global -> shared, completion counted in bytes on an mbarrier; src, dst and bytes multiples of 16
__device__ __forceinline__ void ab_bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(ab_smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(ab_smem_u32(bar))
                 : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void ab_bulk_s2g(void *gdst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gdst), "r"(ab_smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void ab_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all of this thread's bulk stores have finished READING their shared-memory source
__device__ __forceinline__ void ab_bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void ab_bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (bulk stores that read them)
__device__ __forceinline__ void ab_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- thread-block clusters ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ab_cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void ab_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void ab_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void ab_cluster_sync() { ab_cluster_arrive(); ab_cluster_wait(); }
// address of the same shared-memory variable in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t ab_mapa(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void ab_st_cluster_f64(uint32_t addr, double v) {
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
// remote store that completes `8 bytes` of the transaction count of an mbarrier in the SAME remote CTA: the receiver
// learns of the data by waiting on its own barrier (no cluster-wide barrier, no fence)
__device__ __forceinline__ void ab_st_async_f64(uint32_t remote_addr, double v, uint32_t remote_mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];"
                 ::"r"(remote_addr), "d"(v), "r"(remote_mbar) : "memory");
}
__device__ __forceinline__ double ab_ld_cluster_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}

// ---- peer (NVLink) flags: system-scope release / acquire ---------------------------------------------------
__device__ __forceinline__ void ab_st_release_sys_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void ab_st_relaxed_sys_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ab_ld_acquire_sys_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
