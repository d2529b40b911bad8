// fcx_ptx.cuh -- thin inline-PTX wrappers for the sm_100a async-copy machinery
// used by the tile pipeline (fcx_tile.cuh): mbarrier transaction barriers and
// 1-D bulk async copies (the non-tensor TMA path; SASS: UBLKCP / SYNCS).
#pragma once
#include <cstdint>

namespace fcx {

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals)
                 : "memory");
}

// Make the barrier initialisation visible to the async proxy before any bulk
// copy signals it.
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// One arrival + announce `bytes` of pending async-copy traffic.
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}

// global -> shared bulk copy; completion is signalled on `bar` as `bytes`
// transaction bytes.  dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                         uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(smem_dst)),
        "l"(__cvta_generic_to_global(gmem_src)), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// L2 eviction-priority policy for data that is touched exactly once.
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

__device__ __forceinline__ void bulk_g2s_hint(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                              uint64_t *bar, uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(smem_dst)),
        "l"(__cvta_generic_to_global(gmem_src)), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

__device__ __forceinline__ void bulk_s2g_hint(void *gmem_dst, const void *smem_src, uint32_t bytes,
                                              uint64_t policy)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::
                     "l"(__cvta_generic_to_global(gmem_dst)),
                 "r"(smem_u32(smem_src)), "r"(bytes), "l"(policy)
                 : "memory");
}

// shared -> global bulk copy, tracked by the issuing thread's bulk async-group.
__device__ __forceinline__ void bulk_s2g(void *gmem_dst, const void *smem_src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(
                     __cvta_generic_to_global(gmem_dst)),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }

// Wait until all of this thread's bulk groups have finished READING shared
// memory (the global writes may still be in flight).
__device__ __forceinline__ void bulk_wait_read_all()
{
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// Same, but the most recently committed group may still be reading (used when
// that group's source is a never-modified constant block).
__device__ __forceinline__ void bulk_wait_read_1()
{
    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}

// Order generic-proxy shared-memory writes before async-proxy reads of them.
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// Per-thread 8-byte async copy global -> shared (SASS LDGSTS): no register staging, no scoreboard
// stall; completion is tracked by the issuing thread's cp.async groups.  L1-allocating (.ca): the
// FEM kernels gather nodal values that neighbouring cells of the same tile share.
__device__ __forceinline__ void cp_async_8(void *smem_dst, const void *gmem_src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)),
                 "l"(__cvta_generic_to_global(gmem_src))
                 : "memory");
}

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }

// Wait until at most N of this thread's most recent cp.async groups are still pending.
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Streaming 16-byte global store (two doubles), no L1 allocation.
__device__ __forceinline__ void st_stream_v2(double *p, double a, double b)
{
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b)
                 : "memory");
}

}  // namespace fcx
