// stage.cuh -- shared-memory staging of a tile's splat records for the blend kernels.
//
// The blend kernels walk a tile's depth-sorted id list in batches of one entry per thread.  For every batch
// each thread gathers ONE 48-byte SplatRec (three 16-byte asynchronous global->shared
// copies, LDGSTS, no register staging) into an NS-deep ring of SoA-of-float4 buffers:
// a[] = {x, y, half2 footprint, power threshold}, b[] = {conic.x, conic.y, conic.z, opacity},
// c[] = {r, g, b, depth}.  Pipeline: batches 0..NS-2 are issued in the prologue; iteration b
// waits for its own copies of batch b, passes the ONE CTA-wide barrier of the batch (which
// publishes batch b and proves everyone is finished with batch b-1), issues batch b+NS-1
// into the buffer batch b-1 used, and processes batch b.  In the inner loops the lanes of a quarter-warp read the SAME
// entry (a broadcast), and the per-lane cull pass reads a[j] at a 16-byte stride.
#pragma once
#include "common.cuh"

namespace gsb {

// A blend CTA covers a whole 16x16 tile (HALVES = 1: 256 threads) or its upper / lower half
// (HALVES = 2: 128 threads, twice as many CTAs: finer load balance over the 148 SMs and a
// barrier that couples 4 warps instead of 8).  The batch size equals the CTA size.
template <int NS, int BATCH>
struct StageBuf {
    float4 a[NS][BATCH];
    float4 b[NS][BATCH];
    float4 c[NS][BATCH];
};  // 48 B per entry and stage

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem)
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// Issue this thread's gather for one batch (id == 0xffffffff: nothing to fetch).  The caller
// commits the group -- every thread commits exactly one group per batch so wait_group counts
// line up.
template <int NS, int BATCH>
__device__ __forceinline__ void stage_issue(StageBuf<NS, BATCH>& S, int buf, const SplatRec* __restrict__ rec, uint32_t id)
{
    if (id != 0xffffffffu) {
        const SplatRec* r = rec + id;
        cp_async16(&S.a[buf][threadIdx.x], &r->a);
        cp_async16(&S.b[buf][threadIdx.x], &r->b);
        cp_async16(&S.c[buf][threadIdx.x], &r->c);
    }
}

__device__ __forceinline__ float rcp_approx(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Address of the hit word of (window w, block) of a tile whose list is [start, start + len)
// (see BinningLayout::hits / ImageLayout::hits_tail).
__device__ __forceinline__ uint32_t* hit_word(uint32_t* hits_full, uint32_t* hits_tail, uint32_t tile, uint32_t start,
                                              uint32_t len, uint32_t w, uint32_t block)
{
    return w < (len >> 5) ? hits_full + ((size_t)(start >> 5) + w) * HIT_BLOCKS + block
                          : hits_tail + (size_t)tile * HIT_BLOCKS + block;
}

}  // namespace gsb
