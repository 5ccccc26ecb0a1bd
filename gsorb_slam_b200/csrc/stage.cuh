// stage.cuh -- shared-memory staging of a tile's splat records for the blend kernels.
//
// The blend kernels walk a tile's depth-sorted id list in batches of 256.  For every batch
// each thread gathers ONE 48-byte SplatRec (three 16-byte asynchronous global->shared
// copies, LDGSTS, no register staging) into a double-buffered SoA-of-float4 layout:
// a[] = {x, y, conic.x, conic.y}, b[] = {conic.z, opacity, power threshold, depth},
// c[] = {r, g, b, half2 footprint}.  In the inner loops every lane of a warp reads the SAME
// entry (a broadcast, conflict-free), and the per-lane cull pass reads a[j].xy / c[j].w at a
// 16-byte stride (2-way conflicts at worst).
#pragma once
#include "common.cuh"

namespace gsb {

constexpr int BLEND_THREADS = 256;
constexpr int BLEND_BATCH = 256;

struct StageBuf {
    float4 a[2][BLEND_BATCH];
    float4 b[2][BLEND_BATCH];
    float4 c[2][BLEND_BATCH];
};  // 24 KB

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// Issue this thread's gather for one batch (id == 0xffffffff: nothing to fetch) and commit
// the group -- every thread commits exactly one group per batch so wait_group counts line up.
__device__ __forceinline__ void stage_issue(StageBuf& S, int buf, const SplatRec* __restrict__ rec, uint32_t id)
{
    if (id != 0xffffffffu) {
        const SplatRec* r = rec + id;
        cp_async16(&S.a[buf][threadIdx.x], &r->a);
        cp_async16(&S.b[buf][threadIdx.x], &r->b);
        cp_async16(&S.c[buf][threadIdx.x], &r->c);
    }
    cp_async_commit();
}

}  // namespace gsb
