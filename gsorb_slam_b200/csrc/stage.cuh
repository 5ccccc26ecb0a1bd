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

// default staging engine of the blend kernels (GSB_BLEND_STAGE=bulk|ldgsts overrides): see StageRing<> below
#ifndef GSB_DEFAULT_BULK
#define GSB_DEFAULT_BULK false
#endif

// A blend CTA covers a whole 16x16 tile (HALVES = 1: 256 threads) or its upper / lower half
// (HALVES = 2: 128 threads, twice as many CTAs: finer load balance over the 148 SMs and a
// barrier that couples 4 warps instead of 8).  The batch size equals the CTA size.
//
// Two staging engines behind one interface (StageRing<NS, BATCH, BULK>):
//   BULK = false  three 16-byte cp.async (LDGSTS) per record into SoA-of-float4 buffers, completion by
//                 cp.async.wait_group;
//   BULK = true   ONE 48-byte cp.async.bulk (the TMA unit's 1-D bulk copy, UBLKCP in SASS) per record into an
//                 array-of-records buffer, completion by an mbarrier transaction count (thread 0 arms the
//                 stage's barrier with 48 x entries bytes, every copy reports complete_tx on it, consumers
//                 spin on mbarrier.try_wait.parity).  The records are a GATHER (one row per list entry), which
//                 a tensor-map TMA cannot express; the 1-D bulk form can.  A 48-byte record stride keeps the
//                 per-lane 16-byte reads of the cull pass conflict free (8 lanes x 12 words cover the 32 banks).
template <int NS, int BATCH, bool BULK>
struct StageRing;

// One record of padding behind every stage: the forward walk lets a lane with nothing to pop read "entry 32" of its window and
// discard it (blend_fwd.cu); behind the last window of a stage that is this padding, not the first record of the NEXT stage,
// which cp.async may be filling at that moment (harmless, but compute-sanitizer racecheck reports it).
constexpr int RING_PAD = 1;
template <int NS, int BATCH>
struct StageRing<NS, BATCH, false> {
    float4 a[NS][BATCH + RING_PAD];
    float4 b[NS][BATCH + RING_PAD];
    float4 c[NS][BATCH + RING_PAD];
    __device__ __forceinline__ void init() {}
    __device__ __forceinline__ float4 A(int buf, int e) const { return a[buf][e]; }
    __device__ __forceinline__ float4 B(int buf, int e) const { return b[buf][e]; }
    __device__ __forceinline__ float4 C(int buf, int e) const { return c[buf][e]; }
};  // 48 B per entry and stage

template <int NS, int BATCH>
struct StageRing<NS, BATCH, true> {
    SplatRec r[NS][BATCH];
    unsigned long long bar[NS];
    __device__ __forceinline__ void init()
    {
        if (threadIdx.x == 0) {
#pragma unroll
            for (int i = 0; i < NS; i++) {
                const uint32_t m = (uint32_t)__cvta_generic_to_shared(&bar[i]);
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(m) : "memory");
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncthreads();
    }
    __device__ __forceinline__ float4 A(int buf, int e) const { return r[buf][e].a; }
    __device__ __forceinline__ float4 B(int buf, int e) const { return r[buf][e].b; }
    __device__ __forceinline__ float4 C(int buf, int e) const { return r[buf][e].c; }
};

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

// Issue this thread's gather for one batch (id == 0xffffffff: nothing to fetch; `count` = entries of the batch).
// LDGSTS engine: the caller commits the group -- every thread commits exactly one group per batch so
// wait_group counts line up.  Bulk engine: thread 0 arms the stage's mbarrier with the byte count.
template <int NS, int BATCH>
__device__ __forceinline__ void stage_issue(StageRing<NS, BATCH, false>& S, int buf, const SplatRec* __restrict__ rec, uint32_t id, int count)
{
    (void)count;
    if (id != 0xffffffffu) {
        const SplatRec* r = rec + id;
        cp_async16(&S.a[buf][threadIdx.x], &r->a);
        cp_async16(&S.b[buf][threadIdx.x], &r->b);
        cp_async16(&S.c[buf][threadIdx.x], &r->c);
    }
}
template <int NS, int BATCH>
__device__ __forceinline__ void stage_issue(StageRing<NS, BATCH, true>& S, int buf, const SplatRec* __restrict__ rec, uint32_t id, int count)
{
    const uint32_t m = (uint32_t)__cvta_generic_to_shared(&S.bar[buf]);
    if (threadIdx.x == 0 && count > 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(m), "r"((uint32_t)count * (uint32_t)sizeof(SplatRec)) : "memory");
    if (id != 0xffffffffu) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&S.r[buf][threadIdx.x]);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 48, [%2];"
                     ::"r"(dst), "l"(rec + id), "r"(m) : "memory");
    }
}
// Wait until batch number `batch` (0-based, staged into buffer batch % NS) has landed for THIS thread's view.
template <int NS, int BATCH>
__device__ __forceinline__ void stage_wait(StageRing<NS, BATCH, false>&, int, int)
{
    cp_async_wait<NS - 2>();
}
template <int NS, int BATCH>
__device__ __forceinline__ void stage_wait(StageRing<NS, BATCH, true>& S, int buf, int batch)
{
    cp_async_wait<NS - 2>();   // the backward kernel's hit words still travel by cp.async
    const uint32_t m = (uint32_t)__cvta_generic_to_shared(&S.bar[buf]);
    const uint32_t parity = (uint32_t)(batch / NS) & 1u;
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(m), "r"(parity) : "memory");
    } while (!done);
}

__device__ __forceinline__ float rcp_approx(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float ex2_approx(float x)
{
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Row of HIT_PIXELS hit words of window w of a tile whose list is [start, start + len)
// (see BinningLayout::hits / ImageLayout::hits_tail).
__device__ __forceinline__ uint32_t* hit_words(uint32_t* hits_full, uint32_t* hits_tail, uint32_t tile, uint32_t start,
                                               uint32_t len, uint32_t w)
{
    return w < (len >> 5) ? hits_full + ((size_t)(start >> 5) + w) * HIT_PIXELS : hits_tail + (size_t)tile * HIT_PIXELS;
}

}  // namespace gsb
