// exchange.cu -- the ONE exchange step of the sharded render-optimise loop, as a
// single sm_100a kernel over NVLink 5 / NVSwitch peer memory.
//
// The reference is single-GPU; the sharded loop (keyframe-batch or tile-row shard) needs one sum
// all-reduce of the packed [14, P] per-Gaussian gradient block between loss.backward() and the
// Adam step (src/Render.cc:471-475).  The block lives at the same offset of a symmetric
// (peer-mapped) allocation on every rank.  One launch does:
//
//   1. per-CTA pairwise barrier: CTA b of rank r handshakes with CTA b of every other rank
//      (release/acquire CAS on a symmetric u32 scratch).  A CTA of this kernel can only run once
//      everything enqueued before it on its rank's stream (the backward kernels that wrote the
//      local gradients) is complete, so the handshake proves every rank's block is final;
//   2. rank r reduces ITS 1/world slice: with an NVSwitch multicast mapping ONE
//      multimem.ld_reduce.add.v4.f32 per 16 bytes lets the switch add the world copies in flight;
//      without multicast the world copies are read with 128-bit peer loads in rank order;
//   3. the reduced slice is published to every rank (multimem.st, or world peer stores) -- each
//      element is computed by exactly one rank, so all ranks end up bit-identical (replicated
//      parameters stay replicated);
//   4. system fence + the same pairwise barrier again: when the kernel completes on a rank, every
//      CTA of every rank has delivered its part of that rank's block.
//
// NVLink traffic per rank and direction: n * 4 bytes (+ 1/world) with multicast -- every rank's copy
// travels to the switch once and the reduced block comes back once (56 + 7 MB at 1 M Gaussians on
// 8 GPUs) -- against 2 * (world-1)/world * n * 4 for a ring or P2P all-reduce (98 MB).
#include <cstdlib>
#include "common.cuh"

namespace gsb {

constexpr int XCH_MAX_WORLD = 16;
constexpr int XCH_MAX_CTAS = 8 * NUM_SMS;   // handshake slots (all CTAs of a launch are co-resident: a CTA spins on its peers)
constexpr int XCH_CTAS = 2 * NUM_SMS;
constexpr int XCH_THREADS = 256;

struct XchPtrs {
    float* peer[XCH_MAX_WORLD];      // this rank's mapping of every rank's block (own included)
    uint32_t* sync[XCH_MAX_WORLD];   // ... and of every rank's handshake scratch
};

constexpr unsigned long long XCH_TIMEOUT_NS = 60ull * 1000 * 1000 * 1000;   // 60 s of rank skew, then a device-side trap
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint32_t cas_release_sys(uint32_t* p, uint32_t cmp, uint32_t val)
{
    uint32_t old;
    asm volatile("atom.release.sys.global.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(p), "r"(cmp), "r"(val) : "memory");
    return old;
}
__device__ __forceinline__ uint32_t cas_acquire_sys(uint32_t* p, uint32_t cmp, uint32_t val)
{
    uint32_t old;
    asm volatile("atom.acquire.sys.global.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(p), "r"(cmp), "r"(val) : "memory");
    return old;
}

// CTA b of this rank <-> CTA b of every other rank.  Slot layout of the scratch: [cta][sender rank].
__device__ __forceinline__ void pairwise_barrier(const XchPtrs& X, int rank, int world)
{
    __syncthreads();
    if ((int)threadIdx.x < world && (int)threadIdx.x != rank) {
        const int peer = threadIdx.x;
        uint32_t* theirs = X.sync[peer] + (size_t)blockIdx.x * XCH_MAX_WORLD + rank;   // my flag in the peer's scratch
        uint32_t* mine = X.sync[rank] + (size_t)blockIdx.x * XCH_MAX_WORLD + peer;     // the peer's flag in my scratch
        // bounded spins: a rank that never launches (it failed earlier) must surface as a CUDA error here, not hang the box
        const unsigned long long t0 = globaltimer_ns();
        while (cas_release_sys(theirs, 0u, 1u) != 0u)      // put (waits until the previous signal was consumed)
            if (globaltimer_ns() - t0 > XCH_TIMEOUT_NS) __trap();
        while (cas_acquire_sys(mine, 1u, 0u) != 1u)        // wait and consume
            if (globaltimer_ns() - t0 > XCH_TIMEOUT_NS) __trap();
    }
    __syncthreads();
}

template <bool MULTIMEM, int XCH_UNROLL>
__global__ void __launch_bounds__(XCH_THREADS)
exchange_allreduce_kernel(XchPtrs X, float* __restrict__ mc, long long n4, int rank, int world)
{
    pairwise_barrier(X, rank, world);
    const long long lo = n4 * rank / world, hi = n4 * (rank + 1) / world;   // this rank's slice, in float4 units
    // XCH_UNROLL independent 16-byte transactions per thread and pass: NVLink needs megabytes in flight
    const long long stride = (long long)gridDim.x * XCH_THREADS;
    for (long long i0 = lo + (long long)blockIdx.x * XCH_THREADS + threadIdx.x; i0 < hi; i0 += stride * XCH_UNROLL) {
        float4 v[XCH_UNROLL];
        if (MULTIMEM) {
#pragma unroll
            for (int u = 0; u < XCH_UNROLL; u++) {
                const long long i = i0 + u * stride;
                if (i < hi)
                    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(mc + 4 * i) : "memory");
            }
#pragma unroll
            for (int u = 0; u < XCH_UNROLL; u++) {
                const long long i = i0 + u * stride;
                if (i < hi)
                    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                                 :: "l"(mc + 4 * i), "f"(v[u].x), "f"(v[u].y), "f"(v[u].z), "f"(v[u].w) : "memory");
            }
        } else {
#pragma unroll
            for (int u = 0; u < XCH_UNROLL; u++) v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int p = 0; p < world; p++) {   // fixed order: the sum does not depend on which rank computes it
                float4 t[XCH_UNROLL];
#pragma unroll
                for (int u = 0; u < XCH_UNROLL; u++) {
                    const long long i = i0 + u * stride;
                    t[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (i < hi)   // ld.volatile: never served from this SM's L1
                        asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                                     : "=f"(t[u].x), "=f"(t[u].y), "=f"(t[u].z), "=f"(t[u].w) : "l"(X.peer[p] + 4 * i) : "memory");
                }
#pragma unroll
                for (int u = 0; u < XCH_UNROLL; u++) { v[u].x += t[u].x; v[u].y += t[u].y; v[u].z += t[u].z; v[u].w += t[u].w; }
            }
            for (int p = 0; p < world; p++) {
#pragma unroll
                for (int u = 0; u < XCH_UNROLL; u++) {
                    const long long i = i0 + u * stride;
                    if (i < hi) *reinterpret_cast<float4*>(X.peer[p] + 4 * i) = v[u];
                }
            }
        }
    }
    __threadfence_system();
    pairwise_barrier(X, rank, world);
}

}  // namespace gsb

using namespace gsb;

extern "C" {

size_t gsb_exchange_sync_bytes(int world)
{
    (void)world;
    return (size_t)XCH_MAX_CTAS * XCH_MAX_WORLD * sizeof(uint32_t);
}

int gsb_exchange_allreduce(void* multicast_ptr, void* const* peer_ptrs, void* const* sync_ptrs, long long n, int rank,
                           int world, gsb_stream_t stream)
{
    if (world < 1 || world > XCH_MAX_WORLD || rank < 0 || rank >= world || n < 0 || (n & 3)) {
        set_error("exchange_allreduce: need 1 <= world <= %d, 0 <= rank < world, n %% 4 == 0 (got world %d rank %d n %lld)",
                  XCH_MAX_WORLD, world, rank, n);
        return GSB_ERR_INVALID_ARGUMENT;
    }
    if (world == 1 || n == 0) return GSB_OK;
    if (!peer_ptrs || !sync_ptrs) {
        set_error("exchange_allreduce: peer_ptrs / sync_ptrs are required");
        return GSB_ERR_INVALID_ARGUMENT;
    }
    XchPtrs X;
    for (int p = 0; p < XCH_MAX_WORLD; p++) {
        X.peer[p] = p < world ? static_cast<float*>(peer_ptrs[p]) : nullptr;
        X.sync[p] = p < world ? static_cast<uint32_t*>(sync_ptrs[p]) : nullptr;
        if (p < world && (!X.peer[p] || !X.sync[p] || (reinterpret_cast<uintptr_t>(X.peer[p]) & 15))) {
            set_error("exchange_allreduce: peer / sync pointer of rank %d is NULL or not 16-byte aligned", p);
            return GSB_ERR_INVALID_ARGUMENT;
        }
    }
    cudaStream_t s = (cudaStream_t)stream;
    {
        StageTimer _t(ST_OTHER, s);
        int ctas = XCH_CTAS, unroll = 4;
#ifdef GSB_TUNING   // developer builds only
        const char* e1 = getenv("GSB_XCH_CTAS_PER_SM");   // read on every call: a probe sweeps them inside one process
        const char* e2 = getenv("GSB_XCH_UNROLL");
        const int ctas_env = e1 ? atoi(e1) : 2, unroll_env = e2 ? atoi(e2) : 4;
        ctas = (ctas_env < 1 ? 1 : ctas_env > 8 ? 8 : ctas_env) * NUM_SMS;
        unroll = unroll_env;
#endif
        float* mc = static_cast<float*>(multicast_ptr);
#define GSB_XCH(MM, U) exchange_allreduce_kernel<MM, U><<<ctas, XCH_THREADS, 0, s>>>(X, mc, n / 4, rank, world)
        if (multicast_ptr) {
            if (unroll == 2) GSB_XCH(true, 2); else if (unroll == 8) GSB_XCH(true, 8); else GSB_XCH(true, 4);
        } else {
            if (unroll == 2) GSB_XCH(false, 2); else if (unroll == 8) GSB_XCH(false, 8); else GSB_XCH(false, 4);
        }
#undef GSB_XCH
        GSB_LAUNCH_CHECK();
    }
    return GSB_OK;
}

}  // extern "C"
