// scan.cuh -- K2: the per-tile instance counts become tile segments (= the tile ranges) and num_rendered.
// One CTA does it: either the LAST CTA of the preprocess kernel to finish (no extra launch; preprocess.cu) or, for an empty
// map, the stand-alone kernel in binning.cu.  Replaces cub::DeviceScan::InclusiveSum over P values + the blocking D2H copy of
// num_rendered + identifyTileRanges (rasterizer_impl.cu:280-285, :117-139).
#pragma once
#include "common.cuh"

namespace gsb {


// Thread t owns the consecutive tiles [t * per, (t + 1) * per): one pass of loads, a block-wide exclusive scan of the
// per-thread sums, one pass of stores (a single round of memory latency).  Must be called by all THREADS threads of the CTA.
// SCAN_MAX_PER: tiles per thread kept in registers between the two passes; larger images re-read the counts.
template <int THREADS, int SCAN_MAX_PER>
__device__ __forceinline__ void tile_scan_body(const uint32_t* __restrict__ tile_count, uint2* __restrict__ ranges,
                                               uint32_t* __restrict__ cursor, int tiles, GeomHeader* __restrict__ hdr,
                                               uint32_t capacity, int P)
{
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_maxlen;
    if (threadIdx.x == 0) s_maxlen = 0;
    const int per = (tiles + THREADS - 1) / THREADS;
    const int t0 = threadIdx.x * per;
    auto load = [&](int i) { return __ldcg(reinterpret_cast<const uint2*>(tile_count + (size_t)i * TILE_CTR_STRIDE)); };
    uint2 cnt[SCAN_MAX_PER];
    uint32_t sum = 0, maxlen = 0;
    if (per <= SCAN_MAX_PER) {
#pragma unroll
        for (int k = 0; k < SCAN_MAX_PER; k++) cnt[k] = (k < per && t0 + k < tiles) ? load(t0 + k) : make_uint2(0u, 0u);
#pragma unroll
        for (int k = 0; k < SCAN_MAX_PER; k++) {
            const uint32_t x = cnt[k].x + cnt[k].y;   // the tile's instance count (word 1 is unused)
            sum += x;
            maxlen = max(maxlen, x);
        }
    } else {
        for (int k = 0; k < per && t0 + k < tiles; k++) {
            const uint2 c = load(t0 + k);
            sum += c.x + c.y;
            maxlen = max(maxlen, c.x + c.y);
        }
    }
    uint32_t v = sum;   // block-wide inclusive scan of `sum`
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane_id() >= (uint32_t)o) v += u;
    }
    if (lane_id() == 31) s_warp[threadIdx.x >> 5] = v;
    maxlen = __reduce_max_sync(0xffffffffu, maxlen);
    __syncthreads();
    if (lane_id() == 0) atomicMax(&s_maxlen, maxlen);
    if (threadIdx.x < 32) {
        uint32_t w = threadIdx.x < THREADS / 32 ? s_warp[threadIdx.x] : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, w, o);
            if (lane_id() >= (uint32_t)o) w += u;
        }
        s_warp[threadIdx.x] = w;
    }
    __syncthreads();
    uint32_t start = ((threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : 0u) + v - sum;
    auto emit = [&](int i, uint2 c) {
        const uint32_t x = c.x + c.y;
        // a too-small binning capacity truncates the tail of the tile-major list (overflow is latched below)
        ranges[i] = make_uint2(min(start, capacity), min(start + x, capacity));
        cursor[(size_t)i * TILE_CTR_STRIDE] = start;   // the fallback `duplicate` pass claims the segment's slots from its start
        start += x;
    };
    if (per <= SCAN_MAX_PER) {
#pragma unroll
        for (int k = 0; k < SCAN_MAX_PER; k++)
            if (k < per && t0 + k < tiles) emit(t0 + k, cnt[k]);
    } else {
        for (int k = 0; k < per && t0 + k < tiles; k++) emit(t0 + k, load(t0 + k));
    }
    if (threadIdx.x == THREADS - 1) {
        const uint32_t total = s_warp[THREADS / 32 - 1];
        hdr->max_tile_len = s_maxlen;
        hdr->magic = GEOM_MAGIC;
        hdr->P = P;
        hdr->num_rendered = total;
        hdr->num_rendered_clamped = min(total, capacity);
        hdr->overflow = total > capacity ? 1u : 0u;
        hdr->capacity = capacity;
    }
}

}  // namespace gsb
