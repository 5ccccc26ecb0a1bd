// binning.cu -- tile binning and depth ordering (K2-K5) for sm_100a, plus the library's
// stand-alone onesweep radix sort (used by knn.cu).
//
// Replaces (reference file:line, behaviour only):
//   cub::DeviceScan::InclusiveSum + D2H copy  rasterizer_impl.cu:280-285
//   duplicateWithKeys                         rasterizer_impl.cu:71-112
//   cub::DeviceRadixSort::SortPairs           rasterizer_impl.cu:307-315  (6 global passes over
//                                             (tile << 32 | depth bits, id) pairs at 640x480)
//   identifyTileRanges (+ memset)             rasterizer_impl.cu:117-139, :317
//
// B200 design: the reference sorts ONE global list by a 43-45 bit key (six read+write passes
// over 24 B per instance).  Here the tile is known when an instance is created, so
//   1. preprocess counts instances per tile (atomics on T counters),
//   2. one CTA scans the T counts -> every tile's [start, end) segment = the tile ranges
//      (identifyTileRanges disappears) and num_rendered,
//   3. duplicate writes (depth bits << 32 | id) records into the segments -- bucketed by tile, arbitrary order inside
//      a tile; Gaussians touching <= 4 tiles use the slots their counting atomics returned (no second atomic round),
//      larger ones claim slots in the tail of each segment,
//   4. one CTA per tile sorts its segment on ONE 32-bit key per record -- the depth rebased to the tile's minimum and
//      quantised to the bits left above the record's index in the segment -- with a register / shuffle / shared-memory
//      bitonic network (VIMNMX comparators), then repairs the few neighbours whose quantised depths collide by an
//      odd-even transposition on the exact 64-bit records, and writes the ids out.  Size classes: n <= 4096 (256
//      threads, 16 KB), n <= 16384 (1024 threads, 64 KB, one persistent CTA per SM), longer: a 64-bit network in
//      global memory (pathological inputs only).
// Ordering by (depth bits, id) is exactly the order the reference's stable radix sort produces (instances are emitted
// in ascending id, so ties keep id order), hence the sorted list is bit-identical -- while every instance is written
// twice and read twice instead of seven times.
#include <cstdlib>
#include "common.cuh"
#include "scan.cuh"

#ifndef GSB_DEFAULT_TILE_RADIX
#define GSB_DEFAULT_TILE_RADIX false
#endif

namespace gsb {

constexpr int DUP_THREADS = 256;

// ---- K2: the stand-alone scan launch (only an EMPTY map takes it: otherwise the last preprocess CTA scans, scan.cuh) ----
constexpr int SCAN_THREADS = 1024;
__global__ void __launch_bounds__(SCAN_THREADS)
tile_scan_kernel(const uint32_t* __restrict__ tile_count, uint2* __restrict__ ranges, uint32_t* __restrict__ cursor,
                 int tiles, GeomHeader* __restrict__ hdr, uint32_t capacity, int P)
{
    tile_scan_body<SCAN_THREADS, 16>(tile_count, ranges, cursor, tiles, hdr, capacity, P);
}

// ---- K3 (fallback): one thread per Gaussian writes its instances into the tile segments -----------
// Normally there is no K3: preprocess writes every record into its tile's bucket (ImageLayout::buckets) and the tile sort reads
// it there.  Only a frame whose longest tile list exceeds BUCKET_CAP takes this pass: every thread returns at once otherwise.
// All instances claim their slots with cursor atomics (the segment starts come from the scan).
__global__ void __launch_bounds__(DUP_THREADS)
duplicate_kernel(int P, const SplatRec* __restrict__ rec, const int* __restrict__ radii, const uint32_t* __restrict__ tiles_touched,
                 uint32_t* __restrict__ cursor, const GeomHeader* __restrict__ hdr,
                 uint64_t* __restrict__ pairs, int tiles_x, int tiles_y, int band_y0, int band_y1)
{
    if (hdr->max_tile_len <= (uint32_t)BUCKET_CAP) return;   // grid-uniform: the normal case costs one small launch
    const uint32_t cap = hdr->num_rendered_clamped;
    for (int idx = blockIdx.x * DUP_THREADS + threadIdx.x; idx < P; idx += gridDim.x * DUP_THREADS) {   // a few CTAs per SM, grid-stride
        // records of culled Gaussians are stale bytes inside the blob: harmless to read, never used
        const int radius = radii[idx];
        if (radius <= 0 || tiles_touched[idx] == 0) continue;   // (outside this rank's band: no record was written)
        const float4 a = rec[idx].a;
        const float depth = rec[idx].c.w;
        const uint64_t record = ((uint64_t)__float_as_uint(depth) << 32) | (uint32_t)idx;
        uint32_t minx, miny, maxx, maxy;
        get_rect(a.x, a.y, radius, tiles_x, tiles_y, minx, miny, maxx, maxy);
        miny = max(miny, (uint32_t)band_y0);   // the same band clamp as preprocess (tile-row shard)
        maxy = min(maxy, (uint32_t)band_y1);
        if (maxy <= miny) continue;
        for (uint32_t y = miny; y < maxy; y++)
            for (uint32_t x = minx; x < maxx; x++) {
                const uint32_t pos = atomicAdd(&cursor[(size_t)(y * (uint32_t)tiles_x + x) * TILE_CTR_STRIDE], 1u);
                if (pos < cap) pairs[pos] = record;
            }
    }
}

// ---- K4: per-tile sort -----------------------------------------------------------------------------
// Bitonic network with every comparator pointing the same way (first step of each merge is a
// "flip", i ^ (k-1)), so a virtual +inf padding above n needs no storage: comparators whose
// upper element is >= n are skipped.
template <typename F>
__device__ __forceinline__ void bitonic_network(uint32_t n, uint32_t npow2, uint32_t nthreads, F cmpxchg_and_sync)
{
    for (uint32_t k = 2; k <= npow2; k <<= 1) {
        const uint32_t half = k >> 1;
        // flip step: element i of the lower half of each k-block meets base + (k-1-i)
        for (uint32_t p = threadIdx.x; p < (npow2 >> 1); p += nthreads) {
            const uint32_t blk = p / half, off = p % half;
            const uint32_t i = blk * k + off, l = blk * k + (k - 1 - off);
            if (l < n) cmpxchg_and_sync(i, l, false);
        }
        cmpxchg_and_sync(0, 0, true);
        for (uint32_t j = half >> 1; j > 0; j >>= 1) {
            for (uint32_t p = threadIdx.x; p < (npow2 >> 1); p += nthreads) {
                const uint32_t i = ((p & ~(j - 1)) << 1) | (p & (j - 1)), l = i | j;
                if (l < n) cmpxchg_and_sync(i, l, false);
            }
            cmpxchg_and_sync(0, 0, true);
        }
    }
}

__device__ __forceinline__ uint32_t next_pow2(uint32_t n)
{
    return n <= 1 ? 1u : 1u << (32 - __clz(n - 1));
}

constexpr int TSORT_SMALL = 4096;    // entries, 16 KB of 32-bit keys
constexpr int TSORT_MID = 16384;     // entries, 64 KB
constexpr int TSORT_THREADS = 256;

// ---- 32-bit keyed variant (the common size classes, n <= 4096) ----------------------------------
// A tile's records are (depth bits << 32 | id).  Sorting 64-bit records costs a two-instruction
// compare, four selects and two shuffles per comparator.  Here every record is replaced by ONE
// 32-bit key: the depth, rebased to the tile's minimum and shifted so that the tile's depth range
// fits in 32 - log2(npad) bits, above the record's index in the unsorted segment.  The key order
// is the (depth, id) order except between records whose quantised depths collide (about two pairs
// per 2000-entry tile); those are put right afterwards by an odd-even transposition on the exact
// 64-bit records, restricted to neighbours with equal quantised depth.  A comparator is then one
// VIMNMX pair, a shuffle stage one SHFL + one VIMNMX per element.
template <int E, int THREADS>
__device__ __forceinline__ void tile_sort_regs32(uint32_t* __restrict__ s, uint32_t (&a)[E])
{
    constexpr uint32_t N = E * THREADS;
    const uint32_t t = threadIdx.x, lane = t & 31;
#pragma unroll
    for (uint32_t k = 2; k <= N; k <<= 1) {
        // ---- flip step of the merge of size k ----
        if (k <= (uint32_t)E) {
#pragma unroll
            for (int b0 = 0; b0 < E; b0 += (int)k)
#pragma unroll
                for (int o = 0; o < (int)k / 2; o++) {
                    const uint32_t x = a[b0 + o], y = a[b0 + (int)k - 1 - o];
                    a[b0 + o] = min(x, y);
                    a[b0 + (int)k - 1 - o] = max(x, y);
                }
        } else if (k <= 32u * E) {
            const uint32_t m = k / E;  // 2..32 threads per k-block
            uint32_t other[E];
#pragma unroll
            for (int r = 0; r < E; r++) other[r] = __shfl_xor_sync(0xffffffffu, a[E - 1 - r], m - 1);
            const bool lower = (lane & (m >> 1)) == 0;
#pragma unroll
            for (int r = 0; r < E; r++) a[r] = lower ? min(a[r], other[r]) : max(a[r], other[r]);
        }
        uint32_t j = k >> 2;  // first stride of the half-cleaners
        if (k > 32u * E) {
            // cross-warp part of this merge in shared memory: flip, then strides >= 32 E
            __syncthreads();
#pragma unroll
            for (int r = 0; r < E; r++) s[t * E + r] = a[r];
            __syncthreads();
            const uint32_t half = k >> 1;
            for (uint32_t p = t; p < N / 2; p += THREADS) {
                const uint32_t blk = p / half, off = p % half;
                const uint32_t i = blk * k + off, l = blk * k + (k - 1 - off);
                const uint32_t x = s[i], y = s[l];
                s[i] = min(x, y);
                s[l] = max(x, y);
            }
            __syncthreads();
            for (; j >= 32u * E; j >>= 1) {
                for (uint32_t p = t; p < N / 2; p += THREADS) {
                    const uint32_t i = ((p & ~(j - 1)) << 1) | (p & (j - 1)), l = i | j;
                    const uint32_t x = s[i], y = s[l];
                    s[i] = min(x, y);
                    s[l] = max(x, y);
                }
                __syncthreads();
            }
#pragma unroll
            for (int r = 0; r < E; r++) a[r] = s[t * E + r];
        }
        // ---- half-cleaners inside a warp (shuffles) ----
#pragma unroll
        for (uint32_t jj = 16u * E; jj >= (uint32_t)E; jj >>= 1) {
            if (jj <= j && jj >= 1) {
                const uint32_t m = jj / E;
                const bool lower = (lane & m) == 0;
#pragma unroll
                for (int r = 0; r < E; r++) {
                    const uint32_t o = __shfl_xor_sync(0xffffffffu, a[r], m);
                    a[r] = lower ? min(a[r], o) : max(a[r], o);
                }
            }
        }
        // ---- half-cleaners inside a thread (registers) ----
#pragma unroll
        for (int jj = E / 2; jj >= 1; jj >>= 1) {
            if ((uint32_t)jj <= j) {
#pragma unroll
                for (int r = 0; r < E; r++)
                    if ((r & jj) == 0) {
                        const uint32_t x = a[r], y = a[r | jj];
                        a[r] = min(x, y);
                        a[r | jj] = max(x, y);
                    }
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < E; r++) s[t * E + r] = a[r];
    __syncthreads();
}

// Alternative to the comparator network for the same 32-bit keys: a stable LSD radix sort in shared memory over the
// quantised-depth field only (three 7-bit digits; the index bits below it never need sorting: neighbours with equal
// quantised depth are put in exact order by the fix-up pass anyway).  Warp-striped ranking with match.any, as in the
// onesweep pass further down, minus the global look-back.  O(n) per pass instead of O(n log^2 n) compare-exchanges.
constexpr int TR_BITS = 7, TR_RADIX = 1 << TR_BITS, TR_PASSES = 3, TR_QBITS = TR_BITS * TR_PASSES;   // 21 depth bits
template <int E, int THREADS>
__device__ __forceinline__ void tile_radix_sort32(uint32_t* __restrict__ s, uint32_t (&a)[E], uint32_t* __restrict__ cnt /*[THREADS/32][128]*/,
                                                  uint32_t* __restrict__ dstart /*[128 + 32]*/, int idx_bits)
{
    constexpr int WARPS = THREADS / 32;
    const uint32_t t = threadIdx.x, warp = t >> 5, lane = t & 31, lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int k = 0; k < E; k++) s[t * E + k] = a[k];   // any initial order: the segment is unordered
    uint32_t* s_w = dstart + TR_RADIX;                 // per-warp totals of the digit scan
#pragma unroll 1
    for (int pass = 0; pass < TR_PASSES; pass++) {
        const int shift = idx_bits + TR_BITS * pass;
        __syncthreads();   // s[] complete; the previous pass' readers of cnt / dstart are done
        for (uint32_t i = t; i < (uint32_t)(WARPS * TR_RADIX); i += THREADS) cnt[i] = 0;
        uint32_t key[E], loc[E];
#pragma unroll
        for (int it = 0; it < E; it++) key[it] = s[warp * (32 * E) + it * 32 + lane];   // warp-striped: position order = (warp, it, lane)
        __syncthreads();   // counters zeroed, every key is in a register (s[] may be overwritten below)
#pragma unroll
        for (int it = 0; it < E; it++) {
            const uint32_t d = (key[it] >> shift) & (TR_RADIX - 1);
            const uint32_t peers = __match_any_sync(0xffffffffu, d);
            const int leader = __ffs(peers) - 1;
            uint32_t base = 0;
            if ((int)lane == leader) {
                base = cnt[warp * TR_RADIX + d];
                cnt[warp * TR_RADIX + d] = base + __popc(peers);
            }
            base = __shfl_sync(0xffffffffu, base, leader);
            loc[it] = base + __popc(peers & lt_mask);
            __syncwarp();
        }
        __syncthreads();
        // digit t: exclusive prefix over the warps, total count; then an exclusive scan of the totals over the 128 digits
        uint32_t count = 0;
        if (t < (uint32_t)TR_RADIX) {
#pragma unroll 4
            for (int w = 0; w < WARPS; w++) {
                const uint32_t c = cnt[w * TR_RADIX + t];
                cnt[w * TR_RADIX + t] = count;
                count += c;
            }
        }
        uint32_t v = count;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= (uint32_t)o) v += u;
        }
        if (lane == 31 && warp < TR_RADIX / 32) s_w[warp] = v;
        __syncthreads();
        if (t < (uint32_t)TR_RADIX) {
            uint32_t before = 0;
#pragma unroll
            for (int w = 0; w < TR_RADIX / 32; w++)
                if (w < (int)warp) before += s_w[w];
            dstart[t] = before + v - count;
        }
        __syncthreads();
#pragma unroll
        for (int it = 0; it < E; it++) {
            const uint32_t d = (key[it] >> shift) & (TR_RADIX - 1);
            s[dstart[d] + cnt[warp * TR_RADIX + d] + loc[it]] = key[it];
        }
    }
    __syncthreads();
}

// 64-bit network on the tile's own (depth bits << 32 | id) records, in place in global memory (L2): the exact, bounded last
// resort of the 32-bit keyed sort below, and the size class of lists longer than 16384 entries.
template <int THREADS>
__device__ __forceinline__ void tile_sort_exact64(uint64_t* __restrict__ g, uint32_t n, uint32_t* __restrict__ out)
{
    bitonic_network(n, next_pow2(n), THREADS, [&](uint32_t i, uint32_t l, bool sync) {
        if (sync) {
            __threadfence_block();
            __syncthreads();
            return;
        }
        const uint64_t a = g[i], b = g[l];
        if (a > b) {
            g[i] = b;
            g[l] = a;
        }
    });
    for (uint32_t i = threadIdx.x; i < n; i += THREADS) out[i] = (uint32_t)g[i];
}

// Rare path of the 32-bit keyed sort, out of line so that its registers do not weigh on the common path: s[0..n) holds the
// tile's keys (quantised depth << IDX_BITS | slot) sorted, but there are long runs of equal quantised depth.  If every run is
// a run of EXACTLY equal depths (quantised sensor depths: two surfaces in one tile, ids too spread for the exact key) the
// order inside a run is the id order, and a second pass of the same network on (position of the run's first entry,
// id - min id) finishes the job; anything else goes to the exact 64-bit network.
template <int E, int THREADS, int IDX_BITS, bool RADIX>
__device__ __noinline__ void tile_sort_long_runs(uint64_t* __restrict__ seg, uint32_t n, uint32_t* __restrict__ s,
                                                 uint32_t* __restrict__ s_red, uint32_t* __restrict__ out)
{
    constexpr uint32_t IDX_MASK = (1u << IDX_BITS) - 1u;
    const uint32_t t = threadIdx.x, lane = t & 31, warp = t >> 5;
    // (a) last run start at or before each of this thread's E positions (0 = none inside the thread)
    uint32_t start[E];
    uint32_t run = 0;
#pragma unroll
    for (int k = 0; k < E; k++) {
        const uint32_t i = t * E + k;
        if (i < n && i > 0 && (s[i - 1] >> IDX_BITS) != (s[i] >> IDX_BITS)) run = i;
        start[k] = run;
    }
    uint32_t incl = run;   // inclusive max-scan over the warp's threads
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl = max(incl, up);
    }
    uint32_t prev = __shfl_up_sync(0xffffffffu, incl, 1);   // ... exclusive
    if (lane == 0) prev = 0;
    __syncthreads();
    if (lane == 31) s_red[warp] = incl;
    // (b) are all runs pure, and do (run start, id) fit one word?
    uint32_t imn = 0xffffffffu, imx = 0u;
    for (uint32_t i = t; i < n; i += THREADS) {
        const uint32_t id = (uint32_t)seg[i];
        imn = min(imn, id);
        imx = max(imx, id);
    }
    bool impure = false;
#pragma unroll
    for (int k = 0; k < E; k++) {
        const uint32_t i = t * E + k;
        if (i < n && i > 0 && (s[i - 1] >> IDX_BITS) == (s[i] >> IDX_BITS))
            impure |= (uint32_t)(seg[s[i - 1] & IDX_MASK] >> 32) != (uint32_t)(seg[s[i] & IDX_MASK] >> 32);
    }
    imn = __reduce_min_sync(0xffffffffu, imn);
    imx = __reduce_max_sync(0xffffffffu, imx);
    if (lane == 0) {
        s_red[32 + warp] = imn;
        s_red[64 + warp] = imx;
    }
    impure = __syncthreads_or(impure);
    for (uint32_t w = 0; w < warp; w++) prev = max(prev, s_red[w]);   // last run start in the warps before this one
#pragma unroll
    for (int w = 0; w < THREADS / 32; w++) {
        imn = min(imn, s_red[32 + w]);
        imx = max(imx, s_red[64 + w]);
    }
    const int idneed = 32 - __clz(imx - imn);
    if (RADIX || impure || IDX_BITS + idneed > 32) {
        tile_sort_exact64<THREADS>(seg, n, out);
        return;
    }
    // (c) second pass
    uint32_t a[E];
#pragma unroll
    for (int k = 0; k < E; k++) {
        const uint32_t i = t * E + k;
        a[k] = i < n ? (max(start[k], prev) << idneed) | ((uint32_t)seg[s[i] & IDX_MASK] - imn) : 0xffffffffu;
    }
    tile_sort_regs32<E, THREADS>(s, a);   // (its first shared-memory write is behind a CTA barrier: every s[] read above is done)
    const uint32_t idmask = idneed >= 32 ? 0xffffffffu : (1u << idneed) - 1u;
    for (uint32_t i = t; i < n; i += THREADS) out[i] = (s[i] & idmask) + imn;
}

constexpr int TIE_ROUNDS = 8;   // odd-even rounds spent on neighbours whose quantised depths collide before the exact sort takes over

template <int E, int THREADS, bool RADIX>
__device__ __forceinline__ void tile_sort_class32(const uint2 r, uint64_t* __restrict__ seg,
                                                  uint32_t* __restrict__ point_list, uint32_t* __restrict__ s,
                                                  uint32_t* __restrict__ s_red, uint32_t* __restrict__ s_cnt)
{
    constexpr int LOG_E = E == 1 ? 0 : E == 2 ? 1 : E == 4 ? 2 : E == 8 ? 3 : 4;
    constexpr int IDX_BITS = (THREADS == 1024 ? 10 : 8) + LOG_E;
    constexpr int DEPTH_BITS = RADIX ? (32 - IDX_BITS < TR_QBITS ? 32 - IDX_BITS : TR_QBITS) : 32 - IDX_BITS;
    constexpr uint32_t IDX_MASK = (1u << IDX_BITS) - 1u;
    const uint32_t n = r.y - r.x, t = threadIdx.x;   // seg: the tile's n unsorted records (its bucket, or its segment of `pairs`)
    // depth bits of this thread's E consecutive records, tile-wide minimum and maximum
    uint32_t dep[E];
    uint32_t mn = 0xffffffffu, mx = 0u;
#pragma unroll
    for (int k = 0; k < E; k++) {
        const uint32_t i = t * E + k;
        dep[k] = i < n ? (uint32_t)(seg[i] >> 32) : 0u;
        if (i < n) {
            mn = min(mn, dep[k]);
            mx = max(mx, dep[k]);
        }
    }
    mn = __reduce_min_sync(0xffffffffu, mn);
    mx = __reduce_max_sync(0xffffffffu, mx);
    __syncthreads();   // s_red may still be read by the previous tile's threads (persistent callers)
    if ((t & 31) == 0) {
        s_red[t >> 5] = mn;
        s_red[32 + (t >> 5)] = mx;
    }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < THREADS / 32; w++) {
        mn = min(mn, s_red[w]);
        mx = max(mx, s_red[32 + w]);
    }
    // largest depth field must stay below all ones (the padding key is all ones)
    const int need = 32 - __clz(mx - mn + 1u);
    // Exact keys when (depth - min) above (id - min id) fits one word: no quantisation, no ties, no fix-up.  This is what a
    // map straight out of Render::InitWorld gives (raster-ordered ids, a few distinct QUANTISED depths per tile).  The id range
    // is only looked at when the tile's depth range is narrow enough for that to be possible (never on continuous depths).
    bool exact = false;
    uint32_t imn = 0u;
    int idneed = 32;
    if (!RADIX && need <= 20) {   // CTA-uniform
        uint32_t imx = 0u;
        imn = 0xffffffffu;
        for (uint32_t i = t; i < n; i += THREADS) {
            const uint32_t id = (uint32_t)seg[i];
            imn = min(imn, id);
            imx = max(imx, id);
        }
        imn = __reduce_min_sync(0xffffffffu, imn);
        imx = __reduce_max_sync(0xffffffffu, imx);
        if ((t & 31) == 0) {
            s_red[64 + (t >> 5)] = imn;
            s_red[96 + (t >> 5)] = imx;
        }
        __syncthreads();
#pragma unroll
        for (int w = 0; w < THREADS / 32; w++) {
            imn = min(imn, s_red[64 + w]);
            imx = max(imx, s_red[96 + w]);
        }
        idneed = 32 - __clz(imx - imn);   // bits of the id range; 0 when the tile holds one Gaussian
        exact = need + idneed <= 32;
    }
    const int shift = need > DEPTH_BITS ? need - DEPTH_BITS : 0;
    uint32_t a[E];
#pragma unroll
    for (int k = 0; k < E; k++) {
        const uint32_t i = t * E + k;
        uint32_t key = 0xffffffffu;
        if (i < n) {
            if (exact) key = ((dep[k] - mn) << idneed) | ((uint32_t)seg[i] - imn);   // idneed <= 31 here (need >= 1)
            else key = (((dep[k] - mn) >> shift) << IDX_BITS) | i;
        }
        a[k] = key;
    }
    if (RADIX) tile_radix_sort32<E, THREADS>(s, a, s_cnt, s_cnt + (THREADS / 32) * TR_RADIX, IDX_BITS);
    else tile_sort_regs32<E, THREADS>(s, a);
    if (exact) {
        const uint32_t idmask = (1u << idneed) - 1u;
        for (uint32_t i = t; i < n; i += THREADS) point_list[r.x + i] = (s[i] & idmask) + imn;
        return;
    }
    // exact order between neighbours whose quantised depths collide: odd-even transposition on the 64-bit records (runs are
    // 2-3 entries long on continuous depths) -- for at most TIE_ROUNDS rounds; a list with longer runs of equal quantised depth
    // (quantised sensor depths straddling a discontinuity, adversarial inputs) is sorted exactly by the 64-bit network instead
    bool tie = false;
#pragma unroll
    for (int k = 0; k < E; k++) {
        const uint32_t i = t * E + k;
        if (i + 1 < n) tie |= (s[i] >> IDX_BITS) == (s[i + 1] >> IDX_BITS);
    }
    if (const int tied_threads = __syncthreads_count(tie)) {
        bool unsorted = true;
        // ties all over the list (a tile of quantised sensor depths): the bounded repair cannot succeed, skip it
        for (int round = 0; round < TIE_ROUNDS && unsorted && tied_threads <= THREADS / 8; round++) {
            bool swapped = false;
#pragma unroll 1
            for (uint32_t phase = 0; phase < 2; phase++) {
                for (uint32_t i = 2 * t + phase; i + 1 < n; i += 2 * THREADS) {
                    const uint32_t x = s[i], y = s[i + 1];
                    if ((x >> IDX_BITS) == (y >> IDX_BITS) && seg[x & IDX_MASK] > seg[y & IDX_MASK]) {
                        s[i] = y;
                        s[i + 1] = x;
                        swapped = true;
                    }
                }
                __syncthreads();
            }
            unsorted = __syncthreads_or(swapped);
        }
        if (unsorted) {
            tile_sort_long_runs<E, THREADS, IDX_BITS, RADIX>(seg, n, s, s_red, point_list + r.x);
            return;
        }
    }
    for (uint32_t i = t; i < n; i += THREADS) point_list[r.x + i] = (uint32_t)seg[s[i] & IDX_MASK];
}

template <bool RADIX>
__global__ void __launch_bounds__(TSORT_THREADS, 5)
tile_sort_small_kernel(const uint2* __restrict__ ranges, uint64_t* __restrict__ pairs, uint32_t* __restrict__ point_list,
                       uint64_t* __restrict__ buckets, const GeomHeader* __restrict__ hdr)
{
    __shared__ __align__(16) uint32_t s[TSORT_SMALL];
    __shared__ uint32_t s_red[128];
    __shared__ uint32_t s_cnt[RADIX ? (TSORT_THREADS / 32) * TR_RADIX + TR_RADIX + 32 : 1];
    const uint2 r = ranges[blockIdx.x];
    const uint32_t n = r.y - r.x;
    if (n == 0 || n > (uint32_t)TSORT_SMALL) return;   // longer lists: the 128 KB / global classes
    // the tile's unsorted records: its bucket (written by preprocess) unless some list of the frame overflowed its bucket
    uint64_t* seg = hdr->max_tile_len <= (uint32_t)BUCKET_CAP ? buckets + (size_t)blockIdx.x * BUCKET_CAP : pairs + r.x;
    const uint32_t npad = n <= 256 ? 256u : next_pow2(n);
    switch (npad) {
        case 256: tile_sort_class32<1, TSORT_THREADS, false>(r, seg, point_list, s, s_red, s_cnt); break;
        case 512: tile_sort_class32<2, TSORT_THREADS, false>(r, seg, point_list, s, s_red, s_cnt); break;
        case 1024: tile_sort_class32<4, TSORT_THREADS, RADIX>(r, seg, point_list, s, s_red, s_cnt); break;
        case 2048: tile_sort_class32<8, TSORT_THREADS, RADIX>(r, seg, point_list, s, s_red, s_cnt); break;
        default: tile_sort_class32<16, TSORT_THREADS, RADIX>(r, seg, point_list, s, s_red, s_cnt); break;
    }
}

// Size class n > 4096: up to 16384 entries the same 32-bit keyed network with 1024 threads (8 or 16 keys per thread, 32 / 64 KB
// of dynamic shared memory), beyond that a 64-bit network in global memory; one persistent CTA per SM looping over the tiles of
// the class (the launch exits at once when the frame's longest list fits the 256-thread kernel).
template <bool RADIX>
__global__ void __launch_bounds__(1024, 1)
tile_sort_mid_kernel(const uint2* __restrict__ ranges, uint64_t* __restrict__ pairs, uint32_t* __restrict__ point_list, uint64_t* __restrict__ buckets,
                     int tiles, const GeomHeader* __restrict__ hdr)
{
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ uint32_t s_red[128];
    __shared__ uint32_t s_cnt[RADIX ? 32 * TR_RADIX + TR_RADIX + 32 : 1];
    if (hdr->max_tile_len <= (uint32_t)TSORT_SMALL) return;   // no tile of this size class in the frame
    uint32_t* s = reinterpret_cast<uint32_t*>(dyn_smem);
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const uint2 r = ranges[tile];
        const uint32_t n = r.y - r.x;
        if (n <= (uint32_t)TSORT_SMALL) continue;                             // the 256-thread kernel handles this tile
        __syncthreads();                                                      // previous tile's readers are done with s[]
        uint64_t* seg = hdr->max_tile_len <= (uint32_t)BUCKET_CAP ? buckets + (size_t)tile * BUCKET_CAP : pairs + r.x;
        if (n <= 8192u) tile_sort_class32<8, 1024, RADIX>(r, seg, point_list, s, s_red, s_cnt);
        else if (n <= (uint32_t)TSORT_MID) tile_sort_class32<16, 1024, RADIX>(r, seg, point_list, s, s_red, s_cnt);
        else {
            // more than 16384 entries in one tile (pathological inputs): the 64-bit network in place, in global memory
            tile_sort_exact64<1024>(pairs + r.x, n, point_list + r.x);
        }
    }
}

// ---- K4: onesweep radix sort ------------------------------------------------------------------
constexpr uint32_t LB_AGG = 1u << 30;   // look-back word: tile-local count published
constexpr uint32_t LB_INCL = 1u << 31;  // look-back word: inclusive prefix published
constexpr uint32_t LB_MASK = (1u << 30) - 1;

__global__ void __launch_bounds__(SORT_THREADS)
sort_histogram_kernel(const uint64_t* __restrict__ keys, const GeomHeader* __restrict__ hdr, uint32_t* __restrict__ hist,
                      int passes)
{
    __shared__ uint32_t s_hist[SORT_MAX_PASSES * SORT_RADIX];
    for (int i = threadIdx.x; i < passes * SORT_RADIX; i += SORT_THREADS) s_hist[i] = 0;
    __syncthreads();
    const uint32_t R = hdr->num_rendered_clamped;
    for (uint32_t base = blockIdx.x * SORT_TILE; base < R; base += gridDim.x * SORT_TILE) {
#pragma unroll 4
        for (int it = 0; it < SORT_ITEMS; it++) {
            const uint32_t i = base + it * SORT_THREADS + threadIdx.x;
            if (i < R) {
                const uint64_t k = keys[i];
                for (int p = 0; p < passes; p++)
                    atomicAdd(&s_hist[p * SORT_RADIX + (uint32_t)((k >> (p * SORT_RADIX_BITS)) & (SORT_RADIX - 1))], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * SORT_RADIX; i += SORT_THREADS)
        if (s_hist[i]) atomicAdd(&hist[i], s_hist[i]);
}

// Exclusive scan of one value per thread over the 256-thread CTA.
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t x, uint32_t* s_tmp /*[8]*/, uint32_t* total)
{
    uint32_t v = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane_id() >= (uint32_t)o) v += t;
    }
    __syncthreads();  // protect s_tmp from a previous use
    if (lane_id() == 31) s_tmp[threadIdx.x >> 5] = v;
    __syncthreads();
    uint32_t warp_excl = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < SORT_THREADS / 32; w++) {
        const uint32_t c = s_tmp[w];
        if (w < (int)(threadIdx.x >> 5)) warp_excl += c;
        tot += c;
    }
    if (total) *total = tot;
    return warp_excl + v - x;
}

struct SortSmem {
    uint64_t keys[SORT_TILE];                          // 32 KB
    uint32_t vals[SORT_TILE];                          // 16 KB
    uint32_t cnt[SORT_THREADS / 32][SORT_RADIX + 1];   // per-warp digit counters (+1 dummy slot)
    uint32_t digit_start[SORT_RADIX];                  // CTA-local start of each digit
    uint32_t out_off[SORT_RADIX];                      // global index = out_off[d] + local position
    uint32_t tmp[8];
    uint32_t tile;
};

__global__ void __launch_bounds__(SORT_THREADS)
onesweep_pass_kernel(const uint64_t* __restrict__ keys_in, uint64_t* __restrict__ keys_out,
                     const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ vals_out,
                     GeomHeader* __restrict__ hdr, const uint32_t* __restrict__ hist_pass,
                     uint32_t* __restrict__ lookback, int pass, int shift)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SortSmem& S = *reinterpret_cast<SortSmem*>(smem_raw);
    const uint32_t R = hdr->num_rendered_clamped;
    if (threadIdx.x == 0) S.tile = atomicAdd(&hdr->sort_tile_counter[pass], 1u);
    for (int i = threadIdx.x; i < (SORT_THREADS / 32) * (SORT_RADIX + 1); i += SORT_THREADS) (&S.cnt[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = S.tile;
    const uint32_t tile_base = tile * SORT_TILE;
    if (tile_base >= R) return;
    const uint32_t n_valid = min((uint32_t)SORT_TILE, R - tile_base);
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();

    // (b) warp-striped load: item `it` of warp `w` is element w*512 + it*32 + lane of the tile
    uint64_t key[SORT_ITEMS];
    uint32_t val[SORT_ITEMS];
    uint32_t loc[SORT_ITEMS];
#pragma unroll
    for (int it = 0; it < SORT_ITEMS; it++) {
        const uint32_t e = warp * (SORT_ITEMS * 32) + it * 32 + lane;
        if (e < n_valid) {
            key[it] = keys_in[tile_base + e];
            val[it] = vals_in[tile_base + e];
        } else {
            key[it] = ~0ull;
            val[it] = 0;
        }
    }
    // (c) stable rank inside the warp's 512 keys
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int it = 0; it < SORT_ITEMS; it++) {
        const uint32_t e = warp * (SORT_ITEMS * 32) + it * 32 + lane;
        const uint32_t d = e < n_valid ? (uint32_t)((key[it] >> shift) & (SORT_RADIX - 1)) : (uint32_t)SORT_RADIX;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if ((int)lane == leader) {
            base = S.cnt[warp][d];
            S.cnt[warp][d] = base + __popc(peers);
        }
        base = __shfl_sync(0xffffffffu, base, leader);
        loc[it] = base + __popc(peers & lt_mask);
        __syncwarp();
    }
    __syncthreads();
    // (d) per digit: exclusive prefix over the 8 warps, CTA count
    const uint32_t d_own = threadIdx.x;
    uint32_t count = 0;
#pragma unroll
    for (int w = 0; w < SORT_THREADS / 32; w++) {
        const uint32_t c = S.cnt[w][d_own];
        S.cnt[w][d_own] = count;
        count += c;
    }
    // (e) CTA-local digit starts and global digit bases
    const uint32_t dstart = block_excl_scan_256(count, S.tmp, nullptr);
    const uint32_t gbase = block_excl_scan_256(hist_pass[d_own], S.tmp, nullptr);
    // (f) decoupled look-back over preceding tiles for digit d_own
    volatile uint32_t* lb = lookback + (size_t)tile * SORT_RADIX + d_own;
    uint32_t excl = 0;
    if (tile == 0) {
        *lb = LB_INCL | count;
    } else {
        *lb = LB_AGG | count;
        int t = (int)tile - 1;
        while (true) {
            const uint32_t w = *(volatile uint32_t*)(lookback + (size_t)t * SORT_RADIX + d_own);
            if (w & LB_INCL) {
                excl += w & LB_MASK;
                break;
            }
            if (w & LB_AGG) {
                excl += w & LB_MASK;
                t--;
            }
        }
        *lb = LB_INCL | (excl + count);
    }
    S.digit_start[d_own] = dstart;
    S.out_off[d_own] = gbase + excl - dstart;
    __syncthreads();
    // (g) scatter into CTA-sorted order in shared memory, then stream out
#pragma unroll
    for (int it = 0; it < SORT_ITEMS; it++) {
        const uint32_t e = warp * (SORT_ITEMS * 32) + it * 32 + lane;
        if (e < n_valid) {
            const uint32_t d = (uint32_t)((key[it] >> shift) & (SORT_RADIX - 1));
            const uint32_t pos = S.digit_start[d] + S.cnt[warp][d] + loc[it];
            S.keys[pos] = key[it];
            S.vals[pos] = val[it];
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < SORT_ITEMS; it++) {
        const uint32_t pos = it * SORT_THREADS + threadIdx.x;
        if (pos < n_valid) {
            const uint64_t k = S.keys[pos];
            const uint32_t d = (uint32_t)((k >> shift) & (SORT_RADIX - 1));
            const uint32_t out = S.out_off[d] + pos;
            keys_out[out] = k;
            vals_out[out] = S.vals[pos];
        }
    }
}

// Sort hdr->num_rendered_clamped (u64 key, u32 value) pairs, stable, on the low passes*8 key
// bits.  Input in kbuf/vbuf[start]; output in kbuf/vbuf[start ^ (passes & 1)].  `hist` and
// `lookback` must be zeroed, as must hdr->sort_tile_counter[0..passes).
int launch_sort_pairs(GeomHeader* hdr, uint64_t* const kbuf[2], uint32_t* const vbuf[2], int start, int passes,
                      uint32_t* hist, uint32_t* lookback, int sort_tiles, cudaStream_t s)
{
    int cur = start;
    const int hist_grid = sort_tiles < NUM_SMS * 4 ? sort_tiles : NUM_SMS * 4;
    {
        StageTimer _t(ST_SORT_HIST, s);
        sort_histogram_kernel<<<hist_grid, SORT_THREADS, 0, s>>>(kbuf[cur], hdr, hist, passes);
        GSB_LAUNCH_CHECK();
    }
    GSB_SET_ATTR_ONCE(onesweep_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SortSmem));
    for (int pass = 0; pass < passes; pass++) {
        {
            StageTimer _t(ST_SORT_PASS, s);
            onesweep_pass_kernel<<<sort_tiles, SORT_THREADS, sizeof(SortSmem), s>>>(
                kbuf[cur], kbuf[cur ^ 1], vbuf[cur], vbuf[cur ^ 1], hdr, hist + pass * SORT_RADIX,
                lookback + (size_t)pass * sort_tiles * SORT_RADIX, pass, pass * SORT_RADIX_BITS);
            GSB_LAUNCH_CHECK();
        }
        cur ^= 1;
    }
    return GSB_OK;
}

int launch_tile_scan(char* geom, const GeomLayout& GL, char* image, const ImageLayout& IL, uint32_t capacity, int P,
                     cudaStream_t s)
{
    StageTimer _t(ST_SCAN, s);
    tile_scan_kernel<<<1, 1024, 0, s>>>(reinterpret_cast<const uint32_t*>(image + IL.tile_count),
                                        reinterpret_cast<uint2*>(image + IL.ranges),
                                        reinterpret_cast<uint32_t*>(image + IL.tile_cursor), IL.tiles_x * IL.tiles_y,
                                        reinterpret_cast<GeomHeader*>(geom + GL.header), capacity, P);
    GSB_LAUNCH_CHECK();
    return GSB_OK;
}

int launch_binning(const FwdParams& p, char* geom, const GeomLayout& GL, char* binning, const BinningLayout& BL,
                   char* image, const ImageLayout& IL, long long grid_instances, cudaStream_t s)
{
    if (grid_instances <= 0 || p.P <= 0) return GSB_OK;
    const GeomHeader* hdr = reinterpret_cast<const GeomHeader*>(geom + GL.header);
    uint64_t* pairs = reinterpret_cast<uint64_t*>(binning + BL.pairs);
    uint32_t* point_list = reinterpret_cast<uint32_t*>(binning + BL.point_list);
    const uint2* ranges = reinterpret_cast<const uint2*>(image + IL.ranges);
    uint64_t* buckets = reinterpret_cast<uint64_t*>(image + IL.buckets);
    const int tiles = p.tiles_x * p.tiles_y;
    {
        StageTimer _t(ST_DUPLICATE, s);
        duplicate_kernel<<<GL.num_blocks < 8 * NUM_SMS ? GL.num_blocks : 8 * NUM_SMS, DUP_THREADS, 0, s>>>(
            p.P, reinterpret_cast<const SplatRec*>(geom + GL.rec), reinterpret_cast<const int*>(geom + GL.radii),
            reinterpret_cast<const uint32_t*>(geom + GL.tiles_touched), reinterpret_cast<uint32_t*>(image + IL.tile_cursor), hdr, pairs, p.tiles_x, p.tiles_y, p.band_y0, p.band_y1);
        GSB_LAUNCH_CHECK();
    }
    {
        StageTimer _t(ST_TILE_SORT, s);
        // per-tile sort engine: comparator network; the shared-memory radix sort on the same 32-bit keys is a developer option
#ifdef GSB_TUNING
        static const bool radix = [] { const char* e = getenv("GSB_TILE_SORT"); return e ? e[0] == 'r' : GSB_DEFAULT_TILE_RADIX; }();
#else
        constexpr bool radix = false;
#endif
        const int g = tiles < NUM_SMS ? tiles : NUM_SMS;
#ifdef GSB_TUNING
        if (radix) {
            tile_sort_small_kernel<true><<<tiles, TSORT_THREADS, 0, s>>>(ranges, pairs, point_list, buckets, hdr);
            GSB_LAUNCH_CHECK();
            if (grid_instances > TSORT_SMALL) {
                GSB_SET_ATTR_ONCE(tile_sort_mid_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TSORT_MID * 4);
                tile_sort_mid_kernel<true><<<g, 1024, TSORT_MID * 4, s>>>(ranges, pairs, point_list, buckets, tiles, hdr);
                GSB_LAUNCH_CHECK();
            }
        }
#endif
        if (!radix) {
            tile_sort_small_kernel<false><<<tiles, TSORT_THREADS, 0, s>>>(ranges, pairs, point_list, buckets, hdr);
            GSB_LAUNCH_CHECK();
            if (grid_instances > TSORT_SMALL) {   // a longer tile list is only possible then
                GSB_SET_ATTR_ONCE(tile_sort_mid_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TSORT_MID * 4);
                tile_sort_mid_kernel<false><<<g, 1024, TSORT_MID * 4, s>>>(ranges, pairs, point_list, buckets, tiles, hdr);
                GSB_LAUNCH_CHECK();
            }
        }
    }
    return GSB_OK;
}

}  // namespace gsb
