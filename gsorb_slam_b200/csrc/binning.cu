// binning.cu -- tile-instance emission (K3), cub-free onesweep radix sort (K4) and tile
// ranges (K5) for sm_100a.
//
// Replaces (reference file:line, behaviour only):
//   duplicateWithKeys                       rasterizer_impl.cu:71-112
//   cub::DeviceRadixSort::SortPairs         rasterizer_impl.cu:307-315   (stable LSD sort of
//                                           (tile << 32 | depth bits, gaussian id) pairs)
//   identifyTileRanges (+ memset)           rasterizer_impl.cu:117-139, :317
//
// The sort is a hand-written "onesweep" least-significant-digit radix sort: one histogram
// kernel for all digit positions, then one kernel per 8-bit digit that ranks a 4096-key tile
// in shared memory (warp match_any ranking, stable), resolves its global offsets with a
// decoupled look-back over the preceding tiles and scatters keys + values -- each pass reads
// and writes every pair exactly once.  Stability is part of the contract: equal keys keep
// ascending Gaussian order, which is how the reference resolves depth ties.
#include "common.cuh"

namespace gsb {

constexpr int DUP_THREADS = 256;

// ---- K3: one thread per Gaussian writes its (key, id) run -----------------------------------
__global__ void __launch_bounds__(DUP_THREADS)
duplicate_kernel(int P, const SplatRec* __restrict__ rec, const int* __restrict__ radii,
                 const uint32_t* __restrict__ tiles_touched, const uint32_t* __restrict__ block_offsets,
                 const GeomHeader* __restrict__ hdr, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals,
                 int tiles_x, int tiles_y)
{
    __shared__ uint32_t s_warp[DUP_THREADS / 32];
    const int idx = blockIdx.x * DUP_THREADS + threadIdx.x;
    const uint32_t touched = idx < P ? tiles_touched[idx] : 0u;
    uint32_t v = touched;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane_id() >= (uint32_t)o) v += t;
    }
    if (lane_id() == 31) s_warp[threadIdx.x >> 5] = v;
    __syncthreads();
    uint32_t warp_excl = 0;
#pragma unroll
    for (int w = 0; w < DUP_THREADS / 32; w++)
        if (w < (int)(threadIdx.x >> 5)) warp_excl += s_warp[w];
    if (touched == 0) return;
    uint32_t off = block_offsets[blockIdx.x] + warp_excl + v - touched;
    const uint32_t cap = hdr->num_rendered_clamped;
    const float4 a = rec[idx].a;
    const float depth = rec[idx].b.w;
    uint32_t minx, miny, maxx, maxy;
    get_rect(a.x, a.y, radii[idx], tiles_x, tiles_y, minx, miny, maxx, maxy);
    const uint64_t dbits = (uint64_t)__float_as_uint(depth);
    for (uint32_t y = miny; y < maxy; y++)
        for (uint32_t x = minx; x < maxx; x++) {
            if (off < cap) {
                keys[off] = ((uint64_t)(y * (uint32_t)tiles_x + x) << 32) | dbits;
                vals[off] = (uint32_t)idx;
            }
            off++;
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

// ---- K5: tile ranges ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tile_ranges_kernel(const uint64_t* __restrict__ keys, const GeomHeader* __restrict__ hdr, uint2* __restrict__ ranges)
{
    const uint32_t R = hdr->num_rendered_clamped;
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < R; i += gridDim.x * 256) {
        const uint32_t cur = (uint32_t)(keys[i] >> 32);
        if (i == 0) {
            ranges[cur].x = 0;
        } else {
            const uint32_t prev = (uint32_t)(keys[i - 1] >> 32);
            if (cur != prev) {
                ranges[prev].y = i;
                ranges[cur].x = i;
            }
        }
        if (i == R - 1) ranges[cur].y = R;
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
    GSB_CUDA_CHECK(cudaFuncSetAttribute(onesweep_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)sizeof(SortSmem)));
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

int sort_passes_for(int tiles)
{
    int bits = 0;
    while ((1 << bits) < tiles) bits++;
    const int key_bits = 32 + (bits > 0 ? bits : 1);
    return (key_bits + SORT_RADIX_BITS - 1) / SORT_RADIX_BITS;
}

int launch_binning(const FwdParams& p, char* geom, const GeomLayout& GL, char* binning, const BinningLayout& BL,
                   char* image, const ImageLayout& IL, long long grid_instances, cudaStream_t s)
{
    GeomHeader* hdr = reinterpret_cast<GeomHeader*>(geom + GL.header);
    uint64_t* kbuf[2] = {reinterpret_cast<uint64_t*>(binning + BL.keys0), reinterpret_cast<uint64_t*>(binning + BL.keys1)};
    uint32_t* vbuf[2] = {reinterpret_cast<uint32_t*>(binning + BL.vals0), reinterpret_cast<uint32_t*>(binning + BL.vals1)};
    uint32_t* hist = reinterpret_cast<uint32_t*>(binning + BL.hist);
    uint32_t* lookback = reinterpret_cast<uint32_t*>(binning + BL.lookback);
    uint2* ranges = reinterpret_cast<uint2*>(image + IL.ranges);
    const int tiles = p.tiles_x * p.tiles_y;
    const int passes = sort_passes_for(tiles);
    int cur = passes & 1;  // an even number of ping-pong passes ends in buffer 0
    GSB_CUDA_CHECK(cudaMemsetAsync(ranges, 0, (size_t)tiles * sizeof(uint2), s));
    if (grid_instances <= 0 || p.P <= 0) return GSB_OK;
    const int sort_tiles = (int)((grid_instances + SORT_TILE - 1) / SORT_TILE);
    // hist and lookback are adjacent in the blob: one memset
    GSB_CUDA_CHECK(cudaMemsetAsync(hist, 0, BL.lookback - BL.hist + (size_t)passes * sort_tiles * SORT_RADIX * 4, s));

    {

        StageTimer _t(ST_DUPLICATE, s);

        duplicate_kernel<<<GL.num_blocks, DUP_THREADS, 0, s>>>(
            p.P, reinterpret_cast<const SplatRec*>(geom + GL.rec), reinterpret_cast<const int*>(geom + GL.radii),
            reinterpret_cast<const uint32_t*>(geom + GL.tiles_touched), reinterpret_cast<const uint32_t*>(geom + GL.block_offsets),
            hdr, kbuf[cur], vbuf[cur], p.tiles_x, p.tiles_y);
        GSB_LAUNCH_CHECK();

    }
    if (int rc = launch_sort_pairs(hdr, kbuf, vbuf, cur, passes, hist, lookback, sort_tiles, s)) return rc;
    // cur == 0 here
    const int rg = sort_tiles * (SORT_TILE / 256) < NUM_SMS * 8 ? sort_tiles * (SORT_TILE / 256) : NUM_SMS * 8;
    {
        StageTimer _t(ST_RANGES, s);
        tile_ranges_kernel<<<rg, 256, 0, s>>>(kbuf[0], hdr, ranges);
        GSB_LAUNCH_CHECK();
    }
    return GSB_OK;
}

}  // namespace gsb
