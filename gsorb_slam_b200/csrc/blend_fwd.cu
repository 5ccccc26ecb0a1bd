// blend_fwd.cu -- per-tile front-to-back alpha blending (K6) for sm_100a.
//
// Replaces FORWARD::render / renderCUDA<3> (forward.cu:261-401, launch :490-502).
// Semantics kept bit-for-decision: integer pixel centres, power > 0 skip, alpha = min(0.99,
// o * exp(power)), alpha < 1/255 skip, stop BEFORE blending once T(1-alpha) < 1e-4, median
// depth = depth of the last splat blended while T > 0.5, n_contrib = list position of the
// last blended splat, colour = C + T * bg, planar CHW output.
//
// Work decomposition (round 2; the round-1 kernel walked the tile list per 4x2 pixel block, four blocks in lock step
// per warp, and issued 9.6 warp-instructions per blended pair -- 27 % of its lane slots did useful work):
//   * one CTA per 16x16 tile, a warp owns an 8x4 pixel region, a LANE owns ONE pixel and pops ITS OWN candidates;
//   * the cull is split: ONCE per CTA and batch, thread = entry turns the entry's conservative alpha >= 1/255 footprint box
//     (preprocess.cu) into a 16-bit column mask and a 16-bit row mask over the tile (four float->int conversions per entry instead
//     of per entry and warp); per window of 32 staged splats a warp then cuts its 8 columns / 4 rows out of the box words with
//     lane = splat, multiplies them out into the 32 x 32 bit matrix "splat j's box covers pixel p" and TRANSPOSES it across the
//     warp with a five-stage shuffle butterfly (two byte permutes, three rotate + bit-select stages), so that lane = pixel holds
//     the bit mask of its own candidates in a register (bit-reversed, so that the next entry in list order is one FLO away);
//   * every lane then pops its candidates of the window in list order; the warp iterates max-over-lanes of the
//     per-PIXEL candidate counts (159 iterations per warp at the headline workload instead of 195 with per-block
//     lists; tests/decomposition_model.py).  The exact tests of forward.cu:346-362 are applied to every candidate, so
//     results are unchanged; the record's power threshold is exact (preprocess.cu), so the exponential is only evaluated
//     for pairs that are blended (or that saturate the pixel);
//   * the bits a lane actually blended are its hit word of the window: one word per (window, pixel), written
//     coalesced (BinningLayout::hits / ImageLayout::hits_tail); the backward pass replays exactly those pairs and needs
//     no alpha test;
//   * splat records are one 48-byte gather straight into a ring of shared-memory buffers (stage.cuh: 3 x 16-byte
//     cp.async per record; one CTA barrier per batch, ids prefetched one batch ahead of the copies);
//   * CH = 5 blends the depth / silhouette pass of the same iteration in the same walk.
#include <cstdlib>
#include <cstddef>
#include "common.cuh"
#include "stage.cuh"

namespace gsb {

// 32 x 32 bit-matrix transpose across a warp: in: lane j holds row j (bit p = column p); out: lane p holds column p.
// Five butterfly stages; stage s exchanges the off-diagonal s x s blocks of every 2s x 2s block.  The two byte-granular
// stages are ONE byte permute each (selector chosen per lane), the three bit-granular ones a rotate of the partner's
// word (left by s for the lane without bit s, right by s for the other: the bits that wrap land under the mask) and
// one bit-select: 13 instructions instead of the 35 of the select-by-lane formulation.
struct TransposeConsts {
    uint32_t sel16, sel8;      // PRMT selectors
    uint32_t m4, m2, m1;       // bits this lane KEEPS of its own word
    uint32_t r4, r2, r1;       // left-rotation of the partner's word
    __device__ __forceinline__ TransposeConsts(uint32_t lane)
    {
        sel16 = (lane & 16) ? 0x3276u : 0x5410u;   // hi: (x & 0xffff0000) | (y >> 16);  lo: (x & 0xffff) | (y << 16)
        sel8 = (lane & 8) ? 0x3715u : 0x6240u;     // hi: (x & 0xff00ff00) | ((y >> 8) & 0x00ff00ff);  lo: (x & 0x00ff00ff) | ((y << 8) & 0xff00ff00)
        m4 = (lane & 4) ? 0xf0f0f0f0u : 0x0f0f0f0fu; r4 = (lane & 4) ? 28u : 4u;
        m2 = (lane & 2) ? 0xccccccccu : 0x33333333u; r2 = (lane & 2) ? 30u : 2u;
        m1 = (lane & 1) ? 0xaaaaaaaau : 0x55555555u; r1 = (lane & 1) ? 31u : 1u;
        // opaque to the optimiser: the eight words stay in registers instead of being re-derived from the lane id in every window
        asm volatile("" : "+r"(sel16), "+r"(sel8), "+r"(m4), "+r"(m2), "+r"(m1), "+r"(r4), "+r"(r2), "+r"(r1));
    }
    __device__ __forceinline__ uint32_t operator()(uint32_t x) const
    {
        x = __byte_perm(x, __shfl_xor_sync(0xffffffffu, x, 16), sel16);
        x = __byte_perm(x, __shfl_xor_sync(0xffffffffu, x, 8), sel8);
        uint32_t y;
        y = __shfl_xor_sync(0xffffffffu, x, 4); x = (x & m4) | (__funnelshift_l(y, y, r4) & ~m4);
        y = __shfl_xor_sync(0xffffffffu, x, 2); x = (x & m2) | (__funnelshift_l(y, y, r2) & ~m2);
        y = __shfl_xor_sync(0xffffffffu, x, 1); x = (x & m1) | (__funnelshift_l(y, y, r1) & ~m1);
        return x;
    }
};

// ((1 << width) - 1) << pos with pos and width clamped to [0, 32]
__device__ __forceinline__ uint32_t bit_mask(int pos, int width)
{
    uint32_t d;
    asm("bmsk.clamp.b32 %0, %1, %2;" : "=r"(d) : "r"(pos), "r"(width));
    return d;
}

// The CTA's shared memory, in ONE struct so that the order is fixed: a lane with nothing to pop forms the address of "entry 32" of
// its window (FLO of an empty mask is -1), i.e. the first record of the next window or, past the last buffer, the start of the
// next array of the struct -- mapped memory either way (the value is never used).
template <int NS>
struct FwdSmem {
    StageRing<NS, BLEND_THREADS, false> ring;
    uint32_t box[NS][BLEND_THREADS];   // per staged entry: 16-bit column mask | 16-bit row mask << 16 of its footprint box inside the tile
    uint32_t max_contrib;
};

__device__ __forceinline__ float4 lds128(uint32_t addr)   // 32-bit shared-window address: register + immediate, no generic-address arithmetic
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// CH = 3: the reference pass.  CH = 5: the RGB pass and the depth / silhouette pass of one mapping iteration
// (src/Render.cc:445-448: colours [r, g, b] and [z_cam, 1, 0] over the SAME geometry) blended together; channel 3
// accumulates depth * alpha * T, channel 4 alpha * T, with the operation order each has in its own reference pass.
template <int MINB, int NS, int CH>
__global__ void __launch_bounds__(BLEND_THREADS, MINB)
blend_forward_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                     const SplatRec* __restrict__ rec, int W, int H, const float* __restrict__ bg,
                     float* __restrict__ out_color, float* __restrict__ out_depth, float* __restrict__ out_depth_sil,
                     float* __restrict__ final_T,
                     uint32_t* __restrict__ n_contrib, uint32_t* __restrict__ tile_max_contrib,
                     uint32_t* __restrict__ hits_full, uint32_t* __restrict__ hits_tail,
                     GeomHeader* __restrict__ hdr, uint32_t layout_capacity, uint32_t band_y0)
{
    constexpr int BATCH = BLEND_THREADS, WINS = BATCH / 32;
    __shared__ FwdSmem<NS> SM;
    StageRing<NS, BATCH, false>& S = SM.ring;
    const uint32_t tile_y = band_y0 + blockIdx.y;   // the grid covers the band's tile rows
    const uint32_t tile = tile_y * gridDim.x + blockIdx.x;
    const uint2 range = ranges[tile];
    const int n = (int)(range.y - range.x);
    const int batches = (n + BATCH - 1) / BATCH;
    // warp -> 8x4 pixel region of the tile; lane -> pixel (column lane & 7, row lane >> 3) = bit index of the coverage words
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int bx0 = blockIdx.x * TILE_X + (warp & 1) * 8, by0 = tile_y * TILE_Y + (warp >> 1) * 4;
    const int px = bx0 + (int)(lane & 7), py = by0 + (int)(lane >> 3);
    const bool inside = px < W && py < H;
    float pxf = (float)px, pyf = (float)py;
    asm volatile("" : "+f"(pxf), "+f"(pyf));   // stay in registers (not re-derived from the thread id per window)
    if (tid == 0) {
        SM.max_contrib = 0;
        if (blockIdx.x == 0 && blockIdx.y == 0) hdr->layout_capacity = layout_capacity;  // the backward pass locates the hit words with it
    }

    uint32_t done = inside ? 0u : 1u;
    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, C3 = 0.f, C4 = 0.f, D = 0.f;
    uint32_t last = 0;

    const uint32_t* ids = point_list + range.x;
    auto load_id = [&](int b) -> uint32_t {
        const int e = b * BATCH + (int)tid;
        return e < n ? __ldg(ids + e) : 0xffffffffu;
    };
    // prologue: batches 0 .. NS-2 in flight, ids of batch NS-1 in a register
#pragma unroll
    for (int i = 0; i < NS - 1; i++) {
        if (i < batches) stage_issue(S, i, rec, load_id(i), 0);
        cp_async_commit();
    }
    uint32_t id_next = NS - 1 < batches ? load_id(NS - 1) : 0xffffffffu;
    bool warp_done = __all_sync(0xffffffffu, done != 0);
    const TransposeConsts transpose(lane);
    // 32-bit shared-window addresses of the ring (a[0][0]) and of the box words; a stage is BATCH + RING_PAD records, b[] and c[] sit
    // NS stages apart
    const uint32_t sa0 = (uint32_t)__cvta_generic_to_shared(&S.a[0][0]), sbox0 = (uint32_t)__cvta_generic_to_shared(&SM.box[0][0]);
    constexpr uint32_t STAGE_BYTES = (BATCH + RING_PAD) * 16, OFF_B = NS * STAGE_BYTES, OFF_C = 2 * NS * STAGE_BYTES;
    static_assert(offsetof(FwdSmem<NS>, box) == sizeof(StageRing<NS, BLEND_THREADS, false>), "box[] must follow the ring (see FwdSmem)");
    const int tile_x0 = blockIdx.x * TILE_X, tile_y0 = tile_y * TILE_Y;
    uint32_t cshift = (warp & 1) * 8, rshift = 16 + (warp >> 1) * 4;   // this warp's 8 columns / 4 rows inside the box word
    asm volatile("" : "+r"(cshift), "+r"(rshift));
    int buf = 0;
    for (int b = 0; b < batches; b++) {
        stage_wait(S, buf, b);  // this thread's copies of batch b have landed
        // ---- cull, part 1 (thread = entry, once per CTA instead of once per warp): the entry's footprint box (preprocess.cu: conservative
        // alpha >= 1/255 extents) against the tile's pixel centres -> columns [c0, c1] and rows [r0, r1] of the tile's 16 x 16 as two bit masks.
        // Pixel column tile_x0 + c is a candidate iff A.x - e.x <= tile_x0 + c <= A.x + e.x (rows alike); an empty range gives a zero mask.
        {
            uint32_t bw = 0;
            if (b * BATCH + (int)tid < n) {   // else: the slot holds a stale record of an earlier batch
                const float4 A = S.a[buf][tid];
                const float2 e = __half22float2(*reinterpret_cast<const __half2*>(&A.z));
                const int c0 = max(__float2int_ru(A.x - e.x) - tile_x0, 0), c1 = min(__float2int_rd(A.x + e.x) - tile_x0, 15);
                const int r0 = max(__float2int_ru(A.y - e.y) - tile_y0, 0), r1 = min(__float2int_rd(A.y + e.y) - tile_y0, 15);
                bw = bit_mask(c0, max(c1 - c0 + 1, 0)) | (bit_mask(r0, max(r1 - r0 + 1, 0)) << 16);
            }
            SM.box[buf][tid] = bw;
        }
        // one barrier per batch: publishes batch b, and everyone is finished with batch b-1 (whose buffer is reused below)
        if (__syncthreads_count(warp_done) == BATCH) break;  // every pixel of the tile is saturated
        {
            const int nbuf = buf == 0 ? NS - 1 : buf - 1;  // (b + NS - 1) % NS
            if (b + NS - 1 < batches) stage_issue(S, nbuf, rec, id_next, 0);
            cp_async_commit();
            if (b + NS < batches) id_next = load_id(b + NS);
        }
        if (!warp_done) {
            const int cnt = min(BATCH, n - b * BATCH);
            const int nwin = (cnt + 31) >> 5;
            // hit words of the batch: rows of 256 words, window after window; the list's last, partial window has its own row
            uint32_t* hp = hits_full + ((size_t)(range.x >> 5) + (size_t)(b * WINS)) * HIT_PIXELS + tid;
            const int wtail = (n >> 5) - b * WINS;   // first window of this batch that is not a full one
            uint32_t ra = sa0 + (uint32_t)buf * STAGE_BYTES;           // the window's records: a at ra, b at ra + OFF_B, c at ra + OFF_C
            uint32_t rb = sbox0 + (uint32_t)buf * (BATCH * 4) + (31 - lane) * 4;   // lane L culls entry 31 - L (see the walk)
            for (int w = 0; w < nwin; w++, hp += HIT_PIXELS, ra += 32 * 16, rb += 32 * 4) {
                // ---- cull, part 2: lane = splat cuts its box word down to the warp's 8 x 4 region, multiplies the 8 column bits and 4 row bits
                // out into the coverage word over the warp's 32 pixels, and the warp transposes the 32 x 32 bit matrix ----
                uint32_t mask;
                {
                    const uint32_t bw = lds32(rb);
                    const uint32_t colm = (bw >> cshift) & 0xffu, rowm = (bw >> rshift) & 0xfu;
                    const uint32_t pm = ((rowm * 0x00204081u) & 0x01010101u) * colm;   // row bits spread to bytes, each byte = the column bits
                    mask = transpose(pm);
                    if (done) mask = 0;
                }
                // ---- walk: every lane pops ITS candidates of the window, in list order ----
                uint32_t hits = 0;
                const uint32_t pos31 = (uint32_t)(b * BATCH + w * 32) + 32u;   // 1-based list position of the window's entry 31
                const uint32_t ra31 = ra + 31 * 16;
                // (the masks are bit-REVERSED: lane L took entry 31 - L of the window in the cull, so entry f is bit 31 - f and the next entry
                // in list order is the HIGHEST set bit: one FLO instead of BREV + FLO)
                while (__any_sync(0xffffffffu, mask != 0)) {
                    uint32_t h;                                    // 31 - f; 0xffffffff for a lane with nothing to pop: reads "entry 32" (the next window's first record or the stage's padding) and discards it
                    asm("bfind.u32 %0, %1;" : "=r"(h) : "r"(mask));
                    uint32_t rbit, bit;                            // PTX shifts clamp the amount: both are 0 when h = 0xffffffff
                    asm("shl.b32 %0, 1, %1;" : "=r"(rbit) : "r"(h));
                    asm("shr.b32 %0, 0x80000000, %1;" : "=r"(bit) : "r"(h));   // 1 << f: the hit words keep list order
                    const bool act = mask != 0;
                    mask ^= rbit;
                    const uint32_t re = ra31 - h * 16u;            // ra + f * 16
                    const float4 A = lds128(re);
                    const float4 B = lds128(re + OFF_B);
                    const float dx = __fsub_rn(A.x, pxf), dy = __fsub_rn(A.y, pyf);
                    const float power = splat_power(dx, dy, B.x, B.y, B.z);
                    if (act && !(power > 0.0f) && !(power < A.w)) {
                        const float alpha = fminf(0.99f, __fmul_rn(B.w, expf(power)));
                        if (!(alpha < 1.0f / 255.0f)) {
                            const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
                            if (test_T < 0.0001f) {
                                asm volatile("" ::: "memory");   // keeps this rare block a branch target (no speculative select chains in the hot path)
                                mask = 0;
                                done = 1u;
                            } else {
                                const float4 Cc = lds128(re + OFF_C);
                                C0 = fmaf(__fmul_rn(Cc.x, alpha), T, C0);
                                C1 = fmaf(__fmul_rn(Cc.y, alpha), T, C1);
                                C2 = fmaf(__fmul_rn(Cc.z, alpha), T, C2);
                                if (CH == 5) {
                                    C3 = fmaf(__fmul_rn(Cc.w, alpha), T, C3);   // colour z_cam = the splat's view-space depth
                                    C4 = fmaf(alpha, T, C4);                    // colour 1
                                }
                                if (T > 0.5f) D = Cc.w;
                                T = test_T;
                                last = pos31 - h;
                                hits |= bit;
                            }
                        }
                    }
                }
                // hit word of (window, pixel): [window][pixel of the tile], coalesced
                if (w == wtail) {   // once per tile
                    asm volatile("" ::: "memory");
                    hp = hits_tail + (size_t)tile * HIT_PIXELS + tid;
                }
                *hp = hits;
                warp_done = __all_sync(0xffffffffu, done != 0);
                if (warp_done) break;
            }
        }
        buf = buf == NS - 1 ? 0 : buf + 1;
    }
    cp_async_wait<0>();
    if (inside) {
        const size_t HW = (size_t)W * H, pix = (size_t)py * W + px;
        final_T[pix] = T;
        n_contrib[pix] = last;
        out_color[pix] = fmaf(T, __ldg(bg), C0);
        out_color[HW + pix] = fmaf(T, __ldg(bg + 1), C1);
        out_color[2 * HW + pix] = fmaf(T, __ldg(bg + 2), C2);
        out_depth[pix] = D;
        if (CH == 5) {   // the depth pass shares the background tensor: its channels 0 / 1 get bg[0] / bg[1]
            out_depth_sil[pix] = fmaf(T, __ldg(bg), C3);
            out_depth_sil[HW + pix] = fmaf(T, __ldg(bg + 1), C4);
        }
    }
    // highest n_contrib of the tile: the backward pass starts there
    const uint32_t m = __reduce_max_sync(0xffffffffu, last);
    __syncthreads();   // max_contrib was zeroed by thread 0 before this barrier
    if (lane == 0 && m) atomicMax(&SM.max_contrib, m);
    __syncthreads();
    if (tid == 0) tile_max_contrib[2 * tile] = tile_max_contrib[2 * tile + 1] = SM.max_contrib;
}

template <int MINB, int CH>
static void launch_fwd(const FwdParams& p, char* geom, const GeomLayout& GL, char* binning, const BinningLayout& BL, char* image,
                       const ImageLayout& IL, float* out_color, float* out_depth, float* out_depth_sil, cudaStream_t s)
{
    constexpr int NS = 2;
    // many resident CTAs x 24 KB: ask for the largest shared-memory carve-out (once per process and kernel)
    GSB_SET_ATTR_ONCE((blend_forward_kernel<MINB, NS, CH>), cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    blend_forward_kernel<MINB, NS, CH><<<dim3(IL.tiles_x, p.band_y1 - p.band_y0), BLEND_THREADS, 0, s>>>(
        reinterpret_cast<const uint2*>(image + IL.ranges), reinterpret_cast<const uint32_t*>(binning + BL.point_list),
        reinterpret_cast<const SplatRec*>(geom + GL.rec), p.W, p.H, p.background, out_color, out_depth, out_depth_sil,
        reinterpret_cast<float*>(image + IL.final_T), reinterpret_cast<uint32_t*>(image + IL.n_contrib),
        reinterpret_cast<uint32_t*>(image + IL.tile_max_contrib), reinterpret_cast<uint32_t*>(binning + BL.hits),
        reinterpret_cast<uint32_t*>(image + IL.hits_tail), reinterpret_cast<GeomHeader*>(geom + GL.header), (uint32_t)BL.capacity,
        (uint32_t)p.band_y0);
}

int launch_blend_forward(const FwdParams& p, char* geom, const GeomLayout& GL, char* binning, const BinningLayout& BL,
                         char* image, const ImageLayout& IL, float* out_color, float* out_depth, float* out_depth_sil,
                         cudaStream_t s)
{
    if (p.W <= 0 || p.H <= 0 || p.band_y1 <= p.band_y0) return GSB_OK;
    StageTimer _t(ST_BLEND_FWD, s);
#define GSB_FWD(MB)                                                                                                   \
    do {                                                                                                              \
        if (out_depth_sil) launch_fwd<MB, 5>(p, geom, GL, binning, BL, image, IL, out_color, out_depth, out_depth_sil, s); \
        else launch_fwd<MB, 3>(p, geom, GL, binning, BL, image, IL, out_color, out_depth, out_depth_sil, s);          \
    } while (0)
#ifdef GSB_TUNING   // developer builds only (make tune): resident CTAs per SM the compiler must allow = the register budget
    static const int minb = [] { const char* e = getenv("GSB_BLEND_FWD_MINB"); return e ? atoi(e) : 0; }();
    if (minb == 5) GSB_FWD(5); else if (minb == 7) GSB_FWD(7); else if (minb == 8) GSB_FWD(8); else
#endif
    GSB_FWD(6);
#undef GSB_FWD
    GSB_LAUNCH_CHECK();
    return GSB_OK;
}

}  // namespace gsb
