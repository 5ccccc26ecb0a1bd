// blend_fwd.cu -- per-tile front-to-back alpha blending (K6) for sm_100a.
//
// Replaces FORWARD::render / renderCUDA<3> (forward.cu:261-401, launch :490-502).
// Semantics kept bit-for-decision: integer pixel centres, power > 0 skip, alpha = min(0.99,
// o * exp(power)), alpha < 1/255 skip, stop BEFORE blending once T(1-alpha) < 1e-4, median
// depth = depth of the last splat blended while T > 0.5, n_contrib = list position of the
// last blended splat, colour = C + T * bg, planar CHW output.
//
// What is different from the reference kernel (design, not results):
//   * one CTA per 16x16 tile as before, but each QUARTER-WARP owns a 4x2 pixel block: per 32
//     staged splats a warp runs one cull pass (lane j tests splat j's conservative footprint
//     box, computed in preprocess, against the warp's four blocks -> four ballots), then every
//     quarter-warp walks only ITS survivors, so a pixel evaluates the few splats that can
//     reach its block instead of every splat binned to the tile;
//   * a quarter / warp retires as soon as its pixels are saturated;
//   * splat records are one 48-byte gather straight into a ring of shared-memory buffers
//     (stage.cuh: 3 x 16-byte cp.async per record, or one 48-byte bulk copy; one CTA barrier per
//     batch, ids prefetched one batch ahead of the copies) instead of four separate arrays plus a
//     per-pair colour read from global memory;
//   * per window of 32 list entries and per 4x2 block the kernel records WHICH entries were
//     blended into at least one pixel of the block (the "hit words"): the backward pass walks
//     exactly those, and starts at the highest n_contrib of each half tile, also recorded here;
//   * CH = 5 blends the depth / silhouette pass of the same iteration in the same walk.
#include <cstdlib>
#include "common.cuh"
#include "stage.cuh"

namespace gsb {

// CH = 3: the reference pass.  CH = 5: the RGB pass and the depth / silhouette pass of one mapping iteration
// (src/Render.cc:445-448: colours [r, g, b] and [z_cam, 1, 0] over the SAME geometry) blended together; channel 3
// accumulates depth * alpha * T, channel 4 alpha * T, with the operation order each has in its own reference pass.
template <int MINB, int NS, int HALVES, int CH, bool BULK>
__global__ void __launch_bounds__(256 / HALVES, MINB)
blend_forward_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                     const SplatRec* __restrict__ rec, int W, int H, const float* __restrict__ bg,
                     float* __restrict__ out_color, float* __restrict__ out_depth, float* __restrict__ out_depth_sil,
                     float* __restrict__ final_T,
                     uint32_t* __restrict__ n_contrib, uint32_t* __restrict__ tile_max_contrib,
                     uint32_t* __restrict__ hits_full, uint32_t* __restrict__ hits_tail,
                     GeomHeader* __restrict__ hdr, uint32_t layout_capacity, uint32_t band_y0)
{
    constexpr int BLEND_THREADS = 256 / HALVES, BLEND_BATCH = BLEND_THREADS;
    __shared__ StageRing<NS, BLEND_BATCH, BULK> S;
    __shared__ uint32_t s_max[2];
    const uint32_t tile_y = band_y0 + blockIdx.y / HALVES, half = blockIdx.y % HALVES;   // the grid covers the band's tile rows
    const uint32_t tile = tile_y * gridDim.x + blockIdx.x;
    const uint2 range = ranges[tile];
    const int n = (int)(range.y - range.x);
    const int batches = (n + BLEND_BATCH - 1) / BLEND_BATCH;
    // warp (numbered 0..7 over the whole tile) -> 8x4 pixel block of the tile; quarter-warp q -> 4x2 sub-block; lane -> pixel
    const uint32_t warp = half * (8 / HALVES) + (threadIdx.x >> 5), lane = lane_id();
    const uint32_t q = lane >> 3, l8 = lane & 7, qshift = q * 8;
    const int bx0 = blockIdx.x * TILE_X + (warp & 1) * 8, by0 = tile_y * TILE_Y + (warp >> 1) * 4;
    const int px = bx0 + (q & 1) * 4 + (l8 & 3), py = by0 + (q >> 1) * 2 + (l8 >> 2);
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    // sub-block extents used by the cull pass (pixel centres): x halves [bx0, bx0+3], [bx0+4, bx0+7];
    // y halves [by0, by0+1], [by0+2, by0+3]
    const float xa0 = (float)bx0, xa1 = (float)(bx0 + 3), xb0 = (float)(bx0 + 4), xb1 = (float)(bx0 + 7);
    const float ya0 = (float)by0, ya1 = (float)(by0 + 1), yb0 = (float)(by0 + 2), yb1 = (float)(by0 + 3);
    if (threadIdx.x == 0) {
        s_max[0] = s_max[1] = 0;
        if (blockIdx.x == 0 && blockIdx.y == 0) hdr->layout_capacity = layout_capacity;  // the backward pass locates the hit words with it
    }

    bool done = !inside;
    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, C3 = 0.f, C4 = 0.f, D = 0.f;
    uint32_t last = 0;

    const uint32_t* ids = point_list + range.x;
    auto load_id = [&](int b) -> uint32_t {
        const int e = b * BLEND_BATCH + (int)threadIdx.x;
        return e < n ? __ldg(ids + e) : 0xffffffffu;
    };
    auto batch_count = [&](int b) { return min(BLEND_BATCH, n - b * BLEND_BATCH); };
    S.init();
    // prologue: batches 0 .. NS-2 in flight, ids of batch NS-1 in a register
    int issued = -1, waited = -1;   // highest batch issued / waited for (the bulk engine must drain before the CTA exits)
#pragma unroll
    for (int i = 0; i < NS - 1; i++) {
        if (i < batches) {
            stage_issue(S, i, rec, load_id(i), batch_count(i));
            issued = i;
        }
        cp_async_commit();
    }
    uint32_t id_next = NS - 1 < batches ? load_id(NS - 1) : 0xffffffffu;
    uint32_t done_bits = __ballot_sync(0xffffffffu, done);
    bool warp_done = done_bits == 0xffffffffu;
    int buf = 0;
    for (int b = 0; b < batches; b++) {
        stage_wait(S, buf, b);  // batch b has landed (LDGSTS: this thread's copies; bulk: the stage's mbarrier phase)
        waited = b;
        // one barrier per batch: publishes batch b, and everyone is finished with batch b-1 (whose buffer is reused below)
        if (__syncthreads_count(warp_done) == BLEND_THREADS) break;  // every pixel of the tile is saturated
        {
            const int nbuf = buf == 0 ? NS - 1 : buf - 1;  // (b + NS - 1) % NS
            if (b + NS - 1 < batches) {
                stage_issue(S, nbuf, rec, id_next, batch_count(b + NS - 1));
                issued = b + NS - 1;
            }
            cp_async_commit();
            if (b + NS < batches) id_next = load_id(b + NS);
        }
        if (!warp_done) {
            const int cnt = min(BLEND_BATCH, n - b * BLEND_BATCH);
            for (int c0 = 0; c0 < cnt; c0 += 32) {
                // ---- cull pass: lane j tests staged splat c0+j against the four 4x2 sub-blocks ----
                const int j = c0 + (int)lane;
                bool hxa = false, hxb = false, hya = false, hyb = false;
                if (j < cnt) {
                    const float4 A = S.A(buf, j);
                    const float2 e = __half22float2(*reinterpret_cast<const __half2*>(&A.z));
                    const float lox = A.x - e.x, hix = A.x + e.x, loy = A.y - e.y, hiy = A.y + e.y;
                    hxa = !(hix < xa0 || lox > xa1);
                    hxb = !(hix < xb0 || lox > xb1);
                    hya = !(hiy < ya0 || loy > ya1);
                    hyb = !(hiy < yb0 || loy > yb1);
                }
                const uint32_t m0 = __ballot_sync(0xffffffffu, hxa && hya), m1 = __ballot_sync(0xffffffffu, hxb && hya);
                const uint32_t m2 = __ballot_sync(0xffffffffu, hxa && hyb), m3 = __ballot_sync(0xffffffffu, hxb && hyb);
                uint32_t mask = q == 0 ? m0 : q == 1 ? m1 : q == 2 ? m2 : m3;
                if (((done_bits >> qshift) & 0xffu) == 0xffu) mask = 0;  // this quarter is saturated
                uint32_t lane_hits = 0;  // entries of this window blended into THIS pixel
                // ---- blend pass: every quarter-warp walks ITS survivors, in list order ----
                while (__any_sync(0xffffffffu, mask != 0)) {
                    const bool act = mask != 0;
                    const int e = c0 + (act ? __ffs(mask) - 1 : 0);
                    const uint32_t bit = mask & (0u - mask);
                    mask ^= bit;
                    const float4 A = S.A(buf, e);
                    const float4 B = S.B(buf, e);
                    const float dx = __fsub_rn(A.x, pxf), dy = __fsub_rn(A.y, pyf);
                    const float power = splat_power(dx, dy, B.x, B.y, B.z);
                    if (!act || done || power > 0.0f || power < A.w) continue;
                    const float alpha = fminf(0.99f, __fmul_rn(B.w, expf(power)));
                    if (alpha < 1.0f / 255.0f) continue;
                    const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
                    if (test_T < 0.0001f) {
                        done = true;
                        continue;
                    }
                    const float4 Cc = S.C(buf, e);
                    C0 = fmaf(__fmul_rn(Cc.x, alpha), T, C0);
                    C1 = fmaf(__fmul_rn(Cc.y, alpha), T, C1);
                    C2 = fmaf(__fmul_rn(Cc.z, alpha), T, C2);
                    if (CH == 5) {
                        C3 = fmaf(__fmul_rn(Cc.w, alpha), T, C3);   // colour z_cam = the splat's view-space depth
                        C4 = fmaf(alpha, T, C4);                    // colour 1
                    }
                    if (T > 0.5f) D = Cc.w;
                    T = test_T;
                    last = (uint32_t)(b * BLEND_BATCH + e + 1);
                    lane_hits |= bit;
                }
                // hit word of (window, 4x2 block): OR over the quarter's 8 lanes
                lane_hits |= __shfl_xor_sync(0xffffffffu, lane_hits, 4);
                lane_hits |= __shfl_xor_sync(0xffffffffu, lane_hits, 2);
                lane_hits |= __shfl_xor_sync(0xffffffffu, lane_hits, 1);
                if (l8 == 0)
                    *hit_word(hits_full, hits_tail, tile, range.x, (uint32_t)n, (uint32_t)(b * (BLEND_BATCH / 32) + (c0 >> 5)),
                              warp * 4 + q) = lane_hits;
                done_bits = __ballot_sync(0xffffffffu, done);
                if (done_bits == 0xffffffffu) {
                    warp_done = true;
                    break;
                }
            }
        }
        buf = buf == NS - 1 ? 0 : buf + 1;
    }
    cp_async_wait<0>();
    if (BULK)   // copies still in flight target this CTA's shared memory: wait for them before it is released
        for (int k = waited + 1; k <= issued; k++) stage_wait(S, k % NS, k);
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
    uint32_t m = last;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0 && m) atomicMax(&s_max[warp >> 2], m);
    __syncthreads();
    // highest n_contrib of the upper and of the lower half of the tile (the backward pass starts there)
    if (HALVES == 2) {
        if (threadIdx.x == 0) tile_max_contrib[2 * tile + half] = s_max[half];
    } else {
        if (threadIdx.x < 2) tile_max_contrib[2 * tile + threadIdx.x] = s_max[threadIdx.x];
    }
}

int launch_blend_forward(const FwdParams& p, char* geom, const GeomLayout& GL, char* binning, const BinningLayout& BL,
                         char* image, const ImageLayout& IL, float* out_color, float* out_depth, float* out_depth_sil,
                         cudaStream_t s)
{
    const uint32_t* point_list = reinterpret_cast<const uint32_t*>(binning + BL.point_list);
    if (p.W <= 0 || p.H <= 0 || p.band_y1 <= p.band_y0) return GSB_OK;
    // tuning knobs: CTA shape (whole tile / half tile), resident CTAs per SM the compiler must allow (the register
    // budget), depth of the staging ring
    static const int halves = [] { const char* e = getenv("GSB_BLEND_FWD_HALVES"); return e ? atoi(e) : 1; }();
    static const int minb = [] { const char* e = getenv("GSB_BLEND_FWD_MINB"); return e ? atoi(e) : 0; }();
    static const int stages = [] { const char* e = getenv("GSB_BLEND_FWD_STAGES"); return e ? atoi(e) : 2; }();
    {
        StageTimer _t(ST_BLEND_FWD, s);
#define GSB_FWD_LAUNCH(MB, NS, HV) GSB_FWD_LAUNCH_CH(MB, NS, HV, 3, false)
#define GSB_FWD_LAUNCH_CH(MB, NS, HV, CH, BK)                                                                                      \
    do {                                                                                                                \
        static const bool attr_set = [] {  /* many resident CTAs x 12-37 KB: ask for the largest carve-out */          \
            cudaFuncSetAttribute(blend_forward_kernel<MB, NS, HV, CH, BK>, cudaFuncAttributePreferredSharedMemoryCarveout, 100); \
            return true;                                                                                                \
        }();                                                                                                            \
        (void)attr_set;                                                                                                 \
        blend_forward_kernel<MB, NS, HV, CH, BK><<<dim3(IL.tiles_x, (p.band_y1 - p.band_y0) * HV), 256 / HV, 0, s>>>(                    \
            reinterpret_cast<const uint2*>(image + IL.ranges), point_list, reinterpret_cast<const SplatRec*>(geom + GL.rec), \
            p.W, p.H, p.background, out_color, out_depth, out_depth_sil, reinterpret_cast<float*>(image + IL.final_T),  \
            reinterpret_cast<uint32_t*>(image + IL.n_contrib), reinterpret_cast<uint32_t*>(image + IL.tile_max_contrib), \
            reinterpret_cast<uint32_t*>(binning + BL.hits), reinterpret_cast<uint32_t*>(image + IL.hits_tail),          \
            reinterpret_cast<GeomHeader*>(geom + GL.header), (uint32_t)BL.capacity, (uint32_t)p.band_y0);                                    \
    } while (0)
        static const bool bulk = [] { const char* e = getenv("GSB_BLEND_STAGE"); return e ? e[0] == 'b' : GSB_DEFAULT_BULK; }();
        if (out_depth_sil) {
            if (bulk) GSB_FWD_LAUNCH_CH(6, 2, 1, 5, true); else GSB_FWD_LAUNCH_CH(6, 2, 1, 5, false);
        } else if (bulk && halves == 1 && stages != 3 && minb != 8) {
            GSB_FWD_LAUNCH_CH(6, 2, 1, 3, true);
        } else if (halves == 2) {
            if (stages == 3) { if (minb == 12) GSB_FWD_LAUNCH(12, 3, 2); else GSB_FWD_LAUNCH(16, 3, 2); }
            else { if (minb == 12) GSB_FWD_LAUNCH(12, 2, 2); else GSB_FWD_LAUNCH(16, 2, 2); }
        } else {
            if (stages == 3) GSB_FWD_LAUNCH(6, 3, 1);
            else { if (minb == 8) GSB_FWD_LAUNCH(8, 2, 1); else GSB_FWD_LAUNCH(6, 2, 1); }
        }
#undef GSB_FWD_LAUNCH
#undef GSB_FWD_LAUNCH_CH
        GSB_LAUNCH_CHECK();
    }
    return GSB_OK;
}

}  // namespace gsb
