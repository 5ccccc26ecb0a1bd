// blend_bwd.cu -- per-tile back-to-front gradient of the alpha blend (K7) for sm_100a.
//
// Replaces BACKWARD::render / renderCUDA<3> (backward.cu:399-557, launch :641-656).
// Same recurrences as the reference: start from the stored final transmittance and the
// position of the last blended splat, walk the tile's list backwards, T <- T / (1 - alpha),
// running "colour behind" accumulator, background term, gradients w.r.t. colour, 2D mean
// (in NDC units: x 0.5 W, x 0.5 H), conic (slots x, y, w) and opacity.  No depth gradient.
// Per pair only u = G dL/dalpha and its first / second moments in (dx, dy) are formed; the
// linear maps from those 6 sums to d(mean2D), d(conic), d(opacity) use per-Gaussian constants
// and are applied once per Gaussian in gauss_bwd.cu (packed accumulator layout:
// {S u dx, S u dy, S u dx^2, S u dx dy, S u dy^2, S u, S wc d_r, S wc d_g, S wc d_b}).
//
// What is different (design, not results):
//   * the reference issues 9 global atomicAdds per contributing (pixel, splat) pair; here a
//     QUARTER-WARP owns a 4x2 pixel block, reduces the 9 partial sums of a splat across its 8
//     lanes with a shuffle reduce-scatter (10 SHFL) and issues two RED.ADD.F32 instructions
//     into one 48-byte packed accumulator per Gaussian, and only for splats that touched at
//     least one pixel of the block; the four quarters of a warp work on four different splats
//     at once;
//   * no cull pass at all: the forward pass records, per window of 32 list entries and per PIXEL, the bit mask of the
//     entries that were blended into the pixel ("hit words", BinningLayout::hits).  The backward ORs the eight words of a
//     4x2 block into the block's word while it stages a batch (two LDG.128 per thread) and every quarter-warp walks
//     exactly the entries that hit ITS block, back to front -- 25 % fewer (block, splat) visits than the conservative
//     footprint test and no per-batch ballots;
//   * the walk starts at the tile's highest n_contrib (recorded by the forward pass), not at
//     the end of the tile's list;
//   * records are gathered with cp.async into a ring of shared-memory buffers with one CTA
//     barrier per batch (stage.cuh).
#include <cstdlib>
#include "common.cuh"
#include "stage.cuh"

namespace gsb {

// CH = 5: backward of the fused RGB + depth / silhouette pass (see blend_fwd.cu): dL/dalpha sums over five channels,
// and the gradient of the z_cam colour (sum of w * dL/dpix[3]) lands in accumulator slot 9.
template <int MINB, int NS, int HALVES, int CH, bool BULK>
__global__ void __launch_bounds__(256 / HALVES, MINB)
blend_backward_kernel(const uint2* __restrict__ ranges, const char* __restrict__ binning,
                      const SplatRec* __restrict__ rec, int W, int H, const float* __restrict__ bg,
                      const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
                      const uint32_t* __restrict__ tile_max_contrib, const float* __restrict__ dL_dpix,
                      const float* __restrict__ dL_ddepth_sil, float* __restrict__ acc /* [P][12] */, const uint32_t* __restrict__ hits_tail,
                      const GeomHeader* __restrict__ hdr, uint32_t band_y0)
{
    constexpr int HIT_BLOCKS = 32; constexpr int BLEND_THREADS = 256 / HALVES, BLEND_BATCH = BLEND_THREADS, BLOCKS = HIT_BLOCKS / HALVES;
    __shared__ StageRing<NS, BLEND_BATCH, BULK> S;
    __shared__ uint32_t s_ids[NS][BLEND_BATCH];
    __shared__ uint32_t s_hits[NS][(BLEND_BATCH / 32) * BLOCKS];  // [window of the batch][4x2 block of this CTA]
    const uint32_t tile_y = band_y0 + blockIdx.y / HALVES, half = blockIdx.y % HALVES;   // the grid covers the band's tile rows
    const uint32_t tile = tile_y * gridDim.x + blockIdx.x;
    const uint2 range = ranges[tile];
    const uint32_t len = range.y - range.x;
    // entries [0, n) can matter: n = highest n_contrib over this CTA's pixels (recorded per half tile by the forward)
    const uint32_t tmax = HALVES == 2 ? tile_max_contrib[2 * tile + half] : max(tile_max_contrib[2 * tile], tile_max_contrib[2 * tile + 1]);
    const int n = min((int)len, (int)tmax);
    const int batches = (n + BLEND_BATCH - 1) / BLEND_BATCH;
    if (batches == 0) return;
    const BinningLayout BL = BinningLayout::make((long long)hdr->layout_capacity);
    const uint32_t* point_list = reinterpret_cast<const uint32_t*>(binning + BL.point_list);
    const uint32_t* hits_full = reinterpret_cast<const uint32_t*>(binning + BL.hits);
    const uint32_t lwarp = threadIdx.x >> 5, warp = half * (8 / HALVES) + lwarp, lane = lane_id();   // warp: 0..7 over the tile
    const uint32_t q = lane >> 3, l8 = lane & 7, qshift = q * 8;
    const int bx0 = blockIdx.x * TILE_X + (warp & 1) * 8, by0 = tile_y * TILE_Y + (warp >> 1) * 4;
    const int px = bx0 + (q & 1) * 4 + (l8 & 3), py = by0 + (q >> 1) * 2 + (l8 >> 2);
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    const size_t HW = (size_t)W * H, pix = (size_t)py * W + px;

    const float T_final = inside ? final_T[pix] : 0.f;
    float T = T_final;
    const int last_contributor = inside ? (int)n_contrib[pix] : 0;
    float d0 = 0.f, d1 = 0.f, d2 = 0.f;
    if (inside) {
        d0 = dL_dpix[pix];
        d1 = dL_dpix[HW + pix];
        d2 = dL_dpix[2 * HW + pix];
    }
    float d3 = 0.f, d4 = 0.f;
    if (CH == 5 && inside) {
        d3 = dL_ddepth_sil[pix];
        d4 = dL_ddepth_sil[HW + pix];
    }
    float bg_dot_dpixel = __ldg(bg) * d0 + __ldg(bg + 1) * d1 + __ldg(bg + 2) * d2;
    if (CH == 5) bg_dot_dpixel += __ldg(bg) * d3 + __ldg(bg + 1) * d4;   // the depth pass blends over the same background tensor
    const float d8k = (CH == 5 && (l8 & 4)) ? d3 : d2, d8s = (l8 & 4) ? d2 : d3;   // keep / send factors of sums 8 and 9
    const float dk = (l8 & 4) ? d1 : d0, ds = (l8 & 4) ? d0 : d1;   // stage-1 keep / send factors of the colour sums
    // accumulator slot of the sum lane l8 ends up with (GradAcc layout: {Su dx, Su dy, Su dx^2, Su dxdy, Su dy^2, Su, Sw dr, Sw dg, Sw db})
    const int acc_slot = l8 == 0 ? 0 : l8 == 1 ? 2 : l8 == 2 ? 3 : l8 == 3 ? 6 : l8 == 4 ? 1 : l8 == 5 ? 4 : l8 == 6 ? 5 : 7;
    // Colour seen behind the current splat, as ONE scalar: s = dL/dpix . (sum of c_j alpha_j T_j over the splats j already walked)
    // + T_final bg . dL/dpix.  With T_k(1 - alpha_k) = T_after the reference's T_k (c_k - accum_rec) . dL/dpix - T_final/(1 - alpha_k) bg . dL/dpix
    // (backward.cu:505-535: per-channel accum_rec / last_color recurrences) equals (T_k c_k . dL/dpix - s) / (1 - alpha_k).
    float s_behind = T_final * bg_dot_dpixel;
    // last window (of 32 list entries) in which this quarter-warp's 4x2 block blended anything: the forward
    // pass wrote hit words for every window up to it
    int qmax = last_contributor;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) qmax = max(qmax, __shfl_xor_sync(0xffffffffu, qmax, o));
    const int wq_last = (qmax - 1) >> 5;  // -1 when the block never blended

    // batches are walked back to front; inside a batch slot t holds list entry kb*256 + t
    const uint32_t* ids = point_list + range.x;
    auto load_id = [&](int k) -> uint32_t {  // k-th batch in walking order
        const int i = (batches - 1 - k) * BLEND_BATCH + (int)threadIdx.x;
        return (k < batches && i < n) ? __ldg(ids + i) : 0xffffffffu;
    };
    auto stage = [&](int k, int buf, uint32_t id) {
        s_ids[buf][threadIdx.x] = id;
        stage_issue(S, buf, rec, id, min(BLEND_BATCH, n - (batches - 1 - k) * BLEND_BATCH));
        // hit words of the batch: (BLEND_BATCH / 32) windows x BLOCKS blocks, one 4-byte copy per thread
        const uint32_t w = (uint32_t)(batches - 1 - k) * (BLEND_BATCH / 32) + threadIdx.x / BLOCKS;
        if (threadIdx.x < (BLEND_BATCH / 32) * BLOCKS && (int)(w * 32) < n) {
            const uint32_t blk = half * BLOCKS + threadIdx.x % BLOCKS, wp = blk >> 2, qq = blk & 3;
            const uint4* row = reinterpret_cast<const uint4*>(hit_words(const_cast<uint32_t*>(hits_full), const_cast<uint32_t*>(hits_tail), tile, range.x, len, w));
            const uint32_t i4 = (wp * 32 + (qq >> 1) * 16 + (qq & 1) * 4) >> 2;
            const uint4 u = row[i4], v = row[i4 + 2];
            s_hits[buf][threadIdx.x] = u.x | u.y | u.z | u.w | v.x | v.y | v.z | v.w;
        }
    };
    S.init();
#pragma unroll
    for (int i = 0; i < NS - 1; i++) {
        if (i < batches) stage(i, i, load_id(i));
        cp_async_commit();
    }
    uint32_t id_next = load_id(NS - 1);

    int buf = 0;
    for (int k = 0; k < batches; k++) {
        stage_wait(S, buf, k);
        __syncthreads();  // publishes batch k; everyone is finished with batch k-1, whose buffer is reused below
        {
            const int nbuf = buf == 0 ? NS - 1 : buf - 1;  // (k + NS - 1) % NS
            if (k + NS - 1 < batches) stage(k + NS - 1, nbuf, id_next);
            cp_async_commit();
            id_next = load_id(k + NS);
        }
        const int kb = batches - 1 - k;
        // Every quarter-warp walks the hit words of ITS 4x2 block through the whole batch at its own pace (window after
        // window, back to front): the warp iterates max-over-quarters of the BATCH's visit counts instead of the sum over
        // windows of per-window maxima -- about 11 % fewer iterations at the headline workload.
        const int wi_lo = 0;
        // first window of the batch this quarter has to visit: the last one holding entries < n, and not past the quarter's
        // own last hit window (wq_last); negative = nothing for this quarter in the batch
        const int wtop = min(min(BLEND_BATCH / 32 - 1, (n - 1 - kb * BLEND_BATCH) >> 5), wq_last - kb * (BLEND_BATCH / 32));
        auto fetch = [&](int wv) -> uint32_t { return s_hits[buf][wv * BLOCKS + lwarp * 4 + q]; };
        const int e_last = last_contributor - kb * BLEND_BATCH;   // slots [0, e_last) of this batch are at or before this pixel's last contributor
        int wi = max(wtop, 0);
        uint32_t mask = wtop >= 0 ? fetch(wi) : 0u;
        {
            while (true) {
                if (mask == 0u && wi > wi_lo) mask = fetch(--wi);   // this quarter moves on to its next window (one per iteration)
                if (!__any_sync(0xffffffffu, mask != 0u || wi > wi_lo)) break;   // nothing queued and no window left, in any quarter
                const bool act = mask != 0;
                uint32_t eb;                                   // highest queued entry; FLO yields 0xffffffff for an empty mask -> entry 31 (read, not used)
                asm("bfind.u32 %0, %1;" : "=r"(eb) : "r"(mask));
                eb &= 31u;
                mask &= ~(1u << eb);
                const int e = wi * 32 + (int)eb;           // slot of the batch; 1-based list position = kb * BLEND_BATCH + e + 1
                const float4 A = S.A(buf, e);
                const float4 B = S.B(buf, e);
                const float4 Cc = S.C(buf, e);
                const float dx = __fsub_rn(A.x, pxf), dy = __fsub_rn(A.y, pyf);
                const float power = splat_power(dx, dy, B.x, B.y, B.z);
                // Branch-free body: a lane that does not contribute carries u = wc = 0 through the sums and
                // leaves its recurrences untouched (with exact hit words nearly every visit has contributors,
                // so skipping the arithmetic for an all-idle warp is not worth the divergence bookkeeping).
                // exp by one FMUL + MUFU.EX2 (relative error < 1e-6; gradients carry a 1e-3 tolerance).  No alpha >= 1/255 test: the record's
                // power threshold A.w is exact (preprocess.cu), so `power >= A.w` IS the forward's decision -- it must be repeated exactly,
                // one flipped pair shifts T by 0.4 % for every splat in front of it at that pixel -- whatever the exponential's rounding.
                const float G = ex2_approx(power * 1.4426950408889634f);
                const float alpha = fminf(0.99f, __fmul_rn(B.w, G));
                const bool contrib = act && e < e_last && !(power > 0.0f) && !(power < A.w);
                const float inv = rcp_approx(1.0f - alpha);                 // 1 / (1 - alpha), one MUFU.RCP
                const float Tn = T * inv;                                   // T_before = T_after / (1 - alpha)
                float cd = fmaf(Cc.x, d0, fmaf(Cc.y, d1, Cc.z * d2));       // c_k . dL/dpix
                if (CH == 5) cd = fmaf(Cc.w, d3, cd + d4);                  // colours z_cam and 1 of the depth / silhouette pass
                const float dL_dalpha = fmaf(T, cd, -s_behind) * inv;
                // raw moments of u = G * dL/dalpha; the conic / opacity / 0.5 W factors are per-Gaussian
                // constants and are applied once, in gauss_bwd.cu
                const float u = contrib ? G * dL_dalpha : 0.f;
                const float wc = contrib ? alpha * Tn : 0.f;                // d colour_out / d colour_splat
                s_behind = contrib ? fmaf(cd, wc, s_behind) : s_behind;   // a select, not wc = 0: an idle lane may have read a stale (non-finite) record
                T = contrib ? Tn : T;
                const uint32_t cb = __ballot_sync(0xffffffffu, contrib);
                if (cb == 0) continue;
                // Reduce-scatter of the 8 sums {S u dx, S u dx^2, S u dx dy, S w d_r | S u dy, S u dy^2, S u, S w d_g}
                // over the quarter-warp's 8 lanes.  Stage 1 (lane bit 2) pairs sums whose summands differ only in a
                // lane-selectable factor, so "keep" and "send" are formed directly (4 selects instead of 8):
                //   m_k = hi ? dy : dx, m_s = hi ? dx : dy:  keep {u m_k, u m_k^2}, send {u m_s, u m_s^2}
                const bool hi4 = l8 & 4;
                const float mk = hi4 ? dy : dx, ms = hi4 ? dx : dy;
                const float k0 = u * mk, s0 = u * ms;
                const float k1 = k0 * mk, s1 = s0 * ms;
                const float uxy = k0 * ms;                                  // u dx dy (symmetric)
                const float k2 = hi4 ? u : uxy, s2 = hi4 ? uxy : u;
                const float k3 = wc * dk, s3 = wc * ds;                     // dk / ds: d_r, d_g pre-swapped per lane
                float r0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 4);
                float r1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 4);
                float r2 = k2 + __shfl_xor_sync(0xffffffffu, s2, 4);
                float r3 = k3 + __shfl_xor_sync(0xffffffffu, s3, 4);
                // sums 8 (S w d_b) and, with five channels, 9 (S w d_z): lanes 0-3 of the quarter end up with 8, lanes 4-7 with 9
                float v8 = wc * d8k;
                v8 += __shfl_xor_sync(0xffffffffu, CH == 5 ? wc * d8s : v8, 4);
                {
                    const bool hi = l8 & 2;  // keep (r0, r1) on the low pair, (r2, r3) on the high pair
                    const float sa = hi ? r0 : r2, ka = hi ? r2 : r0;
                    const float sb = hi ? r1 : r3, kb2 = hi ? r3 : r1;
                    r0 = ka + __shfl_xor_sync(0xffffffffu, sa, 2);
                    r1 = kb2 + __shfl_xor_sync(0xffffffffu, sb, 2);
                    v8 += __shfl_xor_sync(0xffffffffu, v8, 2);
                }
                {
                    const bool hi = l8 & 1;
                    const float sa = hi ? r0 : r1, ka = hi ? r1 : r0;
                    r0 = ka + __shfl_xor_sync(0xffffffffu, sa, 1);
                    v8 += __shfl_xor_sync(0xffffffffu, v8, 1);
                }
                // lane l8 = (b2 b1 b0) now holds: b2 selects the {dx.. | dy..} family, b1 the pair, b0 the member:
                //   000 S u dx   001 S u dx^2   010 S u dxdy   011 S w d_r   100 S u dy   101 S u dy^2   110 S u   111 S w d_g
                if ((cb >> qshift) & 0xffu) {  // this quarter touched the splat
                    float* dst = acc + (size_t)s_ids[buf][e] * 12;
                    atomicAdd(dst + acc_slot, r0);
                    if (l8 == 0) atomicAdd(dst + 8, v8);
                    if (CH == 5 && l8 == 4) atomicAdd(dst + 9, v8);
                }
            }
        }
        buf = buf == NS - 1 ? 0 : buf + 1;
    }
    cp_async_wait<0>();
}

int launch_blend_backward(const FwdParams& p, char* geom, const GeomLayout& GL, const char* binning,
                          const char* image, const ImageLayout& IL, const float* dL_dpix, const float* dL_ddepth_sil,
                          cudaStream_t s)
{
    if (p.W <= 0 || p.H <= 0 || p.P <= 0) return GSB_OK;
#ifdef GSB_TUNING   // developer builds only: the two-phase experiment (blend_bwd_twophase.cu)
    static const bool twophase = [] { const char* e = getenv("GSB_BLEND_BWD"); return e && e[0] == 't'; }();
    if (twophase) return launch_blend_backward_twophase(p, geom, GL, binning, image, IL, dL_dpix, dL_ddepth_sil, s);
#endif
    GSB_CUDA_CHECK(cudaMemsetAsync(geom + GL.acc, 0, (size_t)p.P * sizeof(GradAcc), s));
    if (p.band_y1 <= p.band_y0) return GSB_OK;   // empty tile-row band: all-zero accumulators
    {
        StageTimer _t(ST_BLEND_BWD, s);
#define GSB_BWD_LAUNCH(MB, NS, HV) GSB_BWD_LAUNCH_CH(MB, NS, HV, 3, false)
#define GSB_BWD_LAUNCH_CH(MB, NS, HV, CH, BK)                                                                                          \
    do {                                                                                                                    \
        GSB_SET_ATTR_ONCE((blend_backward_kernel<MB, NS, HV, CH, BK>), cudaFuncAttributePreferredSharedMemoryCarveout, 100); \
        blend_backward_kernel<MB, NS, HV, CH, BK><<<dim3(IL.tiles_x, (p.band_y1 - p.band_y0) * HV), 256 / HV, 0, s>>>(                       \
            reinterpret_cast<const uint2*>(image + IL.ranges), binning, reinterpret_cast<const SplatRec*>(geom + GL.rec), p.W, \
            p.H, p.background, reinterpret_cast<const float*>(image + IL.final_T),                                          \
            reinterpret_cast<const uint32_t*>(image + IL.n_contrib), reinterpret_cast<const uint32_t*>(image + IL.tile_max_contrib), \
            dL_dpix, dL_ddepth_sil, reinterpret_cast<float*>(geom + GL.acc),                                               \
            reinterpret_cast<const uint32_t*>(image + IL.hits_tail),                                                        \
            reinterpret_cast<const GeomHeader*>(geom + GL.header), (uint32_t)p.band_y0);                                                       \
    } while (0)
#ifdef GSB_TUNING   // developer builds only (make tune): CTA shape (whole / half tile), register budget, staging depth, staging engine
        static const int halves = [] { const char* e = getenv("GSB_BLEND_BWD_HALVES"); return e ? atoi(e) : 1; }();
        static const int minb = [] { const char* e = getenv("GSB_BLEND_BWD_MINB"); return e ? atoi(e) : 0; }();
        static const int stages = [] { const char* e = getenv("GSB_BLEND_BWD_STAGES"); return e ? atoi(e) : 3; }();
        static const bool bulk = [] { const char* e = getenv("GSB_BLEND_STAGE"); return e ? e[0] == 'b' : GSB_DEFAULT_BULK; }();
        if (dL_ddepth_sil) {
            if (bulk) GSB_BWD_LAUNCH_CH(4, 3, 1, 5, true); else GSB_BWD_LAUNCH_CH(4, 3, 1, 5, false);
        } else if (bulk && halves == 1 && stages != 2 && minb != 3) {
            GSB_BWD_LAUNCH_CH(4, 3, 1, 3, true);
        } else if (halves == 2) {
            if (stages == 2) { if (minb == 8) GSB_BWD_LAUNCH(8, 2, 2); else if (minb == 10) GSB_BWD_LAUNCH(10, 2, 2); else GSB_BWD_LAUNCH(6, 2, 2); }
            else { if (minb == 8) GSB_BWD_LAUNCH(8, 3, 2); else if (minb == 10) GSB_BWD_LAUNCH(10, 3, 2); else GSB_BWD_LAUNCH(6, 3, 2); }
        } else {
            if (stages == 2) { if (minb == 3) GSB_BWD_LAUNCH(3, 2, 1); else GSB_BWD_LAUNCH(4, 2, 1); }
            else { if (minb == 3) GSB_BWD_LAUNCH(3, 3, 1); else if (minb == 5) GSB_BWD_LAUNCH(5, 3, 1); else GSB_BWD_LAUNCH(4, 3, 1); }
        }
#else
        // whole-tile CTAs, 4 resident per SM, three-deep cp.async (LDGSTS) staging ring: the measured best
        if (dL_ddepth_sil) GSB_BWD_LAUNCH_CH(4, 3, 1, 5, false);
        else GSB_BWD_LAUNCH_CH(4, 3, 1, 3, false);
#endif
#undef GSB_BWD_LAUNCH
#undef GSB_BWD_LAUNCH_CH
        GSB_LAUNCH_CHECK();
    }
    return GSB_OK;
}

}  // namespace gsb
