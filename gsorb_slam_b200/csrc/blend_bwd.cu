// blend_bwd.cu -- per-tile back-to-front gradient of the alpha blend (K7) for sm_100a.
//
// Replaces BACKWARD::render / renderCUDA<3> (backward.cu:399-557, launch :641-656).
// Same result as the reference's recurrences: start from the stored final transmittance and the
// position of the last blended splat, walk the tile's list backwards, T <- T / (1 - alpha),
// colour accumulated behind the splat, background term, gradients w.r.t. colour, 2D mean
// (in NDC units: x 0.5 W, x 0.5 H), conic (slots x, y, w) and opacity.  No depth gradient.
// Per pair only u = G dL/dalpha and its first / second moments in (dx, dy) are formed; the
// linear maps from those 6 sums to d(mean2D), d(conic), d(opacity) use per-Gaussian constants
// and are applied once per Gaussian in gauss_bwd.cu (packed accumulator layout, 12 floats:
// {S u dx, S u dx^2, S u dx dy, S w d_r | S u dy, S u dy^2, S u, S w d_g | S w d_b, S w d_z}).
//
// What is different (design, not results):
//   * the reference issues 9 global atomicAdds per contributing (pixel, splat) pair; here a GROUP of 4 lanes owns a 2x2 pixel
//     block (8 groups per warp; a 2-lane / 2x1-pixel variant is a tune-build knob), reduces the sums of a splat across its lanes
//     with a two-stage shuffle reduce-scatter and issues one 8-byte vector RED per lane (+ one scalar RED per group) into one
//     48-byte packed accumulator per Gaussian, and only for splats that touched at least one pixel of the block; the eight groups
//     of a warp work on eight different splats at once (round 1 used 8-lane groups on 4x2 blocks: 128 iterations per warp and
//     tile at the headline workload against 110 now, tests/decomposition_model.py);
//   * the per-channel "colour behind" recurrences of the reference (accum_rec / last_color / last_alpha, 7 registers of state,
//     ~20 instructions per pair) are ONE scalar: s = dL/dpix . (sum of c_j alpha_j T_j over the splats behind) + T_final bg . dL/dpix,
//     with dL/dalpha_k = (T_after c_k . dL/dpix - s) / (1 - alpha_k);
//   * no cull pass at all: the forward pass records, per window of 32 list entries and per PIXEL, the bit mask of the
//     entries that were blended into the pixel ("hit words", BinningLayout::hits).  The backward ORs the words of a group's pixels
//     while it stages a batch (two LDG.128 per thread) and every group walks exactly the entries that hit ITS block, back to
//     front, through the whole batch at its own pace;
//   * the alpha >= 1/255 decision is `power >= A.w`: preprocess.cu stores the EXACT power threshold of each Gaussian, so the
//     exponential here can be one FMUL + MUFU.EX2 without ever disagreeing with the forward pass about which pairs were blended;
//   * the walk starts at the tile's highest n_contrib (recorded by the forward pass), not at the end of the tile's list;
//   * records are gathered with cp.async into a ring of shared-memory buffers with one CTA barrier per batch (stage.cuh).
#include <cstdlib>
#include "common.cuh"
#include "stage.cuh"

namespace gsb {

#ifndef GSB_BWD_SPLIT_BARRIER
#define GSB_BWD_SPLIT_BARRIER false   // split (arrive / sync) batch barrier of the backward walk, see the kernel
#endif
#ifndef GSB_BWD_GROUP_LANES
#define GSB_BWD_GROUP_LANES 4   // lanes per group of the backward walk: 4 = 2x2 pixels, 2 = 2x1 pixels
#endif

__device__ __forceinline__ void red_add_v2(float* dst, float x, float y)
{
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dst), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* dst, float x, float y, float z, float w)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

template <int NS, int GL, bool BULK>
struct BwdShared {
    StageRing<NS, 256, BULK> ring;
    uint32_t ids[NS][256];
    alignas(8) uint32_t hits[NS][8 * 8 * (32 / GL)];   // [window of the batch][lane group of the tile]
};

// GL = lanes per group: 4 (a group owns 2x2 pixels) or 2 (2x1 pixels).  CH = 5: backward of the fused RGB + depth / silhouette pass
// (see blend_fwd.cu): dL/dalpha sums over five channels, and the gradient of the z_cam colour (sum of w * dL/dpix[3]) lands in
// accumulator slot 9.
template <int MINB, int NS, int GL, int CH, bool BULK, bool SPLIT>
__global__ void __launch_bounds__(256, MINB)
blend_backward_kernel(const uint2* __restrict__ ranges, const char* __restrict__ binning,
                      const SplatRec* __restrict__ rec, int W, int H, const float* __restrict__ bg,
                      const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
                      const uint32_t* __restrict__ tile_max_contrib, const float* __restrict__ dL_dpix,
                      const float* __restrict__ dL_ddepth_sil, float* __restrict__ acc /* [P][16] */, const uint32_t* __restrict__ hits_tail,
                      const GeomHeader* __restrict__ hdr, uint32_t band_y0)
{
    static_assert(GL == 4 || GL == 2, "lane groups of 4 (2x2 pixels) or 2 (2x1 pixels)");
    constexpr int BLEND_THREADS = 256, BLEND_BATCH = BLEND_THREADS, WINS = BLEND_BATCH / 32;
    constexpr int GPW = 32 / GL;        // lane groups per warp (4 columns of two pixels x 2 or 4 rows)
    constexpr int GROUPS = 8 * GPW;     // lane groups per tile
    // dynamic shared memory (three stages of the 2-lane variant exceed the 48 KB static limit)
    extern __shared__ __align__(16) unsigned char bwd_smem_raw[];
    using Smem = BwdShared<NS, GL, BULK>;
    Smem& SM = *reinterpret_cast<Smem*>(bwd_smem_raw);
    StageRing<NS, BLEND_BATCH, BULK>& S = SM.ring;
    uint32_t (&s_ids)[NS][BLEND_BATCH] = SM.ids;
    uint32_t (&s_hits)[NS][WINS * GROUPS] = SM.hits;
    const uint32_t tile_y = band_y0 + blockIdx.y;   // the grid covers the band's tile rows
    const uint32_t tile = tile_y * gridDim.x + blockIdx.x;
    const uint2 range = ranges[tile];
    const uint32_t len = range.y - range.x;
    // entries [0, n) can matter: n = highest n_contrib over the tile's pixels (recorded by the forward)
    const uint32_t tmax = max(tile_max_contrib[2 * tile], tile_max_contrib[2 * tile + 1]);
    const int n = min((int)len, (int)tmax);
    const int batches = (n + BLEND_BATCH - 1) / BLEND_BATCH;
    if (batches == 0) return;
    const BinningLayout BL = BinningLayout::make((long long)hdr->layout_capacity);
    const uint32_t* point_list = reinterpret_cast<const uint32_t*>(binning + BL.point_list);
    const uint32_t* hits_full = reinterpret_cast<const uint32_t*>(binning + BL.hits);
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();   // warp: 8x4 pixel region of the tile, as in the forward
    const uint32_t g = lane / GL, l = lane % GL, gshift = g * GL;
    const int bx0 = blockIdx.x * TILE_X + (warp & 1) * 8, by0 = tile_y * TILE_Y + (warp >> 1) * 4;
    const int px = bx0 + (g & 3) * 2 + (l & 1), py = by0 + (GL == 4 ? (g >> 2) * 2 + (l >> 1) : (g >> 2));
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
    // The first reduction stage pairs the lanes that differ in the group's top lane bit: the `lo` lane keeps the dx family of sums
    // {S u dx, S u dx^2, S u dx dy, S w d_r}, the `hi` lane the dy family {S u dy, S u dy^2, S u, S w d_g}.
    const bool hi = l & (GL / 2);
    const float d8k = (CH == 5 && hi) ? d3 : d2, d8s = hi ? d2 : d3;   // keep / send factors of sums 8 (S w d_b) and 9 (S w d_z)
    const float dk = hi ? d1 : d0, ds = hi ? d0 : d1;                   // keep / send factors of the colour sums
    // Colour seen behind the current splat, as ONE scalar: s = dL/dpix . (sum of c_j alpha_j T_j over the splats j already walked)
    // + T_final bg . dL/dpix.  With T_k(1 - alpha_k) = T_after the reference's T_k (c_k - accum_rec) . dL/dpix - T_final/(1 - alpha_k) bg . dL/dpix
    // (backward.cu:505-535: per-channel accum_rec / last_color recurrences) equals (T_k c_k . dL/dpix - s) / (1 - alpha_k).
    float s_behind = T_final * bg_dot_dpixel;
    // last window (of 32 list entries) in which this lane group's pixels blended anything: the forward
    // pass wrote hit words for every window up to it
    int qmax = last_contributor;
#pragma unroll
    for (int o = GL / 2; o > 0; o >>= 1) qmax = max(qmax, __shfl_xor_sync(0xffffffffu, qmax, o));
    const int wq_last = (qmax - 1) >> 5;  // -1 when the group never blended

    // batches are walked back to front; inside a batch slot t holds list entry kb*256 + t
    const uint32_t* ids = point_list + range.x;
    auto load_id = [&](int k) -> uint32_t {  // k-th batch in walking order
        const int i = (batches - 1 - k) * BLEND_BATCH + (int)threadIdx.x;
        return (k < batches && i < n) ? __ldg(ids + i) : 0xffffffffu;
    };
    auto stage = [&](int k, int buf, uint32_t id) {
        s_ids[buf][threadIdx.x] = id;
        stage_issue(S, buf, rec, id, min(BLEND_BATCH, n - (batches - 1 - k) * BLEND_BATCH));
        // hit words of the batch: the forward wrote one word per (window, pixel); a thread ORs the words of one 4x2 pixel block of one
        // window (two LDG.128: rows of the forward's [warp][row of 8][column] pixel order) into the words of the block's lane groups
        const uint32_t wv = threadIdx.x >> 5, w = (uint32_t)(batches - 1 - k) * WINS + wv;
        if ((int)(w * 32) < n) {
            const uint32_t blk = threadIdx.x & 31, wp = blk >> 2, qq = blk & 3;
            const uint4* row = reinterpret_cast<const uint4*>(hit_words(const_cast<uint32_t*>(hits_full), const_cast<uint32_t*>(hits_tail), tile, range.x, len, w));
            const uint32_t i4 = (wp * 32 + (qq >> 1) * 16 + (qq & 1) * 4) >> 2;
            const uint4 u = row[i4], v = row[i4 + 2];   // pixel rows 2 (qq >> 1) and 2 (qq >> 1) + 1, columns 4 (qq & 1) .. + 3
            uint32_t* dst = &s_hits[buf][wv * GROUPS + wp * GPW + (qq & 1) * 2];
            if (GL == 4) {
                *reinterpret_cast<uint2*>(dst + (qq >> 1) * 4) = make_uint2(u.x | u.y | v.x | v.y, u.z | u.w | v.z | v.w);
            } else {
                *reinterpret_cast<uint2*>(dst + (qq >> 1) * 8) = make_uint2(u.x | u.y, u.z | u.w);
                *reinterpret_cast<uint2*>(dst + (qq >> 1) * 8 + 4) = make_uint2(v.x | v.y, v.z | v.w);
            }
        }
    };
    // Pipeline.  SPLIT = false: batches k+1 .. k+NS-1 in flight while batch k is walked, ONE CTA barrier per batch (publishes batch k
    // and proves everyone is finished with batch k-1, whose buffer is refilled right after it).
    // SPLIT = true (three buffers, one batch in flight): the barrier of batch k+1 is split into an ARRIVE that a warp issues
    // a few walk iterations into batch k -- its own copies of batch k+1, issued at the top of batch k, have landed by then -- and the
    // SYNC at the top of batch k+1 (named barriers 1..3, one per buffer; 256 arrivals + 256 syncs per phase).  A warp that is
    // ahead therefore only waits for the others to be a few iterations into the PREVIOUS batch, not to have finished it; buffer
    // (k+1) % 3 = (k-2) % 3 is refilled at the top of batch k, which every thread reaches only after all warps left batch k-2.
    static_assert(!SPLIT || (NS == 3 && !BULK), "the split barrier is written for the three-deep cp.async ring");
    constexpr int ARRIVE_AT = 6;   // walk iterations into a batch after which a warp reports its copies of the next one
    // (immediate barrier ids: a register operand makes ptxas reserve all 16 named barriers of the CTA)
    auto bar_arrive = [](int b) {
        if (b == 0) asm volatile("barrier.cta.arrive 1, 512;" ::: "memory");
        else if (b == 1) asm volatile("barrier.cta.arrive 2, 512;" ::: "memory");
        else asm volatile("barrier.cta.arrive 3, 512;" ::: "memory");
    };
    auto bar_sync = [](int b) {
        if (b == 0) asm volatile("barrier.cta.sync 1, 512;" ::: "memory");
        else if (b == 1) asm volatile("barrier.cta.sync 2, 512;" ::: "memory");
        else asm volatile("barrier.cta.sync 3, 512;" ::: "memory");
    };
    S.init();
    uint32_t id_next;
    if (SPLIT) {
        stage(0, 0, load_id(0));
        cp_async_commit();
        id_next = load_id(1);
        cp_async_wait<0>();
        bar_arrive(0);
    } else {
#pragma unroll
        for (int i = 0; i < NS - 1; i++) {
            if (i < batches) stage(i, i, load_id(i));
            cp_async_commit();
        }
        id_next = load_id(NS - 1);
    }

    int buf = 0;
    for (int k = 0; k < batches; k++) {
        const int nbuf = SPLIT ? (buf == NS - 1 ? 0 : buf + 1) : (buf == 0 ? NS - 1 : buf - 1);   // buffer of batch k + 1 / k + NS - 1
        if (SPLIT) {
            bar_sync(buf);   // batch k is complete; every warp is at least ARRIVE_AT iterations into batch k-1
            if (k + 1 < batches) stage(k + 1, nbuf, id_next);
            cp_async_commit();
            id_next = load_id(k + 2);
        } else {
            stage_wait(S, buf, k);
            __syncthreads();  // publishes batch k; everyone is finished with batch k-1, whose buffer is reused below
            if (k + NS - 1 < batches) stage(k + NS - 1, nbuf, id_next);
            cp_async_commit();
            id_next = load_id(k + NS);
        }
        const bool report = SPLIT && k + 1 < batches;   // this warp owes the arrival for batch k + 1
        int it = 0;
        const int kb = batches - 1 - k;
        // Every lane group walks the hit words of ITS pixels through the whole batch at its own pace (window after
        // window, back to front): the warp iterates max-over-groups of the BATCH's visit counts.
        // first window of the batch this group has to visit: the last one holding entries < n, and not past the group's
        // own last hit window (wq_last); negative = nothing for this group in the batch
        const int wtop = min(min(WINS - 1, (n - 1 - kb * BLEND_BATCH) >> 5), wq_last - kb * WINS);
        const uint32_t* const hrow = &s_hits[buf][warp * GPW + g];
        const int e_last = last_contributor - kb * BLEND_BATCH;   // slots [0, e_last) of this batch are at or before this pixel's last contributor
        int wi = max(wtop, 0);
        uint32_t mask = wtop >= 0 ? hrow[wi * GROUPS] : 0u;
        uint32_t nmask = wi > 0 ? hrow[(wi - 1) * GROUPS] : 0u;   // hit word of the group's NEXT window, fetched one window ahead
        float* const acc_l = acc + (GL == 4 ? 4 * l : 8 * l);     // this lane's float4 of the accumulator record (GL = 2: its two float4)
        while (true) {
            if (report && it == ARRIVE_AT) {
                cp_async_wait<0>();
                bar_arrive(nbuf);
            }
            it++;
            {   // a group whose window is exhausted moves on to its next one (one per iteration); written with selects: no divergence
                const bool adv = mask == 0u && wi > 0;
                mask = adv ? nmask : mask;
                wi -= adv ? 1 : 0;
                const uint32_t* nxt = hrow + (wi > 0 ? wi - 1 : 0) * GROUPS;   // always a mapped word
                const uint32_t fetched = *nxt;
                nmask = adv ? fetched : nmask;
            }
            if (!__any_sync(0xffffffffu, mask != 0u || wi > 0)) break;   // nothing queued and no window left, in any group
            const bool act = mask != 0;
            uint32_t eb;                                   // highest queued entry; FLO yields 0xffffffff for an empty mask -> entry 31 (read, not used)
            asm("bfind.u32 %0, %1;" : "=r"(eb) : "r"(mask));
            eb &= 31u;
            mask &= ~(1u << eb);
            const int e = wi * 32 + (int)eb;           // slot of the batch; 1-based list position = kb * BLEND_BATCH + e + 1
            const float4 A = S.A(buf, e);
            const float4 B = S.B(buf, e);
            const float4 Cc = S.C(buf, e);
            const uint32_t gid = s_ids[buf][e];        // the Gaussian's id, needed by the REDs at the end of the iteration
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
            if (cb != 0) {   // some lane of the warp blended the splat it visited
                // Reduce-scatter of the 8 sums {S u dx, S u dx^2, S u dx dy, S w d_r | S u dy, S u dy^2, S u, S w d_g} over the group's
                // lanes.  Stage 1 pairs sums whose summands differ only in a lane-selectable factor, so "keep" and "send" are formed
                // directly (4 selects instead of 8):  m_k = hi ? dy : dx, m_s = hi ? dx : dy:  keep {u m_k, u m_k^2}, send {u m_s, u m_s^2}
                const float mk = hi ? dy : dx, ms = hi ? dx : dy;
                const float k0 = u * mk, s0 = u * ms;
                const float k1 = k0 * mk, s1 = s0 * ms;
                const float uxy = k0 * ms;                                  // u dx dy (symmetric)
                const float k2 = hi ? u : uxy, s2 = hi ? uxy : u;
                const float k3 = wc * dk, s3 = wc * ds;                     // dk / ds: d_r, d_g pre-swapped per lane
                float r0 = k0 + __shfl_xor_sync(0xffffffffu, s0, GL / 2);
                float r1 = k1 + __shfl_xor_sync(0xffffffffu, s1, GL / 2);
                float r2 = k2 + __shfl_xor_sync(0xffffffffu, s2, GL / 2);
                float r3 = k3 + __shfl_xor_sync(0xffffffffu, s3, GL / 2);
                // sums 8 (S w d_b) and, with five channels, 9 (S w d_z): the lo lanes end up with 8, the hi lanes with 9
                float v8 = wc * d8k;
                v8 += __shfl_xor_sync(0xffffffffu, CH == 5 ? wc * d8s : v8, GL / 2);
                const bool touched = (cb >> gshift) & ((1u << GL) - 1u);   // this group blended the splat into at least one of its pixels
                // accumulator record (GradAcc, 16 floats, read by gauss_bwd.cu): lane l of a 4-lane group owns floats 4 l .. 4 l + 3
                if (GL == 4) {
                    const bool b1 = l & 1;   // second stage: the even lane keeps (r0, r1), the odd lane (r2, r3)
                    const float sa = b1 ? r0 : r2, ka = b1 ? r2 : r0;
                    const float sb = b1 ? r1 : r3, kb2 = b1 ? r3 : r1;
                    r0 = ka + __shfl_xor_sync(0xffffffffu, sa, 1);
                    r1 = kb2 + __shfl_xor_sync(0xffffffffu, sb, 1);
                    v8 += __shfl_xor_sync(0xffffffffu, v8, 1);
                    // lane 0: {S u dx, S u dx^2, S w d_b, 0}  lane 1: {S u dxdy, S w d_r, 0, 0}  lane 2: {S u dy, S u dy^2, S w d_z | 0, 0}
                    // lane 3: {S u, S w d_g, 0, 0} -- ONE 16-byte RED per lane and visit
                    const float third = (l == 0 || (CH == 5 && l == 2)) ? v8 : 0.f;
                    if (touched) red_add_v4(acc_l + (size_t)gid * ACC_FLOATS, r0, r1, third, 0.f);
                } else {
                    if (touched) {   // lo lane: floats 0, 1 | 4, 5 (+ 2);  hi lane: floats 8, 9 | 12, 13 (+ 10 with five channels)
                        float* dst = acc_l + (size_t)gid * ACC_FLOATS;
                        red_add_v2(dst, r0, r1);
                        red_add_v2(dst + 4, r2, r3);
                        if (CH == 5 || l == 0) atomicAdd(dst + 2, v8);
                    }
                }
            }
        }
        if (report && it <= ARRIVE_AT) {   // a short walk: report now
            cp_async_wait<0>();
            bar_arrive(nbuf);
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
#define GSB_BWD_LAUNCH_CH(MB, NS, GLN, CH, BK, SP)                                                                                          \
    do {                                                                                                                    \
        GSB_SET_ATTR_ONCE((blend_backward_kernel<MB, NS, GLN, CH, BK, SP>), cudaFuncAttributePreferredSharedMemoryCarveout, 100); \
        GSB_SET_ATTR_ONCE((blend_backward_kernel<MB, NS, GLN, CH, BK, SP>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdShared<NS, GLN, BK>)); \
        blend_backward_kernel<MB, NS, GLN, CH, BK, SP><<<dim3(IL.tiles_x, p.band_y1 - p.band_y0), 256, sizeof(BwdShared<NS, GLN, BK>), s>>>(                       \
            reinterpret_cast<const uint2*>(image + IL.ranges), binning, reinterpret_cast<const SplatRec*>(geom + GL.rec), p.W, \
            p.H, p.background, reinterpret_cast<const float*>(image + IL.final_T),                                          \
            reinterpret_cast<const uint32_t*>(image + IL.n_contrib), reinterpret_cast<const uint32_t*>(image + IL.tile_max_contrib), \
            dL_dpix, dL_ddepth_sil, reinterpret_cast<float*>(geom + GL.acc),                                               \
            reinterpret_cast<const uint32_t*>(image + IL.hits_tail),                                                        \
            reinterpret_cast<const GeomHeader*>(geom + GL.header), (uint32_t)p.band_y0);                                                       \
    } while (0)
#define GSB_BWD_LAUNCH(MB, NS, GLN, SP) do { if (dL_ddepth_sil) GSB_BWD_LAUNCH_CH(MB, NS, GLN, 5, false, SP); else GSB_BWD_LAUNCH_CH(MB, NS, GLN, 3, false, SP); } while (0)
#ifdef GSB_TUNING   // developer builds only (make tune): lanes per group, register budget, staging depth, staging engine, barrier kind
        static const int glanes = [] { const char* e = getenv("GSB_BLEND_BWD_GROUP"); return e ? atoi(e) : GSB_BWD_GROUP_LANES; }();
        static const int minb = [] { const char* e = getenv("GSB_BLEND_BWD_MINB"); return e ? atoi(e) : 4; }();
        static const int stages = [] { const char* e = getenv("GSB_BLEND_BWD_STAGES"); return e ? atoi(e) : 3; }();
        static const bool bulk = [] { const char* e = getenv("GSB_BLEND_STAGE"); return e ? e[0] == 'b' : GSB_DEFAULT_BULK; }();
        static const bool split = [] { const char* e = getenv("GSB_BLEND_BWD_SPLIT"); return e ? atoi(e) != 0 : GSB_BWD_SPLIT_BARRIER; }();
        if (bulk) { if (dL_ddepth_sil) GSB_BWD_LAUNCH_CH(4, 3, GSB_BWD_GROUP_LANES, 5, true, false); else GSB_BWD_LAUNCH_CH(4, 3, GSB_BWD_GROUP_LANES, 3, true, false); }
        else if (glanes == 2) { if (minb == 5) GSB_BWD_LAUNCH(5, 3, 2, false); else if (stages == 2) GSB_BWD_LAUNCH(4, 2, 2, false); else GSB_BWD_LAUNCH(4, 3, 2, false); }
        else if (split) { if (minb == 5) GSB_BWD_LAUNCH(5, 3, 4, true); else GSB_BWD_LAUNCH(4, 3, 4, true); }
        else { if (minb == 5) GSB_BWD_LAUNCH(5, 3, 4, false); else if (minb == 3) GSB_BWD_LAUNCH(3, 3, 4, false); else if (stages == 2) GSB_BWD_LAUNCH(4, 2, 4, false); else GSB_BWD_LAUNCH(4, 3, 4, false); }
#else
        // 4 resident CTAs per SM, three-deep cp.async (LDGSTS) staging ring: the measured best
        GSB_BWD_LAUNCH(4, 3, GSB_BWD_GROUP_LANES, GSB_BWD_SPLIT_BARRIER);
#endif
#undef GSB_BWD_LAUNCH
#undef GSB_BWD_LAUNCH_CH
        GSB_LAUNCH_CHECK();
    }
    return GSB_OK;
}

}  // namespace gsb
