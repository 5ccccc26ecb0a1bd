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
//   * no cull pass at all: the forward pass records, per window of 32 list entries and per 4x2
//     block, the bit mask of the entries that were blended into at least one pixel of the block
//     ("hit words", BinningLayout::hits).  The backward stages the 256 hit words of a batch with
//     the records (one 4-byte cp.async per thread) and every quarter-warp walks exactly the
//     entries that hit ITS block, back to front -- 25 % fewer (block, splat) visits than the
//     conservative footprint test and no per-batch ballots;
//   * the walk starts at the tile's highest n_contrib (recorded by the forward pass), not at
//     the end of the tile's list;
//   * records are gathered with cp.async into a ring of shared-memory buffers with one CTA
//     barrier per batch (stage.cuh).
#include <cstdlib>
#include "common.cuh"
#include "stage.cuh"

namespace gsb {

template <int MINB, int NS>
__global__ void __launch_bounds__(BLEND_THREADS, MINB)
blend_backward_kernel(const uint2* __restrict__ ranges, const char* __restrict__ binning,
                      const SplatRec* __restrict__ rec, int W, int H, const float* __restrict__ bg,
                      const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
                      const uint32_t* __restrict__ tile_max_contrib, const float* __restrict__ dL_dpix,
                      float* __restrict__ acc /* [P][12] */, const uint32_t* __restrict__ hits_tail,
                      const GeomHeader* __restrict__ hdr)
{
    __shared__ StageBuf<NS> S;
    __shared__ uint32_t s_ids[NS][BLEND_BATCH];
    __shared__ uint32_t s_hits[NS][BLEND_BATCH];  // [window of the batch][4x2 block]
    const uint32_t tile = blockIdx.y * gridDim.x + blockIdx.x;
    const uint2 range = ranges[tile];
    const uint32_t len = range.y - range.x;
    const int n = min((int)len, (int)tile_max_contrib[tile]);  // entries [0, n) can matter
    const int batches = (n + BLEND_BATCH - 1) / BLEND_BATCH;
    if (batches == 0) return;
    const BinningLayout BL = BinningLayout::make((long long)hdr->layout_capacity);
    const uint32_t* point_list = reinterpret_cast<const uint32_t*>(binning + BL.point_list);
    const uint32_t* hits_full = reinterpret_cast<const uint32_t*>(binning + BL.hits);
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t q = lane >> 3, l8 = lane & 7, qshift = q * 8;
    const int bx0 = blockIdx.x * TILE_X + (warp & 1) * 8, by0 = blockIdx.y * TILE_Y + (warp >> 1) * 4;
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
    const float bg_dot_dpixel = __ldg(bg) * d0 + __ldg(bg + 1) * d1 + __ldg(bg + 2) * d2;
    float ar0 = 0.f, ar1 = 0.f, ar2 = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, last_alpha = 0.f;
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
        stage_issue(S, buf, rec, id);
        const uint32_t w = (uint32_t)(batches - 1 - k) * (BLEND_BATCH / 32) + (threadIdx.x >> 5);
        if ((int)(w * 32) < n)
            cp_async4(&s_hits[buf][threadIdx.x], hit_word(const_cast<uint32_t*>(hits_full), const_cast<uint32_t*>(hits_tail), tile,
                                                          range.x, len, w, threadIdx.x & 31));
    };
#pragma unroll
    for (int i = 0; i < NS - 1; i++) {
        if (i < batches) stage(i, i, load_id(i));
        cp_async_commit();
    }
    uint32_t id_next = load_id(NS - 1);

    int buf = 0;
    for (int k = 0; k < batches; k++) {
        cp_async_wait<NS - 2>();
        __syncthreads();  // publishes batch k; everyone is finished with batch k-1, whose buffer is reused below
        {
            const int nbuf = buf == 0 ? NS - 1 : buf - 1;  // (k + NS - 1) % NS
            if (k + NS - 1 < batches) stage(k + NS - 1, nbuf, id_next);
            cp_async_commit();
            id_next = load_id(k + NS);
        }
        const int kb = batches - 1 - k;
#pragma unroll 1
        for (int wi = BLEND_BATCH / 32 - 1; wi >= 0; wi--) {
            const int w = kb * (BLEND_BATCH / 32) + wi;
            if (w * 32 >= n) continue;
            uint32_t mask = w <= wq_last ? s_hits[buf][wi * 32 + warp * 4 + q] : 0u;
            // ---- gradient pass: every quarter-warp walks the entries that hit ITS block, back to front ----
            while (__any_sync(0xffffffffu, mask != 0)) {
                const bool act = mask != 0;
                const int eb = act ? 31 - __clz(mask) : 0;
                mask &= ~(1u << eb);
                const int e = wi * 32 + eb;
                const int pos = kb * BLEND_BATCH + e + 1;  // 1-based list position
                const float4 A = S.a[buf][e];
                const float4 B = S.b[buf][e];
                const float dx = __fsub_rn(A.x, pxf), dy = __fsub_rn(A.y, pyf);
                const float power = splat_power(dx, dy, B.x, B.y, B.z);
                float v[8], v8 = 0.f;
#pragma unroll
                for (int kk = 0; kk < 8; kk++) v[kk] = 0.f;
                bool contrib = false;
                if (act && pos <= last_contributor && !(power > 0.0f) && !(power < A.w)) {
                    const float G = expf(power);
                    const float alpha = fminf(0.99f, __fmul_rn(B.w, G));
                    if (!(alpha < 1.0f / 255.0f)) {
                        contrib = true;
                        const float4 Cc = S.c[buf][e];
                        const float inv = __fdividef(1.0f, 1.0f - alpha);   // 1 / (1 - alpha), MUFU.RCP
                        T *= inv;                                           // T_before = T_after / (1 - alpha)
                        const float wc = alpha * T;                         // d colour_out / d colour_splat
                        const float om = 1.f - last_alpha;
                        ar0 = fmaf(last_alpha, lc0, om * ar0);              // colour accumulated behind this splat
                        ar1 = fmaf(last_alpha, lc1, om * ar1);
                        ar2 = fmaf(last_alpha, lc2, om * ar2);
                        lc0 = Cc.x; lc1 = Cc.y; lc2 = Cc.z;
                        last_alpha = alpha;
                        float dL_dalpha = T * fmaf(Cc.x - ar0, d0, fmaf(Cc.y - ar1, d1, (Cc.z - ar2) * d2));
                        dL_dalpha = fmaf(-T_final * inv, bg_dot_dpixel, dL_dalpha);
                        // raw moments of u = G * dL/dalpha; the conic / opacity / 0.5 W factors are
                        // per-Gaussian constants and are applied once, in gauss_bwd.cu
                        const float u = G * dL_dalpha, ux = u * dx, uy = u * dy;
                        v[0] = ux;
                        v[1] = uy;
                        v[2] = ux * dx;
                        v[3] = ux * dy;
                        v[4] = uy * dy;
                        v[5] = u;
                        v[6] = wc * d0;
                        v[7] = wc * d1;
                        v8 = wc * d2;
                    }
                }
                const uint32_t cb = __ballot_sync(0xffffffffu, contrib);
                if (cb == 0) continue;
                // reduce-scatter the 8 sums over the quarter-warp's 8 lanes: 4 + 2 + 1 shuffles,
                // after which lane l8 holds sum number (bit2, bit1, bit0 of l8) complete
                {
                    const bool hi = l8 & 4;
#pragma unroll
                    for (int kk = 0; kk < 4; kk++) {
                        const float send = hi ? v[kk] : v[kk + 4];
                        const float keep = hi ? v[kk + 4] : v[kk];
                        v[kk] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                    }
                }
                {
                    const bool hi = l8 & 2;
#pragma unroll
                    for (int kk = 0; kk < 2; kk++) {
                        const float send = hi ? v[kk] : v[kk + 2];
                        const float keep = hi ? v[kk + 2] : v[kk];
                        v[kk] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
                    }
                }
                {
                    const bool hi = l8 & 1;
                    const float send = hi ? v[0] : v[1];
                    const float keep = hi ? v[1] : v[0];
                    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
                }
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) v8 += __shfl_xor_sync(0xffffffffu, v8, o);
                if ((cb >> qshift) & 0xffu) {  // this quarter touched the splat
                    float* dst = acc + (size_t)s_ids[buf][e] * 12;
                    const int kk = ((l8 >> 2) & 1) * 4 + ((l8 >> 1) & 1) * 2 + (l8 & 1);
                    atomicAdd(dst + kk, v[0]);
                    if (l8 == 0) atomicAdd(dst + 8, v8);
                }
            }
        }
        buf = buf == NS - 1 ? 0 : buf + 1;
    }
    cp_async_wait<0>();
}

int launch_blend_backward(const FwdParams& p, char* geom, const GeomLayout& GL, const char* binning,
                          const char* image, const ImageLayout& IL, const float* dL_dpix, cudaStream_t s)
{
    if (p.W <= 0 || p.H <= 0 || p.P <= 0) return GSB_OK;
    GSB_CUDA_CHECK(cudaMemsetAsync(geom + GL.acc, 0, (size_t)p.P * sizeof(GradAcc), s));
    // tuning knobs: resident CTAs per SM the compiler must allow (the register budget) and the depth of the staging ring
    static const int minb = [] { const char* e = getenv("GSB_BLEND_BWD_MINB"); return e ? atoi(e) : 3; }();
    static const int stages = [] { const char* e = getenv("GSB_BLEND_BWD_STAGES"); return e ? atoi(e) : 3; }();
    dim3 grid(IL.tiles_x, IL.tiles_y);
    {
        StageTimer _t(ST_BLEND_BWD, s);
#define GSB_BWD_LAUNCH(MB, NS)                                                                                              \
    do {                                                                                                                    \
        static const bool attr_set = [] {                                                                                   \
            cudaFuncSetAttribute(blend_backward_kernel<MB, NS>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);       \
            return true;                                                                                                    \
        }();                                                                                                                \
        (void)attr_set;                                                                                                     \
        blend_backward_kernel<MB, NS><<<grid, BLEND_THREADS, 0, s>>>(                                                       \
            reinterpret_cast<const uint2*>(image + IL.ranges), binning, reinterpret_cast<const SplatRec*>(geom + GL.rec), p.W, \
            p.H, p.background, reinterpret_cast<const float*>(image + IL.final_T),                                          \
            reinterpret_cast<const uint32_t*>(image + IL.n_contrib), reinterpret_cast<const uint32_t*>(image + IL.tile_max_contrib), \
            dL_dpix, reinterpret_cast<float*>(geom + GL.acc), reinterpret_cast<const uint32_t*>(image + IL.hits_tail),      \
            reinterpret_cast<const GeomHeader*>(geom + GL.header));                                                         \
    } while (0)
        if (stages == 2) {
            if (minb == 3) GSB_BWD_LAUNCH(3, 2); else if (minb == 5) GSB_BWD_LAUNCH(5, 2); else GSB_BWD_LAUNCH(4, 2);
        } else {
            if (minb == 3) GSB_BWD_LAUNCH(3, 3); else if (minb == 5) GSB_BWD_LAUNCH(5, 3); else GSB_BWD_LAUNCH(4, 3);
        }
#undef GSB_BWD_LAUNCH
        GSB_LAUNCH_CHECK();
    }
    return GSB_OK;
}

}  // namespace gsb
