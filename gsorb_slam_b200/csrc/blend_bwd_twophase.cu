// blend_bwd_twophase.cu -- EXPERIMENT (developer builds only: `make tune`, GSB_BLEND_BWD=twophase): a two-phase decomposition of
// the backward blend.  Correct (passes every parity test) but SLOWER than blend_bwd.cu on a B200: 428 us against 272 us at the
// headline workload (231 M warp-instructions, 50 % issue-active, 3.9 warps per issue stalled at the phase barriers, 3 CTAs per SM
// because of its 71 KB of shared memory) -- profiles/r02_blend_experiments.md.  Kept as the record of what was tried.
//
// Per-tile back-to-front gradient of the alpha blend (K7) for sm_100a.
//
// Replaces BACKWARD::render / renderCUDA<3> (backward.cu:399-557, launch :641-656).
// Same recurrences as the reference: start from the stored final transmittance and the
// position of the last blended splat, walk the tile's list backwards, T <- T / (1 - alpha),
// running "colour behind" accumulator, background term, gradients w.r.t. colour, 2D mean
// (in NDC units: x 0.5 W, x 0.5 H), conic (slots x, y, w) and opacity.  No depth gradient.
// Per pair only u = G dL/dalpha, w = alpha T and the moments of u in (dx, dy) are formed; the
// linear maps from the sums to d(mean2D), d(conic), d(opacity) use per-Gaussian constants
// and are applied once per Gaussian in gauss_bwd.cu (packed accumulator layout:
// {S u dx, S u dy, S u dx^2, S u dx dy, S u dy^2, S u, S w d_r, S w d_g, S w d_b, S w d_z}).
//
// Work decomposition (round 2).  The backward has two incompatible needs: the recurrences run PER PIXEL, back to
// front, while the gradient sums run PER SPLAT, over the pixels it was blended into.  The reference resolves that with
// 9 global atomics per (pixel, splat) pair; the round-1 kernel gave a 4x2 pixel block to a quarter-warp and reduced
// with shuffles (37 % of its lane slots carried a pair, 20 k warp-instructions per warp).  Here the two needs get one
// phase each, per batch of 256 list entries, joined through shared memory:
//   * the forward pass left one hit word per (window of 32 entries, pixel): exactly the pairs that were blended;
//   * phase 1, thread = PIXEL: every lane pops ITS hit bits back to front at its own pace (no alpha test: a hit is a
//     contribution), runs the recurrences and appends (u, w) to the pixel's slice of a pair buffer;
//   * the hit words are transposed per warp (the forward's shuffle butterfly) into one pixel mask per (entry, 8x4
//     region);
//   * phase 2, thread = ENTRY: every lane walks the pixels of ITS splat, looks the pair up (slot = the pixel's slice
//     start + number of the pixel's hits above the entry: one popcount), forms the nine sums in registers and issues
//     THREE vector reductions (RED.ADD.F32x4) per splat and tile -- 11 k reductions per tile instead of 40 k;
//   * the pair buffer holds 4096 pairs; a batch with more is processed in chunks of whole half-windows (16 entries x
//     256 pixels always fit).
#include "common.cuh"
#include "stage.cuh"

namespace gsb {

constexpr int BWD_CAP = 4096;                 // pairs between the phases
constexpr int BWD_WINS = BLEND_THREADS / 32;  // windows per batch

// 32 x 32 bit-matrix transpose across a warp (see blend_fwd.cu)
struct BwdTranspose {
    uint32_t sel16, sel8, m4, m2, m1, r4, r2, r1;
    __device__ __forceinline__ BwdTranspose(uint32_t lane)
    {
        sel16 = (lane & 16) ? 0x3276u : 0x5410u;
        sel8 = (lane & 8) ? 0x3715u : 0x6240u;
        m4 = (lane & 4) ? 0xf0f0f0f0u : 0x0f0f0f0fu; r4 = (lane & 4) ? 28u : 4u;
        m2 = (lane & 2) ? 0xccccccccu : 0x33333333u; r2 = (lane & 2) ? 30u : 2u;
        m1 = (lane & 1) ? 0xaaaaaaaau : 0x55555555u; r1 = (lane & 1) ? 31u : 1u;
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

struct BwdSmem {
    float4 a[BLEND_THREADS], b[BLEND_THREADS], c[BLEND_THREADS];   // the batch's records
    uint32_t ids[BLEND_THREADS];
    uint32_t hits[BWD_WINS][BLEND_THREADS];      // [window][pixel]: entries of the window blended into the pixel
    uint32_t hitsT[BWD_WINS][BLEND_THREADS];     // [window][region * 32 + entry]: pixels of the region the entry was blended into
    uint16_t base[BWD_WINS][BLEND_THREADS];      // [window][pixel]: pair slot of the pixel's highest hit of the window (this chunk)
    float2 pairs[BWD_CAP];                       // (u, w) of the chunk's pairs, pixel after pixel, back to front inside a pixel
    float4 pixd[BLEND_THREADS];                  // dL/dpixel of the tile's pixels: r, g, b, z (five-channel pass)
    float2 pixxy[BLEND_THREADS];                 // pixel centres
    uint32_t half_total[2 * BWD_WINS];           // pairs per half window (16 entries) of the batch
    uint32_t warp_sum[BLEND_THREADS / 32];
};

__device__ __forceinline__ void red_add_v4(float* dst, float x, float y, float z, float w)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

// CH = 5: backward of the fused RGB + depth / silhouette pass (see blend_fwd.cu): dL/dalpha sums over five channels,
// and the gradient of the z_cam colour (sum of w * dL/dpix[3]) lands in accumulator slot 9.
template <int MINB, int CH>
__global__ void __launch_bounds__(BLEND_THREADS, MINB)
blend_backward_twophase_kernel(const uint2* __restrict__ ranges, const char* __restrict__ binning,
                      const SplatRec* __restrict__ rec, int W, int H, const float* __restrict__ bg,
                      const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
                      const uint32_t* __restrict__ tile_max_contrib, const float* __restrict__ dL_dpix,
                      const float* __restrict__ dL_ddepth_sil, float* __restrict__ acc /* [P][16] */, const uint32_t* __restrict__ hits_tail,
                      const GeomHeader* __restrict__ hdr, uint32_t band_y0)
{
    extern __shared__ __align__(16) unsigned char bwd_smem_raw[];
    BwdSmem& S = *reinterpret_cast<BwdSmem*>(bwd_smem_raw);
    const uint32_t tile_y = band_y0 + blockIdx.y;   // the grid covers the band's tile rows
    const uint32_t tile = tile_y * gridDim.x + blockIdx.x;
    const uint2 range = ranges[tile];
    const uint32_t len = range.y - range.x;
    // entries [0, n) can matter: n = highest n_contrib over the tile's pixels (recorded by the forward)
    const int n = min((int)len, (int)tile_max_contrib[2 * tile]);
    const int batches = (n + BLEND_THREADS - 1) / BLEND_THREADS;
    if (batches == 0) return;
    const BinningLayout BL = BinningLayout::make((long long)hdr->layout_capacity);
    const uint32_t* point_list = reinterpret_cast<const uint32_t*>(binning + BL.point_list);
    uint32_t* hits_full = const_cast<uint32_t*>(reinterpret_cast<const uint32_t*>(binning + BL.hits));
    // thread = pixel (phase 1): warp -> 8x4 region, lane -> row-major pixel of the region, as in the forward;
    // thread = entry (phase 2): warp -> window of the batch, lane -> entry of the window
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int px = blockIdx.x * TILE_X + (warp & 1) * 8 + (int)(lane & 7), py = tile_y * TILE_Y + (warp >> 1) * 4 + (int)(lane >> 3);
    const bool inside = px < W && py < H;
    const size_t HW = (size_t)W * H, pix = (size_t)py * W + px;

    const float T_final = inside ? final_T[pix] : 0.f;
    float T = T_final;
    const int last_contributor = inside ? (int)n_contrib[pix] : 0;
    const int my_last_window = (last_contributor - 1) >> 5;   // -1: this pixel blended nothing
    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f, d4 = 0.f;
    if (inside) {
        d0 = dL_dpix[pix];
        d1 = dL_dpix[HW + pix];
        d2 = dL_dpix[2 * HW + pix];
        if (CH == 5) {
            d3 = dL_ddepth_sil[pix];
            d4 = dL_ddepth_sil[HW + pix];
        }
    }
    float bg_dot_dpixel = __ldg(bg) * d0 + __ldg(bg + 1) * d1 + __ldg(bg + 2) * d2;
    if (CH == 5) bg_dot_dpixel += __ldg(bg) * d3 + __ldg(bg + 1) * d4;   // the depth pass blends over the same background tensor
    S.pixd[tid] = make_float4(d0, d1, d2, d3);
    S.pixxy[tid] = make_float2((float)px, (float)py);
    const float pxf = (float)px, pyf = (float)py;
    float ar0 = 0.f, ar1 = 0.f, ar2 = 0.f, ar3 = 0.f, ar4 = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, lc3 = 0.f, lc4 = 0.f, last_alpha = 0.f;
    const BwdTranspose transpose(lane);
    const uint32_t* ids = point_list + range.x;

    for (int kb = batches - 1; kb >= 0; kb--) {
        const int cnt = min(BLEND_THREADS, n - kb * BLEND_THREADS);
        const int nwin = (cnt + 31) >> 5;
        // ---- stage the batch: one record, one id and one column of hit words per thread ----
        {
            const int e = kb * BLEND_THREADS + (int)tid;
            const uint32_t id = e < n ? __ldg(ids + e) : 0xffffffffu;
            S.ids[tid] = id;
            if (id != 0xffffffffu) {
                const SplatRec* r = rec + id;
                cp_async16(&S.a[tid], &r->a);
                cp_async16(&S.b[tid], &r->b);
                cp_async16(&S.c[tid], &r->c);
            }
            if (tid < 2 * BWD_WINS) S.half_total[tid] = 0;
#pragma unroll
            for (int w = 0; w < BWD_WINS; w++) {
                const int gw = kb * BWD_WINS + w;
                // words of windows past the one this pixel stopped in were never written by the forward (its warp had retired)
                if (w < nwin && gw <= my_last_window)
                    cp_async4(&S.hits[w][tid], hit_words(hits_full, const_cast<uint32_t*>(hits_tail), tile, range.x, len, (uint32_t)gw) + tid);
                else
                    S.hits[w][tid] = 0;
            }
            cp_async_commit();
            cp_async_wait<0>();
        }
        __syncthreads();
        // ---- per-entry pixel masks (transpose of this warp's region) and the pair count of every half window ----
#pragma unroll
        for (int w = 0; w < BWD_WINS; w++) {
            if (w < nwin) {
                const uint32_t x = S.hits[w][tid];
                S.hitsT[w][tid] = transpose(x);
                const uint32_t lo = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(x & 0xffffu));
                const uint32_t hi = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(x >> 16));
                if (lane == 0) {
                    if (lo) atomicAdd(&S.half_total[2 * w], lo);
                    if (hi) atomicAdd(&S.half_total[2 * w + 1], hi);
                }
            }
        }
        __syncthreads();
        // ---- chunks of whole half windows, back to front, each at most BWD_CAP pairs ----
        int hw_hi = 2 * nwin - 1;
        while (hw_hi >= 0) {
            int hw_lo = hw_hi;
            uint32_t chunk_pairs = S.half_total[hw_hi];
            while (hw_lo > 0 && chunk_pairs + S.half_total[hw_lo - 1] <= (uint32_t)BWD_CAP) chunk_pairs += S.half_total[--hw_lo];
            if (chunk_pairs == 0) {   // CTA-uniform
                hw_hi = hw_lo - 1;
                continue;
            }
            const int w_hi = hw_hi >> 1, w_lo = hw_lo >> 1;
            const uint32_t top_mask = (hw_hi & 1) ? 0xffffffffu : 0x0000ffffu;   // bits of window w_hi inside the chunk
            const uint32_t bot_mask = (hw_lo & 1) ? 0xffff0000u : 0xffffffffu;   // bits of window w_lo inside the chunk
            auto chunk_bits = [&](int w, uint32_t x) -> uint32_t {
                if (w == w_hi) x &= top_mask;
                if (w == w_lo) x &= bot_mask;
                return x;
            };
            // -- slice of the pair buffer of every pixel, and of every (pixel, window) inside it --
            uint32_t mine = 0, nzw = 0;   // pairs of this pixel in the chunk; windows in which it has any
            for (int w = w_hi; w >= w_lo; w--) {
                const uint32_t c = __popc(chunk_bits(w, S.hits[w][tid]));
                mine += c;
                nzw |= (c ? 1u : 0u) << w;
            }
            uint32_t incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= (uint32_t)o) incl += up;
            }
            if (lane == 31) S.warp_sum[warp] = incl;
            __syncthreads();   // warp sums visible (the previous chunk's readers of `pairs` / `base` / `warp_sum` passed the barrier that ended it)
            uint32_t slot = incl - mine;
            for (uint32_t w = 0; w < warp; w++) slot += S.warp_sum[w];
            {
                uint32_t run = slot;
                for (int w = w_hi; w >= w_lo; w--) {
                    S.base[w][tid] = (uint16_t)run;
                    run += __popc(chunk_bits(w, S.hits[w][tid]));
                }
            }
            // -- phase 1: thread = pixel.  Pop the pixel's hits back to front, run the recurrences, append (u, w) --
            {
                int wcur = 0;
                uint32_t mask = 0;
                while (__any_sync(0xffffffffu, (mask | nzw) != 0)) {
                    if (mask == 0 && nzw != 0) {   // next (lower) window in which this pixel blended anything
                        wcur = 31 - __clz(nzw);
                        nzw ^= 1u << wcur;
                        mask = chunk_bits(wcur, S.hits[wcur][tid]);
                    }
                    if (mask != 0) {
                        const int f = 31 - __clz(mask);
                        mask ^= 1u << f;
                        const int e = wcur * 32 + f;
                        const float4 A = S.a[e];
                        const float4 B = S.b[e];
                        const float4 Cc = S.c[e];
                        const float dx = __fsub_rn(A.x, pxf), dy = __fsub_rn(A.y, pyf);
                        const float power = splat_power(dx, dy, B.x, B.y, B.z);
                        const float G = expf(fminf(power, 0.0f));
                        const float alpha = fminf(0.99f, __fmul_rn(B.w, G));
                        const float inv = rcp_approx(1.0f - alpha);                 // 1 / (1 - alpha), one MUFU.RCP
                        const float Tn = T * inv;                                   // T_before = T_after / (1 - alpha)
                        const float om = 1.f - last_alpha;
                        const float a0 = fmaf(last_alpha, lc0, om * ar0);           // colour accumulated behind this splat
                        const float a1 = fmaf(last_alpha, lc1, om * ar1);
                        const float a2 = fmaf(last_alpha, lc2, om * ar2);
                        float chan = fmaf(Cc.x - a0, d0, fmaf(Cc.y - a1, d1, (Cc.z - a2) * d2));
                        if (CH == 5) {
                            const float a3 = fmaf(last_alpha, lc3, om * ar3);
                            const float a4 = fmaf(last_alpha, lc4, om * ar4);
                            chan = fmaf(Cc.w - a3, d3, fmaf(1.0f - a4, d4, chan));
                            ar3 = a3; ar4 = a4;
                            lc3 = Cc.w; lc4 = 1.0f;
                        }
                        float dL_dalpha = Tn * chan;
                        dL_dalpha = fmaf(-T_final * inv, bg_dot_dpixel, dL_dalpha);
                        S.pairs[slot++] = make_float2(G * dL_dalpha, alpha * Tn);   // u, w
                        T = Tn;
                        ar0 = a0; ar1 = a1; ar2 = a2;
                        lc0 = Cc.x; lc1 = Cc.y; lc2 = Cc.z;
                        last_alpha = alpha;
                    }
                }
            }
            __syncthreads();
            // -- phase 2: thread = entry.  Walk the pixels the splat was blended into, sum, reduce into the accumulator --
            {
                const int hw = (int)(tid >> 4);   // half window of this thread's entry
                const bool active = hw >= hw_lo && hw <= hw_hi && (int)tid < cnt;
                const float4 A = S.a[tid];
                float m_dx = 0.f, m_dy = 0.f, m_dx2 = 0.f, m_dxdy = 0.f, m_dy2 = 0.f, m_u = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
                uint32_t nzr = 0;   // regions holding pixels of this entry
                if (active) {
#pragma unroll
                    for (int rr = 0; rr < 8; rr++) nzr |= (S.hitsT[warp][rr * 32 + lane] ? 1u : 0u) << rr;
                }
                int r = 0;
                uint32_t mask = 0;
                const bool any = nzr != 0;
                while (__any_sync(0xffffffffu, (mask | nzr) != 0)) {
                    if (mask == 0 && nzr != 0) {
                        r = __ffs(nzr) - 1;
                        nzr &= nzr - 1;
                        mask = S.hitsT[warp][r * 32 + lane];
                    }
                    if (mask != 0) {
                        const int p = __ffs(mask) - 1;
                        mask &= mask - 1;
                        const int P = r * 32 + p;
                        const uint32_t above = (chunk_bits((int)warp, S.hits[warp][P]) >> lane) >> 1;   // the pixel's hits above this entry
                        const float2 uw = S.pairs[(uint32_t)S.base[warp][P] + (uint32_t)__popc(above)];
                        const float4 d = S.pixd[P];
                        const float2 xy = S.pixxy[P];
                        const float dx = __fsub_rn(A.x, xy.x), dy = __fsub_rn(A.y, xy.y);
                        const float ux = uw.x * dx, uy = uw.x * dy;
                        m_dx += ux; m_dy += uy;
                        m_dx2 = fmaf(ux, dx, m_dx2); m_dxdy = fmaf(ux, dy, m_dxdy); m_dy2 = fmaf(uy, dy, m_dy2);
                        m_u += uw.x;
                        c0 = fmaf(uw.y, d.x, c0); c1 = fmaf(uw.y, d.y, c1); c2 = fmaf(uw.y, d.z, c2);
                        if (CH == 5) c3 = fmaf(uw.y, d.w, c3);
                    }
                }
                if (any) {
                    float* dst = acc + (size_t)S.ids[tid] * ACC_FLOATS;   // GradAcc layout (common.cuh)
                    red_add_v4(dst, m_dx, m_dx2, c2, 0.f);
                    red_add_v4(dst + 4, m_dxdy, c0, 0.f, 0.f);
                    red_add_v4(dst + 8, m_dy, m_dy2, CH == 5 ? c3 : 0.f, 0.f);
                    red_add_v4(dst + 12, m_u, c1, 0.f, 0.f);
                }
            }
            hw_hi = hw_lo - 1;
            if (hw_hi >= 0) __syncthreads();   // the next chunk rewrites `warp_sum`, `base` and `pairs`
        }
        __syncthreads();   // the next batch rewrites the records and the hit words
    }
}

template <int CH>
static void launch_bwd2(const FwdParams& p, char* geom, const GeomLayout& GL, const char* binning, const char* image,
                       const ImageLayout& IL, const float* dL_dpix, const float* dL_ddepth_sil, cudaStream_t s)
{
    constexpr int MINB = 3;
    GSB_SET_ATTR_ONCE((blend_backward_twophase_kernel<MINB, CH>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdSmem));
    blend_backward_twophase_kernel<MINB, CH><<<dim3(IL.tiles_x, p.band_y1 - p.band_y0), BLEND_THREADS, sizeof(BwdSmem), s>>>(
        reinterpret_cast<const uint2*>(image + IL.ranges), binning, reinterpret_cast<const SplatRec*>(geom + GL.rec), p.W, p.H,
        p.background, reinterpret_cast<const float*>(image + IL.final_T), reinterpret_cast<const uint32_t*>(image + IL.n_contrib),
        reinterpret_cast<const uint32_t*>(image + IL.tile_max_contrib), dL_dpix, dL_ddepth_sil,
        reinterpret_cast<float*>(geom + GL.acc), reinterpret_cast<const uint32_t*>(image + IL.hits_tail),
        reinterpret_cast<const GeomHeader*>(geom + GL.header), (uint32_t)p.band_y0);
}

int launch_blend_backward_twophase(const FwdParams& p, char* geom, const GeomLayout& GL, const char* binning,
                          const char* image, const ImageLayout& IL, const float* dL_dpix, const float* dL_ddepth_sil,
                          cudaStream_t s)
{
    if (p.W <= 0 || p.H <= 0 || p.P <= 0) return GSB_OK;
    GSB_CUDA_CHECK(cudaMemsetAsync(geom + GL.acc, 0, (size_t)p.P * sizeof(GradAcc), s));
    if (p.band_y1 <= p.band_y0) return GSB_OK;   // empty tile-row band: all-zero accumulators
    {
        StageTimer _t(ST_BLEND_BWD, s);
        if (dL_ddepth_sil) launch_bwd2<5>(p, geom, GL, binning, image, IL, dL_dpix, dL_ddepth_sil, s);
        else launch_bwd2<3>(p, geom, GL, binning, image, IL, dL_dpix, dL_ddepth_sil, s);
        GSB_LAUNCH_CHECK();
    }
    return GSB_OK;
}

}  // namespace gsb
