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
//   * splat records are one 48-byte gather (3 x 16-byte cp.async straight into shared
//     memory, double buffered, ids prefetched two batches ahead) instead of four separate
//     arrays plus a per-pair colour read from global memory;
//   * the tile's highest n_contrib is recorded for the backward pass.
#include <cstdlib>
#include "common.cuh"
#include "stage.cuh"

namespace gsb {

template <int MINB>
__global__ void __launch_bounds__(BLEND_THREADS, MINB)
blend_forward_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                     const SplatRec* __restrict__ rec, int W, int H, const float* __restrict__ bg,
                     float* __restrict__ out_color, float* __restrict__ out_depth, float* __restrict__ final_T,
                     uint32_t* __restrict__ n_contrib, uint32_t* __restrict__ tile_max_contrib)
{
    __shared__ StageBuf S;
    __shared__ uint32_t s_max;
    const uint32_t tile = blockIdx.y * gridDim.x + blockIdx.x;
    const uint2 range = ranges[tile];
    const int n = (int)(range.y - range.x);
    const int batches = (n + BLEND_BATCH - 1) / BLEND_BATCH;
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    // warp -> 8x4 pixel block of the tile; quarter-warp q -> 4x2 sub-block; lane -> pixel
    const uint32_t q = lane >> 3, l8 = lane & 7, qshift = q * 8;
    const int bx0 = blockIdx.x * TILE_X + (warp & 1) * 8, by0 = blockIdx.y * TILE_Y + (warp >> 1) * 4;
    const int px = bx0 + (q & 1) * 4 + (l8 & 3), py = by0 + (q >> 1) * 2 + (l8 >> 2);
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    // sub-block extents used by the cull pass (pixel centres): x halves [bx0, bx0+3], [bx0+4, bx0+7];
    // y halves [by0, by0+1], [by0+2, by0+3]
    const float xa0 = (float)bx0, xa1 = (float)(bx0 + 3), xb0 = (float)(bx0 + 4), xb1 = (float)(bx0 + 7);
    const float ya0 = (float)by0, ya1 = (float)(by0 + 1), yb0 = (float)(by0 + 2), yb1 = (float)(by0 + 3);
    if (threadIdx.x == 0) s_max = 0;

    bool done = !inside;
    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, D = 0.f;
    uint32_t last = 0;

    const uint32_t* ids = point_list + range.x;
    uint32_t id_next = 0;
    if (batches > 0) {
        const uint32_t id0 = (int)threadIdx.x < n ? __ldg(ids + threadIdx.x) : 0xffffffffu;
        stage_issue(S, 0, rec, id0);
        if (batches > 1) id_next = BLEND_BATCH + (int)threadIdx.x < n ? __ldg(ids + BLEND_BATCH + threadIdx.x) : 0xffffffffu;
    }
    uint32_t done_bits = __ballot_sync(0xffffffffu, done);
    bool warp_done = done_bits == 0xffffffffu;
    for (int b = 0; b < batches; b++) {
        const int buf = b & 1;
        if (b + 1 < batches) stage_issue(S, buf ^ 1, rec, id_next);
        else cp_async_commit();
        if (b + 2 < batches) {
            const int e = (b + 2) * BLEND_BATCH + (int)threadIdx.x;
            id_next = e < n ? __ldg(ids + e) : 0xffffffffu;
        }
        cp_async_wait<1>();
        if (__syncthreads_count(warp_done) == BLEND_THREADS) break;  // every pixel of the tile is saturated
        if (!warp_done) {
            const int cnt = min(BLEND_BATCH, n - b * BLEND_BATCH);
            for (int c0 = 0; c0 < cnt; c0 += 32) {
                // ---- cull pass: lane j tests staged splat c0+j against the four 4x2 sub-blocks ----
                const int j = c0 + (int)lane;
                bool hxa = false, hxb = false, hya = false, hyb = false;
                if (j < cnt) {
                    const float4 A = S.a[buf][j];
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
                // ---- blend pass: every quarter-warp walks ITS survivors, in list order ----
                while (__any_sync(0xffffffffu, mask != 0)) {
                    const bool act = mask != 0;
                    const int e = c0 + (act ? __ffs(mask) - 1 : 0);
                    mask &= mask - 1;
                    const float4 A = S.a[buf][e];
                    const float4 B = S.b[buf][e];
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
                    const float4 Cc = S.c[buf][e];
                    C0 = fmaf(__fmul_rn(Cc.x, alpha), T, C0);
                    C1 = fmaf(__fmul_rn(Cc.y, alpha), T, C1);
                    C2 = fmaf(__fmul_rn(Cc.z, alpha), T, C2);
                    if (T > 0.5f) D = Cc.w;
                    T = test_T;
                    last = (uint32_t)(b * BLEND_BATCH + e + 1);
                }
                done_bits = __ballot_sync(0xffffffffu, done);
                if (done_bits == 0xffffffffu) {
                    warp_done = true;
                    break;
                }
            }
        }
        __syncthreads();  // everyone is finished with `buf` before batch b+2 is staged into it
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
    }
    uint32_t m = last;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0 && m) atomicMax(&s_max, m);
    __syncthreads();
    if (threadIdx.x == 0) tile_max_contrib[tile] = s_max;
}

int launch_blend_forward(const FwdParams& p, const char* geom, const GeomLayout& GL, const uint32_t* point_list,
                         char* image, const ImageLayout& IL, float* out_color, float* out_depth, cudaStream_t s)
{
    if (p.W <= 0 || p.H <= 0) return GSB_OK;
    // tuning knob (resident CTAs per SM the compiler must allow, i.e. the register budget)
    static const int minb = [] { const char* e = getenv("GSB_BLEND_FWD_MINB"); return e ? atoi(e) : 6; }();
    dim3 grid(IL.tiles_x, IL.tiles_y);
    {
        StageTimer _t(ST_BLEND_FWD, s);
        switch (minb) {
            case 4: blend_forward_kernel<4><<<grid, BLEND_THREADS, 0, s>>>(reinterpret_cast<const uint2*>(image + IL.ranges), point_list, reinterpret_cast<const SplatRec*>(geom + GL.rec),
            p.W, p.H, p.background, out_color, out_depth, reinterpret_cast<float*>(image + IL.final_T),
            reinterpret_cast<uint32_t*>(image + IL.n_contrib), reinterpret_cast<uint32_t*>(image + IL.tile_max_contrib)); break;
            case 8: blend_forward_kernel<8><<<grid, BLEND_THREADS, 0, s>>>(reinterpret_cast<const uint2*>(image + IL.ranges), point_list, reinterpret_cast<const SplatRec*>(geom + GL.rec),
            p.W, p.H, p.background, out_color, out_depth, reinterpret_cast<float*>(image + IL.final_T),
            reinterpret_cast<uint32_t*>(image + IL.n_contrib), reinterpret_cast<uint32_t*>(image + IL.tile_max_contrib)); break;
            default: blend_forward_kernel<6><<<grid, BLEND_THREADS, 0, s>>>(reinterpret_cast<const uint2*>(image + IL.ranges), point_list, reinterpret_cast<const SplatRec*>(geom + GL.rec),
            p.W, p.H, p.background, out_color, out_depth, reinterpret_cast<float*>(image + IL.final_T),
            reinterpret_cast<uint32_t*>(image + IL.n_contrib), reinterpret_cast<uint32_t*>(image + IL.tile_max_contrib)); break;
        }
        GSB_LAUNCH_CHECK();
    }
    return GSB_OK;
}

}  // namespace gsb
