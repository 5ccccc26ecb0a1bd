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
//   * the walk starts at the tile's highest n_contrib (recorded by the forward pass), not at
//     the end of the tile's list, and each warp culls staged splats against its four pixel
//     blocks and their highest n_contrib with four ballots per 32 splats;
//   * records are gathered with cp.async into double-buffered shared memory (stage.cuh).
#include <cstdlib>
#include "common.cuh"
#include "stage.cuh"

namespace gsb {

template <int MINB>
__global__ void __launch_bounds__(BLEND_THREADS, MINB)
blend_backward_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                      const SplatRec* __restrict__ rec, int W, int H, const float* __restrict__ bg,
                      const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
                      const uint32_t* __restrict__ tile_max_contrib, const float* __restrict__ dL_dpix,
                      float* __restrict__ acc /* [P][12] */)
{
    __shared__ StageBuf S;
    __shared__ uint32_t s_ids[2][BLEND_BATCH];
    const uint32_t tile = blockIdx.y * gridDim.x + blockIdx.x;
    const uint2 range = ranges[tile];
    const int n = min((int)(range.y - range.x), (int)tile_max_contrib[tile]);  // entries [0, n) can matter
    const int batches = (n + BLEND_BATCH - 1) / BLEND_BATCH;
    if (batches == 0) return;
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t q = lane >> 3, l8 = lane & 7, qshift = q * 8;
    const int bx0 = blockIdx.x * TILE_X + (warp & 1) * 8, by0 = blockIdx.y * TILE_Y + (warp >> 1) * 4;
    const int px = bx0 + (q & 1) * 4 + (l8 & 3), py = by0 + (q >> 1) * 2 + (l8 >> 2);
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    const float xa0 = (float)bx0, xa1 = (float)(bx0 + 3), xb0 = (float)(bx0 + 4), xb1 = (float)(bx0 + 7);
    const float ya0 = (float)by0, ya1 = (float)(by0 + 1), yb0 = (float)(by0 + 2), yb1 = (float)(by0 + 3);
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
    // highest n_contrib of each quarter-warp (the cull pass needs all four) and of the warp
    int qmax = last_contributor;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) qmax = max(qmax, __shfl_xor_sync(0xffffffffu, qmax, o));
    const int qm0 = __shfl_sync(0xffffffffu, qmax, 0), qm1 = __shfl_sync(0xffffffffu, qmax, 8);
    const int qm2 = __shfl_sync(0xffffffffu, qmax, 16), qm3 = __shfl_sync(0xffffffffu, qmax, 24);
    const int warp_max = max(max(qm0, qm1), max(qm2, qm3));

    // slot t of batch k holds list entry (n - 1 - k*256 - t): slots run back to front
    const uint32_t* ids = point_list + range.x;
    auto load_id = [&](int k) -> uint32_t {
        const int i = n - 1 - k * BLEND_BATCH - (int)threadIdx.x;
        return i >= 0 ? __ldg(ids + i) : 0xffffffffu;
    };
    uint32_t id_cur = load_id(0);
    s_ids[0][threadIdx.x] = id_cur;
    stage_issue(S, 0, rec, id_cur);
    uint32_t id_next = batches > 1 ? load_id(1) : 0xffffffffu;

    for (int b = 0; b < batches; b++) {
        const int buf = b & 1;
        if (b + 1 < batches) {
            s_ids[buf ^ 1][threadIdx.x] = id_next;
            stage_issue(S, buf ^ 1, rec, id_next);
        } else {
            cp_async_commit();
        }
        if (b + 2 < batches) id_next = load_id(b + 2);
        cp_async_wait<1>();
        __syncthreads();
        const int first_pos = n - b * BLEND_BATCH;  // 1-based list position of slot 0
        const int cnt = min(BLEND_BATCH, first_pos);
        if (first_pos - cnt < warp_max) {  // some slot of this batch has pos <= warp_max
            for (int c0 = 0; c0 < cnt; c0 += 32) {
                // ---- cull pass ----
                const int j = c0 + (int)lane;
                const int posj = first_pos - j;
                bool hxa = false, hxb = false, hya = false, hyb = false;
                if (j < cnt && posj <= warp_max) {
                    const float4 A = S.a[buf][j];
                    const float2 e = __half22float2(*reinterpret_cast<const __half2*>(&A.z));
                    const float lox = A.x - e.x, hix = A.x + e.x, loy = A.y - e.y, hiy = A.y + e.y;
                    hxa = !(hix < xa0 || lox > xa1);
                    hxb = !(hix < xb0 || lox > xb1);
                    hya = !(hiy < ya0 || loy > ya1);
                    hyb = !(hiy < yb0 || loy > yb1);
                }
                const uint32_t m0 = __ballot_sync(0xffffffffu, hxa && hya && posj <= qm0);
                const uint32_t m1 = __ballot_sync(0xffffffffu, hxb && hya && posj <= qm1);
                const uint32_t m2 = __ballot_sync(0xffffffffu, hxa && hyb && posj <= qm2);
                const uint32_t m3 = __ballot_sync(0xffffffffu, hxb && hyb && posj <= qm3);
                uint32_t mask = q == 0 ? m0 : q == 1 ? m1 : q == 2 ? m2 : m3;
                // ---- gradient pass: every quarter-warp walks its survivors back to front ----
                while (__any_sync(0xffffffffu, mask != 0)) {
                    const bool act = mask != 0;
                    const int e = c0 + (act ? __ffs(mask) - 1 : 0);
                    mask &= mask - 1;
                    const int pos = first_pos - e;
                    const float4 A = S.a[buf][e];
                    const float4 B = S.b[buf][e];
                    const float dx = __fsub_rn(A.x, pxf), dy = __fsub_rn(A.y, pyf);
                    const float power = splat_power(dx, dy, B.x, B.y, B.z);
                    float v[8], v8 = 0.f;
#pragma unroll
                    for (int k = 0; k < 8; k++) v[k] = 0.f;
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
                        for (int k = 0; k < 4; k++) {
                            const float send = hi ? v[k] : v[k + 4];
                            const float keep = hi ? v[k + 4] : v[k];
                            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                        }
                    }
                    {
                        const bool hi = l8 & 2;
#pragma unroll
                        for (int k = 0; k < 2; k++) {
                            const float send = hi ? v[k] : v[k + 2];
                            const float keep = hi ? v[k + 2] : v[k];
                            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
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
                        const int k = ((l8 >> 2) & 1) * 4 + ((l8 >> 1) & 1) * 2 + (l8 & 1);
                        atomicAdd(dst + k, v[0]);
                        if (l8 == 0) atomicAdd(dst + 8, v8);
                    }
                }
            }
        }
        __syncthreads();
    }
    cp_async_wait<0>();
}

int launch_blend_backward(const FwdParams& p, char* geom, const GeomLayout& GL, const uint32_t* point_list,
                          const char* image, const ImageLayout& IL, const float* dL_dpix, cudaStream_t s)
{
    if (p.W <= 0 || p.H <= 0 || p.P <= 0) return GSB_OK;
    GSB_CUDA_CHECK(cudaMemsetAsync(geom + GL.acc, 0, (size_t)p.P * sizeof(GradAcc), s));
    // tuning knob (resident CTAs per SM the compiler must allow, i.e. the register budget)
    static const int minb = [] { const char* e = getenv("GSB_BLEND_BWD_MINB"); return e ? atoi(e) : 4; }();
    dim3 grid(IL.tiles_x, IL.tiles_y);
    {
        StageTimer _t(ST_BLEND_BWD, s);
        switch (minb) {
            case 3: blend_backward_kernel<3><<<grid, BLEND_THREADS, 0, s>>>(reinterpret_cast<const uint2*>(image + IL.ranges), point_list, reinterpret_cast<const SplatRec*>(geom + GL.rec),
            p.W, p.H, p.background, reinterpret_cast<const float*>(image + IL.final_T),
            reinterpret_cast<const uint32_t*>(image + IL.n_contrib), reinterpret_cast<const uint32_t*>(image + IL.tile_max_contrib),
            dL_dpix, reinterpret_cast<float*>(geom + GL.acc)); break;
            case 5: blend_backward_kernel<5><<<grid, BLEND_THREADS, 0, s>>>(reinterpret_cast<const uint2*>(image + IL.ranges), point_list, reinterpret_cast<const SplatRec*>(geom + GL.rec),
            p.W, p.H, p.background, reinterpret_cast<const float*>(image + IL.final_T),
            reinterpret_cast<const uint32_t*>(image + IL.n_contrib), reinterpret_cast<const uint32_t*>(image + IL.tile_max_contrib),
            dL_dpix, reinterpret_cast<float*>(geom + GL.acc)); break;
            case 6: blend_backward_kernel<6><<<grid, BLEND_THREADS, 0, s>>>(reinterpret_cast<const uint2*>(image + IL.ranges), point_list, reinterpret_cast<const SplatRec*>(geom + GL.rec),
            p.W, p.H, p.background, reinterpret_cast<const float*>(image + IL.final_T),
            reinterpret_cast<const uint32_t*>(image + IL.n_contrib), reinterpret_cast<const uint32_t*>(image + IL.tile_max_contrib),
            dL_dpix, reinterpret_cast<float*>(geom + GL.acc)); break;
            default: blend_backward_kernel<4><<<grid, BLEND_THREADS, 0, s>>>(reinterpret_cast<const uint2*>(image + IL.ranges), point_list, reinterpret_cast<const SplatRec*>(geom + GL.rec),
            p.W, p.H, p.background, reinterpret_cast<const float*>(image + IL.final_T),
            reinterpret_cast<const uint32_t*>(image + IL.n_contrib), reinterpret_cast<const uint32_t*>(image + IL.tile_max_contrib),
            dL_dpix, reinterpret_cast<float*>(geom + GL.acc)); break;
        }
        GSB_LAUNCH_CHECK();
    }
    return GSB_OK;
}

}  // namespace gsb
