// blend_bwd.cu -- per-tile back-to-front gradient of the alpha blend (K7) for sm_100a.
//
// Replaces BACKWARD::render / renderCUDA<3> (backward.cu:399-557, launch :641-656).
// Same recurrences as the reference: start from the stored final transmittance and the
// position of the last blended splat, walk the tile's list backwards, T <- T / (1 - alpha),
// running "colour behind" accumulator, background term, gradients w.r.t. colour, 2D mean
// (in NDC units: x 0.5 W, x 0.5 H), conic (slots x, y, w) and opacity.  No depth gradient.
//
// What is different (design, not results):
//   * the reference issues 9 global atomicAdds per contributing (pixel, splat) pair; here a
//     warp owns an 8x4 pixel block, reduces the 9 partial sums of a splat across its lanes
//     with a shuffle reduce-scatter (14 SHFL instead of 45) and issues ONE predicated
//     RED.ADD.F32 instruction (9 lanes, one 48-byte packed accumulator per Gaussian), and only
//     for splats that touched at least one pixel of the block;
//   * the walk starts at the tile's highest n_contrib (recorded by the forward pass), not at
//     the end of the tile's list, and each warp culls staged splats against its pixel block
//     and its own highest n_contrib with one ballot per 32 splats;
//   * records are gathered with cp.async into double-buffered shared memory (stage.cuh).
#include "common.cuh"
#include "stage.cuh"

namespace gsb {

__global__ void __launch_bounds__(BLEND_THREADS)
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
    const int bx0 = blockIdx.x * TILE_X + (warp & 1) * 8, by0 = blockIdx.y * TILE_Y + (warp >> 1) * 4;
    const int px = bx0 + (lane & 7), py = by0 + (lane >> 3);
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    const float fx0 = (float)bx0, fx1 = (float)(bx0 + 7), fy0 = (float)by0, fy1 = (float)(by0 + 3);
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
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
    int warp_max = last_contributor;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) warp_max = max(warp_max, __shfl_xor_sync(0xffffffffu, warp_max, o));

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
                const int j = c0 + (int)lane;
                bool hit = false;
                if (j < cnt && first_pos - j <= warp_max) {
                    const float2 c = *reinterpret_cast<const float2*>(&S.a[buf][j]);
                    const float2 e = __half22float2(*reinterpret_cast<const __half2*>(&S.c[buf][j].w));
                    hit = !(c.x + e.x < fx0 || c.x - e.x > fx1 || c.y + e.y < fy0 || c.y - e.y > fy1);
                }
                uint32_t mask = __ballot_sync(0xffffffffu, hit);
                while (mask) {
                    const int jj = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const int e = c0 + jj;
                    const int pos = first_pos - e;
                    const float4 A = S.a[buf][e];
                    const float4 B = S.b[buf][e];
                    const float dx = __fsub_rn(A.x, pxf), dy = __fsub_rn(A.y, pyf);
                    const float power = splat_power(dx, dy, A.z, A.w, B.x);
                    float v[8], v8 = 0.f;
#pragma unroll
                    for (int k = 0; k < 8; k++) v[k] = 0.f;
                    bool contrib = false;
                    if (pos <= last_contributor && !(power > 0.0f) && !(power < B.z)) {
                        const float G = expf(power);
                        const float alpha = fminf(0.99f, __fmul_rn(B.y, G));
                        if (!(alpha < 1.0f / 255.0f)) {
                            contrib = true;
                            const float4 Cc = S.c[buf][e];
                            T = __fdiv_rn(T, __fsub_rn(1.f, alpha));
                            const float dchannel_dcolor = alpha * T;
                            ar0 = last_alpha * lc0 + (1.f - last_alpha) * ar0;
                            ar1 = last_alpha * lc1 + (1.f - last_alpha) * ar1;
                            ar2 = last_alpha * lc2 + (1.f - last_alpha) * ar2;
                            lc0 = Cc.x; lc1 = Cc.y; lc2 = Cc.z;
                            float dL_dalpha = (Cc.x - ar0) * d0 + (Cc.y - ar1) * d1 + (Cc.z - ar2) * d2;
                            dL_dalpha *= T;
                            last_alpha = alpha;
                            dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot_dpixel;
                            const float dL_dG = B.y * dL_dalpha;
                            const float gdx = G * dx, gdy = G * dy;
                            const float dG_ddelx = -gdx * A.z - gdy * A.w;
                            const float dG_ddely = -gdy * B.x - gdx * A.w;
                            v[0] = dL_dG * dG_ddelx * ddelx_dx;
                            v[1] = dL_dG * dG_ddely * ddely_dy;
                            v[2] = -0.5f * gdx * dx * dL_dG;
                            v[3] = -0.5f * gdx * dy * dL_dG;
                            v[4] = -0.5f * gdy * dy * dL_dG;
                            v[5] = G * dL_dalpha;
                            v[6] = dchannel_dcolor * d0;
                            v[7] = dchannel_dcolor * d1;
                            v8 = dchannel_dcolor * d2;
                        }
                    }
                    if (!__any_sync(0xffffffffu, contrib)) continue;
                    // reduce-scatter 8 values over the warp: 4 + 2 + 1 + 1 + 1 shuffles
                    {
                        const bool hi = lane & 16;
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const float send = hi ? v[k] : v[k + 4];
                            const float keep = hi ? v[k + 4] : v[k];
                            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                        }
                    }
                    {
                        const bool hi = lane & 8;
#pragma unroll
                        for (int k = 0; k < 2; k++) {
                            const float send = hi ? v[k] : v[k + 2];
                            const float keep = hi ? v[k + 2] : v[k];
                            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                        }
                    }
                    {
                        const bool hi = lane & 4;
                        const float send = hi ? v[0] : v[1];
                        const float keep = hi ? v[1] : v[0];
                        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                    }
                    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
                    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v8 += __shfl_xor_sync(0xffffffffu, v8, o);
                    // lanes 0,4,...,28 hold sums 0..7 (index = bits 4,3,2 of the lane); lane 1 adds sum 8
                    const uint32_t id = s_ids[buf][e];
                    float* dst = acc + (size_t)id * 12;
                    if ((lane & 3) == 0) {
                        const int k = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
                        atomicAdd(dst + k, v[0]);
                    } else if (lane == 1) {
                        atomicAdd(dst + 8, v8);
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
    dim3 grid(IL.tiles_x, IL.tiles_y);
    {
        StageTimer _t(ST_BLEND_BWD, s);
        blend_backward_kernel<<<grid, BLEND_THREADS, 0, s>>>(
            reinterpret_cast<const uint2*>(image + IL.ranges), point_list, reinterpret_cast<const SplatRec*>(geom + GL.rec),
            p.W, p.H, p.background, reinterpret_cast<const float*>(image + IL.final_T),
            reinterpret_cast<const uint32_t*>(image + IL.n_contrib), reinterpret_cast<const uint32_t*>(image + IL.tile_max_contrib),
            dL_dpix, reinterpret_cast<float*>(geom + GL.acc));
        GSB_LAUNCH_CHECK();
    }
    return GSB_OK;
}

}  // namespace gsb
