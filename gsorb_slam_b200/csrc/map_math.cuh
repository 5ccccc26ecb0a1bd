// map_math.cuh -- the activation prologue of Render::StartSplatting (src/Render.cc:750-759) and the torch::optim::Adam update
// (src/Gaussian.cc:131-175) as row-level device functions, shared by the stand-alone kernels of extras.cu and the fused
// map_update_kernel (map_update.cu) so that both see the same values bit for bit.
#pragma once
#include "common.cuh"

namespace gsb {

// row r of Tcw [mean; 1]
__device__ __forceinline__ float to_camera(const float* __restrict__ Tcw, int r, float x, float y, float z)
{
    return fmaf(Tcw[4 * r + 2], z, fmaf(Tcw[4 * r + 1], y, Tcw[4 * r] * x)) + Tcw[4 * r + 3];
}
__device__ __forceinline__ float sigmoid_act(float logit) { return 1.0f / (1.0f + expf(-logit)); }
__device__ __forceinline__ float quat_norm(float a, float b, float c, float d)
{
    return fmaxf(sqrtf(a * a + b * b + c * c + d * d), 1e-12f);  // F::normalize eps
}

__device__ __forceinline__ void adam_update(float& p, float gr, float& mi, float& vi, float step_size, float omb1, float beta2, float omb2,
                                            float eps, float inv_sqrt_bc2)
{
    mi = mi + (gr - mi) * omb1;
    vi = vi * beta2 + omb2 * gr * gr;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    p = p - step_size * (mi / denom);
}

// The same update with sqrt.approx + rcp.approx instead of the IEEE sequences (about 10 instead of 45 instructions per element:
// map_update_kernel is bound by instruction issue, the stand-alone Adam kernel by HBM).  The quotient is off by <= 3 ulp, i.e.
// the parameter moves by step_size * (1 +- 4e-7) * m / denom; vi = 0 -> denom = eps and mi = 0 -> no move, as above.
__device__ __forceinline__ void adam_update_fast(float& p, float gr, float& mi, float& vi, float step_size, float omb1, float beta2,
                                                 float omb2, float eps, float inv_sqrt_bc2)
{
    mi = mi + (gr - mi) * omb1;
    vi = vi * beta2 + omb2 * gr * gr;
    float rt;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rt) : "f"(vi));
    const float denom = rt * inv_sqrt_bc2 + eps;
    p = p - step_size * __fdividef(mi, denom);
}

// Warp-reduce 12 partial sums by halving (xor 16: 12 -> 6 values per lane, xor 8: 6 -> 3, then a butterfly over xor 4, 2, 1:
// 18 shuffles instead of 60): lanes 0 / 8 / 16 / 24 end up with the warp's sums 0-2 / 3-5 / 6-8 / 9-11 in w[0..2].
__device__ __forceinline__ void warp_reduce12(const float* v, float* w)
{
    const int lane = lane_id();
    const bool h16 = lane & 16, h8 = lane & 8;
    float u[6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        const float send = h16 ? v[k] : v[k + 6], keep = h16 ? v[k + 6] : v[k];
        u[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float send = h8 ? u[k] : u[k + 3], keep = h8 ? u[k + 3] : u[k];
        w[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1)
#pragma unroll
        for (int k = 0; k < 3; k++) w[k] += __shfl_xor_sync(0xffffffffu, w[k], o);
}
// ... and over the CTA, added to out[12] with one atomic per value per CTA
template <int THREADS>
__device__ __forceinline__ void reduce12_scatter_and_add(const float* v, float* __restrict__ out)
{
    __shared__ float s_part[THREADS / 32][12];
    float w[3];
    warp_reduce12(v, w);
    const int lane = lane_id();
    if ((lane & 7) == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) s_part[threadIdx.x >> 5][3 * (lane >> 3) + k] = w[k];
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        float x = 0.f;
#pragma unroll
        for (int k = 0; k < THREADS / 32; k++) x += s_part[k][threadIdx.x];
        atomicAdd(out + threadIdx.x, x);
    }
}

// Block-reduce 12 partial sums and add them to out[12] with one atomic per value per CTA.
template <int THREADS>
__device__ __forceinline__ void reduce12_and_add(float* v, float* __restrict__ out)
{
    __shared__ float s_part[THREADS / 32][12];
#pragma unroll
    for (int k = 0; k < 12; k++) {
        float x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane_id() == 0) s_part[threadIdx.x >> 5][k] = x;
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        float x = 0.f;
#pragma unroll
        for (int w = 0; w < THREADS / 32; w++) x += s_part[w][threadIdx.x];
        atomicAdd(out + threadIdx.x, x);
    }
}

}  // namespace gsb
