// densify.cu -- GPU twin of the reference's CPU back-projection, sm_100a.
//
// Replaces the per-pixel double loops Render::ProjectPixel (src/Render.cc:617-655) and Render::InitGaussianPoint
// (:666-707) -- executed on the host for every frame, with a cv::Mat 4x4 product per pixel -- plus the parameter
// initialisation of Gaussian::AddGaussianPoints (src/Gaussian.cc:50-74, initScalarMethod = SinglePixel, the method of
// every shipped YAML): for every selected pixel with positive depth, in ROW-MAJOR pixel order (the order the CPU loop
// emits, which fixes the Gaussian ids and therefore the sort's tie rule),
//     p_cam = ((j - cx) z / fx, (i - cy) z / fy, z),  p_world = Twc [p_cam; 1],
//     mean = p_world, rgb = image[:, i, j], log_scale = log(sqrt((p_world.z / ((fx + fy) / 2))^2)) x 3,
//     unnorm_quat = (1, 0, 0, 0), logit_opacity = 1.
// Order-preserving stream compaction in three small kernels (per-block counts, one-CTA scan, ballot-ranked scatter);
// the row count and max camera depth (Render::mMaxZ) stay on the device.
#include "common.cuh"

namespace gsb {

constexpr int DN_THREADS = 256;

__device__ __forceinline__ bool dn_selected(const uint8_t* __restrict__ mask, const float* __restrict__ depth, int i)
{
    return (!mask || mask[i] >= 250) && depth[i] > 0.f;   // Render.cc:627-631
}

__global__ void __launch_bounds__(DN_THREADS)
densify_count_kernel(int HW, const uint8_t* __restrict__ mask, const float* __restrict__ depth, uint32_t* __restrict__ block_count)
{
    const int i = blockIdx.x * DN_THREADS + threadIdx.x;
    const int c = __syncthreads_count(i < HW && dn_selected(mask, depth, i));
    if (threadIdx.x == 0) block_count[blockIdx.x] = (uint32_t)c;
}

__global__ void __launch_bounds__(1024)
densify_scan_kernel(int blocks, uint32_t* __restrict__ block_count /* in: counts, out: exclusive offsets */, int* __restrict__ count_out)
{
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < blocks; base += 1024) {
        const int i = base + threadIdx.x;
        const uint32_t x = i < blocks ? block_count[i] : 0u;
        uint32_t v = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane_id() >= (uint32_t)o) v += t;
        }
        if (lane_id() == 31) s_warp[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = s_warp[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane_id() >= (uint32_t)o) w += t;
            }
            s_warp[threadIdx.x] = w;
        }
        __syncthreads();
        const uint32_t warp_excl = (threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : 0u;
        const uint32_t carry = s_carry;
        if (i < blocks) block_count[i] = carry + warp_excl + v - x;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + warp_excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *count_out = (int)s_carry;
}

struct DensifyParams {
    int W, H, capacity;
    const uint8_t* mask;
    const float* depth;
    const float* image;   // [3,H,W]
    float fx, fy, cx, cy;
    float Twc[16];        // row-major
    float *means, *rgb, *log_scales, *quats, *logit_opacities;
    float* max_z;
};

__global__ void __launch_bounds__(DN_THREADS)
densify_scatter_kernel(DensifyParams q, const uint32_t* __restrict__ block_offset)
{
    __shared__ uint32_t s_warp[DN_THREADS / 32];
    const int HW = q.W * q.H;
    const int i = blockIdx.x * DN_THREADS + threadIdx.x;
    const bool sel = i < HW && dn_selected(q.mask, q.depth, i);
    const uint32_t b = __ballot_sync(0xffffffffu, sel);
    if (lane_id() == 0) s_warp[threadIdx.x >> 5] = __popc(b);
    __syncthreads();
    uint32_t pre = block_offset[blockIdx.x];
    for (uint32_t w = 0; w < (threadIdx.x >> 5); w++) pre += s_warp[w];
    float z = 0.f;
    if (sel) {
        const uint32_t row = pre + __popc(b & ((1u << lane_id()) - 1u));
        z = q.depth[i];
        if ((int)row < q.capacity) {
            const int pi = i / q.W, pj = i % q.W;
            const float x = __fdiv_rn(__fmul_rn((float)pj - q.cx, z), q.fx);   // (j - cx) * z / fx, Render.cc:635
            const float y = __fdiv_rn(__fmul_rn((float)pi - q.cy, z), q.fy);
            const float* T = q.Twc;
            const float xw = T[0] * x + T[1] * y + T[2] * z + T[3];
            const float yw = T[4] * x + T[5] * y + T[6] * z + T[7];
            const float zw = T[8] * x + T[9] * y + T[10] * z + T[11];
            const size_t r = row;
            q.means[3 * r] = xw; q.means[3 * r + 1] = yw; q.means[3 * r + 2] = zw;
            if (q.rgb) {
                q.rgb[3 * r] = q.image[i]; q.rgb[3 * r + 1] = q.image[(size_t)HW + i]; q.rgb[3 * r + 2] = q.image[2 * (size_t)HW + i];
            }
            if (q.log_scales) {   // Gaussian.cc:66-69: the WORLD z of the point, not its camera depth
                const float t = zw / ((q.fx + q.fy) * 0.5f);
                const float ls = logf(sqrtf(t * t));
                q.log_scales[3 * r] = ls; q.log_scales[3 * r + 1] = ls; q.log_scales[3 * r + 2] = ls;
            }
            if (q.quats) { q.quats[4 * r] = 1.f; q.quats[4 * r + 1] = 0.f; q.quats[4 * r + 2] = 0.f; q.quats[4 * r + 3] = 0.f; }
            if (q.logit_opacities) q.logit_opacities[r] = 1.f;
        }
    }
    if (q.max_z) {   // Render::mMaxZ (positive floats order like their bit patterns)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) z = fmaxf(z, __shfl_xor_sync(0xffffffffu, z, o));
        if (lane_id() == 0 && z > 0.f) atomicMax(reinterpret_cast<int*>(q.max_z), __float_as_int(z));
    }
}

// ---- prune: order-preserving compaction of per-Gaussian rows --------------------------------------------
// Gaussian::RemovePoints / PruneOptimizer (src/Gaussian.cc:209-239) index_select every parameter tensor and both Adam
// moments by the surviving rows, one tensor at a time; here one scan of the keep flags and ONE scatter kernel move up to
// 16 row-major tensors (parameters, exp_avg, exp_avg_sq of every group).
constexpr int PRUNE_MAX_TENSORS = 16;
struct PruneTensors {
    const float* src[PRUNE_MAX_TENSORS];
    float* dst[PRUNE_MAX_TENSORS];
    int width[PRUNE_MAX_TENSORS];
    int n;
};

__global__ void __launch_bounds__(DN_THREADS)
prune_count_kernel(int P, const uint8_t* __restrict__ keep, uint32_t* __restrict__ block_count)
{
    const int i = blockIdx.x * DN_THREADS + threadIdx.x;
    const int c = __syncthreads_count(i < P && keep[i] != 0);
    if (threadIdx.x == 0) block_count[blockIdx.x] = (uint32_t)c;
}

__global__ void __launch_bounds__(DN_THREADS)
prune_scatter_kernel(int P, const uint8_t* __restrict__ keep, const uint32_t* __restrict__ block_offset, PruneTensors T)
{
    __shared__ uint32_t s_warp[DN_THREADS / 32];
    const int i = blockIdx.x * DN_THREADS + threadIdx.x;
    const bool sel = i < P && keep[i] != 0;
    const uint32_t b = __ballot_sync(0xffffffffu, sel);
    if (lane_id() == 0) s_warp[threadIdx.x >> 5] = __popc(b);
    __syncthreads();
    uint32_t row = block_offset[blockIdx.x];
    for (uint32_t w = 0; w < (threadIdx.x >> 5); w++) row += s_warp[w];
    if (!sel) return;
    row += __popc(b & ((1u << lane_id()) - 1u));
    for (int t = 0; t < T.n; t++) {
        const int w = T.width[t];
        const float* s = T.src[t] + (size_t)i * w;
        float* d = T.dst[t] + (size_t)row * w;
        for (int k = 0; k < w; k++) d[k] = s[k];
    }
}

__global__ void __launch_bounds__(DN_THREADS)
low_opacity_keep_kernel(int P, const float* __restrict__ logit, float threshold, uint8_t* __restrict__ keep)
{
    const int i = blockIdx.x * DN_THREADS + threadIdx.x;
    if (i < P) keep[i] = (1.0f / (1.0f + expf(-logit[i])) < threshold) ? 0 : 1;   // Gaussian::RemoveLowOpcitiesGaussian
}

}  // namespace gsb

using namespace gsb;

extern "C" {

size_t gsb_backproject_scratch_bytes(int width, int height)
{
    if (width <= 0 || height <= 0) return 0;
    return align_up(((size_t)width * height + DN_THREADS - 1) / DN_THREADS * sizeof(uint32_t), 256);
}

int gsb_backproject(int width, int height, const uint8_t* mask, const float* depth, const float* image, float fx, float fy,
                    float cx, float cy, const float* Twc_host, int capacity, float* means, float* rgb, float* log_scales,
                    float* unnorm_quats, float* logit_opacities, int* count, float* max_z, void* scratch, size_t scratch_bytes,
                    gsb_stream_t stream)
{
    if (width <= 0 || height <= 0 || !depth || !Twc_host || !means || !count || capacity < 0 || (rgb && !image)) {
        set_error("backproject: size, depth, Twc, means, count are required (and image when rgb is requested)");
        return GSB_ERR_INVALID_ARGUMENT;
    }
    if (!scratch || scratch_bytes < gsb_backproject_scratch_bytes(width, height)) {
        set_error("backproject: scratch too small");
        return GSB_ERR_WORKSPACE;
    }
    const int HW = width * height, blocks = (HW + DN_THREADS - 1) / DN_THREADS;
    cudaStream_t s = (cudaStream_t)stream;
    uint32_t* bc = static_cast<uint32_t*>(scratch);
    DensifyParams q;
    q.W = width; q.H = height; q.capacity = capacity; q.mask = mask; q.depth = depth; q.image = image;
    q.fx = fx; q.fy = fy; q.cx = cx; q.cy = cy;
    for (int k = 0; k < 16; k++) q.Twc[k] = Twc_host[k];
    q.means = means; q.rgb = rgb; q.log_scales = log_scales; q.quats = unnorm_quats; q.logit_opacities = logit_opacities;
    q.max_z = max_z;
    StageTimer _t(ST_OTHER, s);
    densify_count_kernel<<<blocks, DN_THREADS, 0, s>>>(HW, mask, depth, bc);
    GSB_LAUNCH_CHECK();
    densify_scan_kernel<<<1, 1024, 0, s>>>(blocks, bc, count);
    GSB_LAUNCH_CHECK();
    densify_scatter_kernel<<<blocks, DN_THREADS, 0, s>>>(q, bc);
    GSB_LAUNCH_CHECK();
    return GSB_OK;
}

size_t gsb_prune_scratch_bytes(int P)
{
    return align_up(((size_t)(P > 0 ? P : 0) + DN_THREADS - 1) / DN_THREADS * sizeof(uint32_t) + sizeof(uint32_t), 256);
}

int gsb_low_opacity_keep(int P, const float* logit_opacities, float threshold, uint8_t* keep, gsb_stream_t stream)
{
    if (P < 0 || (P > 0 && (!logit_opacities || !keep))) {
        set_error("low_opacity_keep: bad arguments");
        return GSB_ERR_INVALID_ARGUMENT;
    }
    if (P == 0) return GSB_OK;
    low_opacity_keep_kernel<<<(P + DN_THREADS - 1) / DN_THREADS, DN_THREADS, 0, (cudaStream_t)stream>>>(P, logit_opacities, threshold, keep);
    GSB_LAUNCH_CHECK();
    return GSB_OK;
}

int gsb_prune_rows(int P, const uint8_t* keep, int ntensors, const float* const* src_host, float* const* dst_host,
                   const int* widths_host, int* count, void* scratch, size_t scratch_bytes, gsb_stream_t stream)
{
    if (P < 0 || ntensors < 0 || ntensors > PRUNE_MAX_TENSORS || !count || (P > 0 && !keep) ||
        (ntensors > 0 && (!src_host || !dst_host || !widths_host))) {
        set_error("prune_rows: need P >= 0, 0 <= ntensors <= %d, keep, count and the tensor tables", PRUNE_MAX_TENSORS);
        return GSB_ERR_INVALID_ARGUMENT;
    }
    if (!scratch || scratch_bytes < gsb_prune_scratch_bytes(P)) {
        set_error("prune_rows: scratch too small");
        return GSB_ERR_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    PruneTensors T;
    T.n = ntensors;
    for (int t = 0; t < PRUNE_MAX_TENSORS; t++) {
        T.src[t] = t < ntensors ? src_host[t] : nullptr;
        T.dst[t] = t < ntensors ? dst_host[t] : nullptr;
        T.width[t] = t < ntensors ? widths_host[t] : 0;
        if (t < ntensors && (P > 0) && (!T.src[t] || !T.dst[t] || T.width[t] <= 0)) {
            set_error("prune_rows: tensor %d has a NULL pointer or a non-positive width", t);
            return GSB_ERR_INVALID_ARGUMENT;
        }
    }
    const int blocks = (P + DN_THREADS - 1) / DN_THREADS;
    uint32_t* bc = static_cast<uint32_t*>(scratch);
    if (blocks == 0) {
        GSB_CUDA_CHECK(cudaMemsetAsync(count, 0, sizeof(int), s));
        return GSB_OK;
    }
    StageTimer _t(ST_OTHER, s);
    prune_count_kernel<<<blocks, DN_THREADS, 0, s>>>(P, keep, bc);
    GSB_LAUNCH_CHECK();
    densify_scan_kernel<<<1, 1024, 0, s>>>(blocks, bc, count);
    GSB_LAUNCH_CHECK();
    prune_scatter_kernel<<<blocks, DN_THREADS, 0, s>>>(P, keep, bc, T);
    GSB_LAUNCH_CHECK();
    return GSB_OK;
}

}  // extern "C"
