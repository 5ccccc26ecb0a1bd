// extras.cu -- fused caller-side pieces around the rasterizer, sm_100a.
//
//   prologue           Render::StartSplatting's torch prologue, src/Render.cc:750-759
//                      (Tcw.repeat(N,1,1).bmm([mean;1]), sigmoid, normalize, exp) in ONE pass
//                      with no N x 4 x 4 intermediate.
//   prologue_backward  what autograd does for that prologue, plus the camera-pose gradient
//                      dL/dTcw[0:3,:] = sum_i g_i [p_i;1]^T reduced on chip.
//   pose_grad          that reduction alone.
//   adam_step          torch::optim::Adam step of src/Gaussian.cc:131-175 over a flat block.
#include "common.cuh"
#include "map_math.cuh"

namespace gsb {

constexpr int EX_THREADS = 256;

__global__ void __launch_bounds__(EX_THREADS)
prologue_kernel(int P, const float* __restrict__ Tcw, const float* __restrict__ means_world,
                const float* __restrict__ logit, const float* __restrict__ quats, const float* __restrict__ log_scales,
                float* __restrict__ means_cam, float* __restrict__ opac, float* __restrict__ rot, float* __restrict__ scales)
{
    const int idx = blockIdx.x * EX_THREADS + threadIdx.x;
    if (idx >= P) return;
    const size_t i = (size_t)idx;
    if (means_cam) {
        const float x = means_world[3 * i], y = means_world[3 * i + 1], z = means_world[3 * i + 2];
#pragma unroll
        for (int r = 0; r < 3; r++) means_cam[3 * i + r] = to_camera(Tcw, r, x, y, z);
    }
    if (opac) opac[i] = sigmoid_act(logit[i]);
    if (rot) {
        const float a = quats[4 * i], b = quats[4 * i + 1], c = quats[4 * i + 2], d = quats[4 * i + 3];
        const float nrm = quat_norm(a, b, c, d);
        rot[4 * i] = a / nrm; rot[4 * i + 1] = b / nrm; rot[4 * i + 2] = c / nrm; rot[4 * i + 3] = d / nrm;
    }
    if (scales) {
#pragma unroll
        for (int k = 0; k < 3; k++) scales[3 * i + k] = expf(log_scales[3 * i + k]);
    }
}

__global__ void __launch_bounds__(EX_THREADS)
prologue_backward_kernel(int P, const float* __restrict__ Tcw, const float* __restrict__ means_world,
                         const float* __restrict__ logit, const float* __restrict__ quats,
                         const float* __restrict__ log_scales, const float* __restrict__ g_means,
                         const float* __restrict__ g_opac, const float* __restrict__ g_rot, const float* __restrict__ g_scales,
                         float* __restrict__ d_means, float* __restrict__ d_logit, float* __restrict__ d_quats,
                         float* __restrict__ d_log_scales, float* __restrict__ dTcw)
{
    float part[12];
#pragma unroll
    for (int k = 0; k < 12; k++) part[k] = 0.f;
    for (int idx = blockIdx.x * EX_THREADS + threadIdx.x; idx < P; idx += gridDim.x * EX_THREADS) {
        const size_t i = (size_t)idx;
        if (g_means) {
            const float g0 = g_means[3 * i], g1 = g_means[3 * i + 1], g2 = g_means[3 * i + 2];
            if (d_means) {
#pragma unroll
                for (int c = 0; c < 3; c++) d_means[3 * i + c] = Tcw[c] * g0 + Tcw[4 + c] * g1 + Tcw[8 + c] * g2;
            }
            if (dTcw) {
                const float x = means_world[3 * i], y = means_world[3 * i + 1], z = means_world[3 * i + 2];
                part[0] += g0 * x; part[1] += g0 * y; part[2] += g0 * z; part[3] += g0;
                part[4] += g1 * x; part[5] += g1 * y; part[6] += g1 * z; part[7] += g1;
                part[8] += g2 * x; part[9] += g2 * y; part[10] += g2 * z; part[11] += g2;
            }
        }
        if (d_logit && g_opac) {
            const float s = sigmoid_act(logit[i]);
            d_logit[i] = g_opac[i] * s * (1.0f - s);
        }
        if (d_quats && g_rot) {
            const float a = quats[4 * i], b = quats[4 * i + 1], c = quats[4 * i + 2], d = quats[4 * i + 3];
            const float nrm = quat_norm(a, b, c, d);
            const float na = a / nrm, nb = b / nrm, nc = c / nrm, nd = d / nrm;
            const float ga = g_rot[4 * i], gb = g_rot[4 * i + 1], gc = g_rot[4 * i + 2], gd = g_rot[4 * i + 3];
            const float dot = na * ga + nb * gb + nc * gc + nd * gd;
            d_quats[4 * i] = (ga - na * dot) / nrm;
            d_quats[4 * i + 1] = (gb - nb * dot) / nrm;
            d_quats[4 * i + 2] = (gc - nc * dot) / nrm;
            d_quats[4 * i + 3] = (gd - nd * dot) / nrm;
        }
        if (d_log_scales && g_scales) {
#pragma unroll
            for (int k = 0; k < 3; k++) d_log_scales[3 * i + k] = g_scales[3 * i + k] * expf(log_scales[3 * i + k]);
        }
    }
    if (dTcw) reduce12_and_add<EX_THREADS>(part, dTcw);
}

// The whole packed block in ONE launch: group g covers [bound[g-1], bound[g]) and has its own learning rate
// (torch::optim::Adam with one param group per tensor, src/Gaussian.cc:158-175).
struct AdamGroups {
    long long bound[8];   // exclusive end of each group
    float step_size[8];   // lr_g / bias_correction1
    int n;
};
__device__ __forceinline__ float adam_group_step(const AdamGroups& G, long long i)
{
    float step_size = G.step_size[0];
#pragma unroll
    for (int g = 1; g < 8; g++)
        if (g < G.n && i >= G.bound[g - 1]) step_size = G.step_size[g];
    return step_size;
}
// VEC = true: 16-byte accesses, four elements per thread and iteration (the kernel moves 7 x 4 bytes per element and nothing else:
// 392 MB per step at 1 M Gaussians); the group of a quad is looked up once unless a group boundary falls inside it.  total4 quads,
// then the scalar tail.  VEC = false: unaligned pointers.
template <bool VEC>
__global__ void __launch_bounds__(EX_THREADS)
adam_groups_kernel(long long total, float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ m,
                   float* __restrict__ v, float omb1, float beta2, float omb2, float eps, float inv_sqrt_bc2, AdamGroups G)
{
    const long long tid = (long long)blockIdx.x * EX_THREADS + threadIdx.x, stride = (long long)gridDim.x * EX_THREADS;
    long long done = 0;
    if (VEC) {
        const long long total4 = total >> 2;
        for (long long q = tid; q < total4; q += stride) {
            const long long i = q << 2;
            float4 p4 = reinterpret_cast<float4*>(param)[q], m4 = reinterpret_cast<float4*>(m)[q], v4 = reinterpret_cast<float4*>(v)[q];
            const float4 g4 = reinterpret_cast<const float4*>(grad)[q];
            const float s0 = adam_group_step(G, i), s3 = adam_group_step(G, i + 3);
            const float s1 = s0 == s3 ? s0 : adam_group_step(G, i + 1), s2 = s0 == s3 ? s0 : adam_group_step(G, i + 2);
            adam_update(p4.x, g4.x, m4.x, v4.x, s0, omb1, beta2, omb2, eps, inv_sqrt_bc2);
            adam_update(p4.y, g4.y, m4.y, v4.y, s1, omb1, beta2, omb2, eps, inv_sqrt_bc2);
            adam_update(p4.z, g4.z, m4.z, v4.z, s2, omb1, beta2, omb2, eps, inv_sqrt_bc2);
            adam_update(p4.w, g4.w, m4.w, v4.w, s3, omb1, beta2, omb2, eps, inv_sqrt_bc2);
            reinterpret_cast<float4*>(m)[q] = m4;
            reinterpret_cast<float4*>(v)[q] = v4;
            reinterpret_cast<float4*>(param)[q] = p4;
        }
        done = total4 << 2;
    }
    for (long long i = done + tid; i < total; i += stride) {
        float pi = param[i], mi = m[i], vi = v[i];
        adam_update(pi, grad[i], mi, vi, adam_group_step(G, i), omb1, beta2, omb2, eps, inv_sqrt_bc2);
        m[i] = mi;
        v[i] = vi;
        param[i] = pi;
    }
}

// ---- scale regularisers of the mapping loss (src/Render.cc:462-469) --------------------------------------------------
//   big   = where(exp(log_scales) > maxScalar)[0]      row index once per AXIS that exceeds (a row can appear 1-3 times)
//   reg_scalar = sum_big (max_axis exp(ls) - maxScalar),   reg_long = mean_big (max_axis exp(ls) - min_axis exp(ls))
// acc[0] = number of selected (row, axis) pairs C, acc[1] = reg_scalar, acc[2] = sum of (max - min) over the selection.
__global__ void __launch_bounds__(EX_THREADS)
scale_reg_sum_kernel(int P, const float* __restrict__ log_scales, float max_scalar, float* __restrict__ acc)
{
    float c = 0.f, a = 0.f, b = 0.f;
    for (int i = blockIdx.x * EX_THREADS + threadIdx.x; i < P; i += gridDim.x * EX_THREADS) {
        const float s0 = expf(log_scales[3 * (size_t)i]), s1 = expf(log_scales[3 * (size_t)i + 1]), s2 = expf(log_scales[3 * (size_t)i + 2]);
        const float n = (s0 > max_scalar ? 1.f : 0.f) + (s1 > max_scalar ? 1.f : 0.f) + (s2 > max_scalar ? 1.f : 0.f);
        if (n > 0.f) {
            const float mx = fmaxf(s0, fmaxf(s1, s2)), mn = fminf(s0, fminf(s1, s2));
            c += n;
            a += n * (mx - max_scalar);
            b += n * (mx - mn);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c += __shfl_xor_sync(0xffffffffu, c, o);
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane_id() == 0 && c > 0.f) {
        atomicAdd(acc, c);
        atomicAdd(acc + 1, a);
        atomicAdd(acc + 2, b);
    }
}

// d(w_scalar reg_scalar + w_long reg_long) / d(log_scales), ADDED to d_log_scales; terms = {reg_scalar, reg_long, C, 0}
__global__ void __launch_bounds__(EX_THREADS)
scale_reg_apply_kernel(int P, const float* __restrict__ log_scales, float max_scalar, float w_scalar, float w_long,
                       const float* __restrict__ acc, float* __restrict__ d_log_scales, float* __restrict__ terms)
{
    const float C = acc[0];
    if (blockIdx.x == 0 && threadIdx.x == 0 && terms) {
        terms[0] = acc[1];
        terms[1] = acc[2] / C;   // mean over an empty selection is NaN in the reference too (the gradients are unaffected)
        terms[2] = C;
        terms[3] = 0.f;
    }
    if (!(C > 0.f) || !d_log_scales) return;
    const float wl = w_long / C;
    for (int i = blockIdx.x * EX_THREADS + threadIdx.x; i < P; i += gridDim.x * EX_THREADS) {
        const float s[3] = {expf(log_scales[3 * (size_t)i]), expf(log_scales[3 * (size_t)i + 1]), expf(log_scales[3 * (size_t)i + 2])};
        const float n = (s[0] > max_scalar ? 1.f : 0.f) + (s[1] > max_scalar ? 1.f : 0.f) + (s[2] > max_scalar ? 1.f : 0.f);
        if (n > 0.f) {
            int imax = 0, imin = 0;   // first maximum / first minimum, as torch.max / torch.min report them
            if (s[1] > s[imax]) imax = 1;
            if (s[2] > s[imax]) imax = 2;
            if (s[1] < s[imin]) imin = 1;
            if (s[2] < s[imin]) imin = 2;
            d_log_scales[3 * (size_t)i + imax] += n * (w_scalar + wl) * s[imax];   // d exp(ls) / d ls = exp(ls)
            d_log_scales[3 * (size_t)i + imin] -= n * wl * s[imin];
        }
    }
}

static int grid_for(long long n)
{
    long long g = (n + EX_THREADS - 1) / EX_THREADS;
    const long long cap = (long long)NUM_SMS * 16;
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

int launch_prologue(int P, const float* Tcw, const float* means_world, const float* logit, const float* quats,
                    const float* log_scales, float* means_cam, float* opac, float* rot, float* scales, cudaStream_t s)
{
    if (P <= 0) return GSB_OK;
    prologue_kernel<<<(P + EX_THREADS - 1) / EX_THREADS, EX_THREADS, 0, s>>>(P, Tcw, means_world, logit, quats, log_scales,
                                                                             means_cam, opac, rot, scales);
    GSB_LAUNCH_CHECK();
    return GSB_OK;
}

int launch_prologue_backward(int P, const float* Tcw, const float* means_world, const float* logit, const float* quats,
                             const float* log_scales, const float* g_means, const float* g_opac, const float* g_rot,
                             const float* g_scales, float* d_means, float* d_logit, float* d_quats, float* d_log_scales,
                             float* dTcw, cudaStream_t s)
{
    if (dTcw) GSB_CUDA_CHECK(cudaMemsetAsync(dTcw, 0, 12 * sizeof(float), s));
    if (P <= 0) return GSB_OK;
    prologue_backward_kernel<<<grid_for(P), EX_THREADS, 0, s>>>(P, Tcw, means_world, logit, quats, log_scales, g_means, g_opac,
                                                                g_rot, g_scales, d_means, d_logit, d_quats, d_log_scales, dTcw);
    GSB_LAUNCH_CHECK();
    return GSB_OK;
}

int launch_adam(long long n, float* param, const float* grad, float* m, float* v, double lr, double beta1, double beta2,
                double eps, long long step, cudaStream_t s)
{
    if (n <= 0) return GSB_OK;
    // torch keeps the hyper-parameters as doubles and rounds each derived scalar once (1 - beta, lr / bias_correction1, ...)
    const double bc1 = 1.0 - pow(beta1, (double)step);
    const double bc2 = 1.0 - pow(beta2, (double)step);
    const float step_size = (float)(lr / bc1);
    const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    AdamGroups G;   // one group: the same (vectorised) kernel as the packed block
    G.n = 1;
    for (int g = 0; g < 8; g++) { G.bound[g] = n; G.step_size[g] = step_size; }
    const bool vec = ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    if (vec)
        adam_groups_kernel<true><<<grid_for((n + 3) / 4), EX_THREADS, 0, s>>>(n, param, grad, m, v, (float)(1.0 - beta1), (float)beta2,
                                                                            (float)(1.0 - beta2), (float)eps, inv_sqrt_bc2, G);
    else
        adam_groups_kernel<false><<<grid_for(n), EX_THREADS, 0, s>>>(n, param, grad, m, v, (float)(1.0 - beta1), (float)beta2,
                                                                   (float)(1.0 - beta2), (float)eps, inv_sqrt_bc2, G);
    GSB_LAUNCH_CHECK();
    return GSB_OK;
}

int launch_adam_groups(int ngroups, const long long* sizes, const float* lrs, float* param, const float* grad, float* m, float* v,
                       double beta1, double beta2, double eps, long long step, cudaStream_t s)
{
    AdamGroups G;
    G.n = ngroups;
    long long total = 0;
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    for (int g = 0; g < 8; g++) {
        if (g < ngroups) total += sizes[g];
        G.bound[g] = total;
        G.step_size[g] = g < ngroups ? (float)((double)lrs[g] / bc1) : 0.f;
    }
    if (total == 0) return GSB_OK;
    StageTimer _t(ST_OTHER, s);
    const bool vec = ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    if (vec)
        adam_groups_kernel<true><<<grid_for((total + 3) / 4), EX_THREADS, 0, s>>>(total, param, grad, m, v, (float)(1.0 - beta1), (float)beta2,
                                                                                (float)(1.0 - beta2), (float)eps, (float)(1.0 / sqrt(bc2)), G);
    else
        adam_groups_kernel<false><<<grid_for(total), EX_THREADS, 0, s>>>(total, param, grad, m, v, (float)(1.0 - beta1), (float)beta2,
                                                                       (float)(1.0 - beta2), (float)eps, (float)(1.0 / sqrt(bc2)), G);
    GSB_LAUNCH_CHECK();
    return GSB_OK;
}

// the reduction pass alone: acc[0..2] = {C, reg_scalar, sum of (max - min)} (map_update.cu folds the apply pass into its own kernel)
int launch_scale_regulariser_sum(int P, const float* log_scales, float max_scalar, float* acc, cudaStream_t s)
{
    GSB_CUDA_CHECK(cudaMemsetAsync(acc, 0, 4 * sizeof(float), s));
    if (P <= 0) return GSB_OK;
    StageTimer _t(ST_OTHER, s);
    scale_reg_sum_kernel<<<grid_for(P), EX_THREADS, 0, s>>>(P, log_scales, max_scalar, acc);
    GSB_LAUNCH_CHECK();
    return GSB_OK;
}

int launch_scale_regulariser(int P, const float* log_scales, float max_scalar, float w_scalar, float w_long, float* d_log_scales,
                             float* terms, float* acc, cudaStream_t s)
{
    if (int rc = launch_scale_regulariser_sum(P, log_scales, max_scalar, acc, s)) return rc;
    if (P <= 0) return GSB_OK;
    StageTimer _t(ST_OTHER, s);
    scale_reg_apply_kernel<<<grid_for(P), EX_THREADS, 0, s>>>(P, log_scales, max_scalar, w_scalar, w_long, acc, d_log_scales, terms);
    GSB_LAUNCH_CHECK();
    return GSB_OK;
}

}  // namespace gsb
