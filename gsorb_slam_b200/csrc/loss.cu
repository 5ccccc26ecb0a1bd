// loss.cu -- the producer of dL/dpixel for one mapping iteration, fused, sm_100a.
//
// Replaces the libtorch expression of Render::RenderForFrame (src/Render.cc:454-469) and its autograd:
//
//   image_loss    = lambda * mean|I - G| + (1 - lambda) * (1 - SSIM(I, G))          src/Utils.cc:39-44, :81-100
//   depth_loss    = mean over {G_d > 0}               |D - G_d|                      (D = depth pass channel 0)
//   surdepth_loss = mean over {G_d > 0, S > 0.99}     |M - G_d|                      (M = median depth, S = silhouette;
//                                                                                    no gradient: Rasterizer.cuh:210)
//   loss          = w_im * image_loss + w_d * depth_loss + w_sur * surdepth_loss     (+ scale regularisers: caller side)
//
// SSIM as the reference builds it: 11 x 11 window = outer product of the normalised 1-D weights
// exp(-floor((x - 11) / 2)^2 / (2 * 1.5^2)), x = 0..10 -- NOT centred: the offset comes from the reference's
// GaussianGenerator (src/Utils.cc:68-74) and is kept; five zero-padded grouped convolutions (mu1, mu2, E[x^2], E[y^2], E[xy]),
// C1 = 0.01^2, C2 = 0.03^2, mean over all pixels and channels.  libtorch launches ~40 kernels for the forward and backward of
// this expression; here two kernels do it with separable 11-tap passes in shared memory:
//   pass 1  per 16x16 tile and channel: moments -> SSIM value and its partials w.r.t. (mu1, E[x^2], E[xy]); L1 / depth sums;
//   pass 2  transposed (flipped-window) convolution of the three partial maps -> dL/dI, plus the L1 and depth gradients.
#include "common.cuh"

namespace gsb {

constexpr int LT = 16;            // output tile edge
constexpr int LW = 11;            // window taps
constexpr int LH = LT + LW - 1;   // halo edge (26)
constexpr float SSIM_C1 = 0.01f * 0.01f, SSIM_C2 = 0.03f * 0.03f;

struct LossTotals {   // device, zeroed by the launch
    float l1_sum, ssim_sum, depth_sum, sur_sum;
    float n_valid, n_valid_sur, pad0, pad1;
};

struct LossParams {
    int W, H;
    const float* color;      // [3,H,W] rendered
    const float* depth_sil;  // [2,H,W] rendered depth / silhouette (may be NULL: no depth terms)
    const float* median;     // [1,H,W] (may be NULL)
    const float* gt_color;   // [3,H,W]
    const float* gt_depth;   // [H,W]   (may be NULL)
    float lambda_, w_image, w_depth, w_sur;
    float win[LW];           // normalised 1-D window
    float* maps;             // [3 partials][3 channels][H][W] scratch
    LossTotals* totals;
    float* dL_dcolor;        // [3,H,W]
    float* dL_ddepth_sil;    // [2,H,W] or NULL
    float* loss_terms;       // [8]: l1, ssim, depth_l1, surdepth_l1, total, n_valid, n_valid_sur, -
};

__device__ __forceinline__ float block_sum_256(float v, float* s_red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; w++) t += s_red[w];
    return t;
}

__global__ void __launch_bounds__(LT * LT)
loss_stats_kernel(LossParams q)
{
    __shared__ float s_x[LH][LH + 1], s_y[LH][LH + 1];
    __shared__ float s_h[5][LH][LT + 1];
    __shared__ float s_red[8];
    const int tx = threadIdx.x % LT, ty = threadIdx.x / LT;
    const int x0 = blockIdx.x * LT, y0 = blockIdx.y * LT;
    const int px = x0 + tx, py = y0 + ty;
    const bool inside = px < q.W && py < q.H;
    const size_t HW = (size_t)q.W * q.H, pix = (size_t)py * q.W + px;
    float l1 = 0.f, ssim_acc = 0.f;
    const int ch = blockIdx.z;   // one CTA per (tile, channel): 3 x 1200 short CTAs fill the 148 x 8 slots better than 1200 long ones
    {
        const float* I = q.color + ch * HW;
        const float* G = q.gt_color + ch * HW;
        for (int i = threadIdx.x; i < LH * LH; i += LT * LT) {   // zero-padded halo (conv2d padding 5)
            const int hx = i % LH, hy = i / LH, gx = x0 + hx - LW / 2, gy = y0 + hy - LW / 2;
            const bool in = gx >= 0 && gx < q.W && gy >= 0 && gy < q.H;
            s_x[hy][hx] = in ? I[(size_t)gy * q.W + gx] : 0.f;
            s_y[hy][hx] = in ? G[(size_t)gy * q.W + gx] : 0.f;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < LH * LT; i += LT * LT) {   // horizontal 11-tap pass of the five moments
            const int cx = i % LT, hy = i / LT;
            float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
            for (int k = 0; k < LW; k++) {
                const float w = q.win[k], a = s_x[hy][cx + k], b = s_y[hy][cx + k];
                m1 = fmaf(w, a, m1); m2 = fmaf(w, b, m2);
                e11 = fmaf(w, a * a, e11); e22 = fmaf(w, b * b, e22); e12 = fmaf(w, a * b, e12);
            }
            s_h[0][hy][cx] = m1; s_h[1][hy][cx] = m2; s_h[2][hy][cx] = e11; s_h[3][hy][cx] = e22; s_h[4][hy][cx] = e12;
        }
        __syncthreads();
        float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
        for (int k = 0; k < LW; k++) {   // vertical pass
            const float w = q.win[k];
            m1 = fmaf(w, s_h[0][ty + k][tx], m1); m2 = fmaf(w, s_h[1][ty + k][tx], m2);
            e11 = fmaf(w, s_h[2][ty + k][tx], e11); e22 = fmaf(w, s_h[3][ty + k][tx], e22);
            e12 = fmaf(w, s_h[4][ty + k][tx], e12);
        }
        if (inside) {
            const float A = 2.f * m1 * m2 + SSIM_C1, B = 2.f * (e12 - m1 * m2) + SSIM_C2;
            const float D = m1 * m1 + m2 * m2 + SSIM_C1, E = (e11 - m1 * m1) + (e22 - m2 * m2) + SSIM_C2;
            const float inv = 1.f / (D * E), ssim = A * B * inv;
            ssim_acc += ssim;
            // partials w.r.t. the raw window moments of the rendered image (mu1, E[x^2], E[xy])
            q.maps[(0 * 3 + ch) * HW + pix] = 2.f * m2 * (B - A) * inv - ssim * 2.f * m1 * (E - D) * inv;
            q.maps[(1 * 3 + ch) * HW + pix] = -ssim / E;
            q.maps[(2 * 3 + ch) * HW + pix] = 2.f * A * inv;
            l1 += fabsf(s_x[ty + LW / 2][tx + LW / 2] - s_y[ty + LW / 2][tx + LW / 2]);
        }
    }
    float dsum = 0.f, ssum = 0.f, nv = 0.f, nvs = 0.f;
    if (ch == 0 && inside && q.gt_depth && q.depth_sil) {
        const float gd = q.gt_depth[pix];
        if (gd > 0.f) {
            nv = 1.f;
            dsum = fabsf(q.depth_sil[pix] - gd);
            if (q.median && q.depth_sil[HW + pix] > 0.99f) {
                nvs = 1.f;
                ssum = fabsf(q.median[pix] - gd);
            }
        }
    }
    const float t0 = block_sum_256(l1, s_red), t1 = block_sum_256(ssim_acc, s_red);
    if (threadIdx.x == 0) { atomicAdd(&q.totals->l1_sum, t0); atomicAdd(&q.totals->ssim_sum, t1); }
    if (ch == 0) {   // the depth terms travel with channel 0
        const float t2 = block_sum_256(dsum, s_red), t3 = block_sum_256(ssum, s_red), t4 = block_sum_256(nv, s_red), t5 = block_sum_256(nvs, s_red);
        if (threadIdx.x == 0) {
            atomicAdd(&q.totals->depth_sum, t2); atomicAdd(&q.totals->sur_sum, t3); atomicAdd(&q.totals->n_valid, t4);
            atomicAdd(&q.totals->n_valid_sur, t5);
        }
    }
}

__global__ void __launch_bounds__(LT * LT)
loss_grad_kernel(LossParams q)
{
    __shared__ float s_m[3][LH][LH + 1];
    __shared__ float s_h[3][LH][LT + 1];
    const int tx = threadIdx.x % LT, ty = threadIdx.x / LT;
    const int x0 = blockIdx.x * LT, y0 = blockIdx.y * LT;
    const int px = x0 + tx, py = y0 + ty;
    const bool inside = px < q.W && py < q.H;
    const size_t HW = (size_t)q.W * q.H, pix = (size_t)py * q.W + px;
    const float n_all = 3.f * (float)HW;
    const float g_ssim = -(1.f - q.lambda_) * q.w_image / n_all;   // dL / d ssim_p
    const float g_l1 = q.lambda_ * q.w_image / n_all;
    const int ch = blockIdx.z;
    {
        // d mu_p / d x_q = win[q - p + 5]: the partial maps are correlated with the FLIPPED window around q
        for (int i = threadIdx.x; i < LH * LH; i += LT * LT) {
            const int hx = i % LH, hy = i / LH, gx = x0 + hx - LW / 2, gy = y0 + hy - LW / 2;
            const bool in = gx >= 0 && gx < q.W && gy >= 0 && gy < q.H;   // no SSIM output outside the image
            const size_t gp = (size_t)gy * q.W + gx;
#pragma unroll
            for (int m = 0; m < 3; m++) s_m[m][hy][hx] = in ? q.maps[(m * 3 + ch) * HW + gp] : 0.f;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < LH * LT; i += LT * LT) {
            const int cx = i % LT, hy = i / LT;
            float a = 0.f, b = 0.f, c = 0.f;
#pragma unroll
            for (int k = 0; k < LW; k++) {
                const float w = q.win[LW - 1 - k];
                a = fmaf(w, s_m[0][hy][cx + k], a); b = fmaf(w, s_m[1][hy][cx + k], b); c = fmaf(w, s_m[2][hy][cx + k], c);
            }
            s_h[0][hy][cx] = a; s_h[1][hy][cx] = b; s_h[2][hy][cx] = c;
        }
        __syncthreads();
        if (inside) {
            float a = 0.f, b = 0.f, c = 0.f;
#pragma unroll
            for (int k = 0; k < LW; k++) {
                const float w = q.win[LW - 1 - k];
                a = fmaf(w, s_h[0][ty + k][tx], a); b = fmaf(w, s_h[1][ty + k][tx], b); c = fmaf(w, s_h[2][ty + k][tx], c);
            }
            const float x = q.color[ch * HW + pix], y = q.gt_color[ch * HW + pix];
            const float d = x - y, sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
            q.dL_dcolor[ch * HW + pix] = g_ssim * (a + 2.f * x * b + y * c) + g_l1 * sgn;
        }
    }
    const LossTotals t = *q.totals;
    if (ch == 0 && inside && q.dL_ddepth_sil) {
        float gd0 = 0.f;
        if (q.gt_depth && q.depth_sil && t.n_valid > 0.f) {
            const float gd = q.gt_depth[pix];
            if (gd > 0.f) {
                const float d = q.depth_sil[pix] - gd;
                gd0 = q.w_depth / t.n_valid * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
            }
        }
        q.dL_ddepth_sil[pix] = gd0;
        q.dL_ddepth_sil[HW + pix] = 0.f;   // the silhouette only gates a mask (detached, src/Render.cc:455)
    }
    if (blockIdx.x == 0 && blockIdx.y == 0 && ch == 0 && threadIdx.x == 0 && q.loss_terms) {
        const float l1 = t.l1_sum / n_all, ssim = t.ssim_sum / n_all;
        const float dl = t.n_valid > 0.f ? t.depth_sum / t.n_valid : 0.f, sl = t.n_valid_sur > 0.f ? t.sur_sum / t.n_valid_sur : 0.f;
        q.loss_terms[0] = l1; q.loss_terms[1] = ssim; q.loss_terms[2] = dl; q.loss_terms[3] = sl;
        q.loss_terms[4] = q.w_image * (q.lambda_ * l1 + (1.f - q.lambda_) * (1.f - ssim)) + q.w_depth * dl + q.w_sur * sl;
        q.loss_terms[5] = t.n_valid; q.loss_terms[6] = t.n_valid_sur; q.loss_terms[7] = 0.f;
    }
}


// ---- tracking iteration (src/Render.cc:1075-1093; L1LossForTracking, src/Utils.cc:45-52) ----------------------------------------
// mask = silhouette > 0.99 and gt depth is not NaN ("uncertainDepth"); image term = sum over the mask of |I - G| (three channels),
// depth term = sum over the mask of |D - G_d| with D the median depth (no gradient: include/Rasterizer.cuh:210) or the blended depth.
// One pass: per-pixel gradients sign(.) * weight * mask, block sums, one atomicAdd per block and term; the last block to finish
// forms the loss.  terms = {image_l1, depth_l1, loss, n_mask, -, -, -, completion counter (bits)}.
constexpr int TRK_THREADS = 256;
__global__ void __launch_bounds__(TRK_THREADS)
tracking_loss_kernel(int HW, const float* __restrict__ color, const float* __restrict__ depth_sil, const float* __restrict__ median,
                     const float* __restrict__ gt_color, const float* __restrict__ gt_depth, float w_image, float w_depth,
                     int use_sur, float* __restrict__ dC, float* __restrict__ dD, float* __restrict__ terms)
{
    float s_img = 0.f, s_dep = 0.f, s_n = 0.f;
    for (int i = blockIdx.x * TRK_THREADS + threadIdx.x; i < HW; i += gridDim.x * TRK_THREADS) {
        const float gd = gt_depth[i];
        const bool m = depth_sil[HW + i] > 0.99f && !(gd != gd);
        float dsum = 0.f;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float d = color[c * HW + i] - gt_color[c * HW + i];
            dC[c * HW + i] = m ? w_image * ((d > 0.f) - (d < 0.f)) : 0.f;
            dsum += fabsf(d);
        }
        const float dd = (use_sur ? median[i] : depth_sil[i]) - gd;
        if (dD) {
            dD[i] = (m && !use_sur) ? w_depth * ((dd > 0.f) - (dd < 0.f)) : 0.f;
            dD[HW + i] = 0.f;
        }
        if (m) { s_img += dsum; s_dep += fabsf(dd); s_n += 1.f; }
    }
    __shared__ float red[3][TRK_THREADS / 32];
    __shared__ bool last;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s_img += __shfl_xor_sync(0xffffffffu, s_img, o);
        s_dep += __shfl_xor_sync(0xffffffffu, s_dep, o);
        s_n += __shfl_xor_sync(0xffffffffu, s_n, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s_img; red[1][threadIdx.x >> 5] = s_dep; red[2][threadIdx.x >> 5] = s_n; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f, c = 0.f;
        for (int w = 0; w < TRK_THREADS / 32; w++) { a += red[0][w]; b += red[1][w]; c += red[2][w]; }
        atomicAdd(terms + 0, a); atomicAdd(terms + 1, b); atomicAdd(terms + 3, c);
        __threadfence();
        last = atomicAdd(reinterpret_cast<unsigned int*>(terms + 7), 1u) == gridDim.x - 1;
        if (last) {
            __threadfence();
            const float img = __ldcg(terms + 0), dep = __ldcg(terms + 1);
            terms[2] = w_image * img + w_depth * dep;
        }
    }
}

}  // namespace gsb

using namespace gsb;

extern "C" {

size_t gsb_loss_scratch_bytes(int width, int height)
{
    if (width <= 0 || height <= 0) return 0;
    return align_up(sizeof(LossTotals), 256) + (size_t)9 * width * height * sizeof(float);
}

int gsb_mapping_loss(int width, int height, const float* color, const float* depth_sil, const float* median_depth,
                     const float* gt_color, const float* gt_depth, float lambda_, float w_image, float w_depth, float w_surdepth,
                     float* dL_dcolor, float* dL_ddepth_sil, float* loss_terms, void* scratch, size_t scratch_bytes,
                     gsb_stream_t stream)
{
    if (width <= 0 || height <= 0 || !color || !gt_color || !dL_dcolor) {
        set_error("mapping_loss: image size, color, gt_color and dL_dcolor are required");
        return GSB_ERR_INVALID_ARGUMENT;
    }
    if (!scratch || scratch_bytes < gsb_loss_scratch_bytes(width, height)) {
        set_error("mapping_loss: scratch too small (%zu < %zu)", scratch_bytes, gsb_loss_scratch_bytes(width, height));
        return GSB_ERR_WORKSPACE;
    }
    LossParams q;
    q.W = width; q.H = height; q.color = color; q.depth_sil = depth_sil; q.median = median_depth; q.gt_color = gt_color;
    q.gt_depth = gt_depth; q.lambda_ = lambda_; q.w_image = w_image; q.w_depth = w_depth; q.w_sur = w_surdepth;
    // src/Utils.cc:68-74 in float, as the reference evaluates it
    float sum = 0.f;
    for (int x = 0; x < LW; x++) {
        const float f = floorf((float)(x - LW) / 2.f);
        q.win[x] = expf(-(f * f) / (2.f * 1.5f * 1.5f));
        sum += q.win[x];
    }
    for (int x = 0; x < LW; x++) q.win[x] /= sum;
    char* sc = static_cast<char*>(scratch);
    q.totals = reinterpret_cast<LossTotals*>(sc);
    q.maps = reinterpret_cast<float*>(sc + align_up(sizeof(LossTotals), 256));
    q.dL_dcolor = dL_dcolor; q.dL_ddepth_sil = dL_ddepth_sil; q.loss_terms = loss_terms;
    cudaStream_t s = (cudaStream_t)stream;
    GSB_CUDA_CHECK(cudaMemsetAsync(q.totals, 0, sizeof(LossTotals), s));
    const dim3 grid((width + LT - 1) / LT, (height + LT - 1) / LT, 3);   // z = colour channel
    {
        StageTimer _t(ST_OTHER, s);
        loss_stats_kernel<<<grid, LT * LT, 0, s>>>(q);
        GSB_LAUNCH_CHECK();
        loss_grad_kernel<<<grid, LT * LT, 0, s>>>(q);
        GSB_LAUNCH_CHECK();
    }
    return GSB_OK;
}


int gsb_tracking_loss(int width, int height, const float* color, const float* depth_sil, const float* median_depth,
                      const float* gt_color, const float* gt_depth, float w_image, float w_depth, int use_surdepth,
                      float* dL_dcolor, float* dL_ddepth_sil, float* loss_terms, gsb_stream_t stream)
{
    if (width <= 0 || height <= 0 || !color || !depth_sil || !gt_color || !gt_depth || !dL_dcolor || !loss_terms || (use_surdepth && !median_depth)) {
        set_error("tracking_loss: image size, color, depth_sil, gt_color, gt_depth, dL_dcolor, loss_terms (and median_depth with use_surdepth) are required");
        return GSB_ERR_INVALID_ARGUMENT;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int HW = width * height;
    GSB_CUDA_CHECK(cudaMemsetAsync(loss_terms, 0, 8 * sizeof(float), s));
    {
        StageTimer _t(ST_OTHER, s);
        const int blocks = (HW + TRK_THREADS - 1) / TRK_THREADS;
        tracking_loss_kernel<<<blocks < 2 * NUM_SMS ? blocks : 2 * NUM_SMS, TRK_THREADS, 0, s>>>(
            HW, color, depth_sil, median_depth, gt_color, gt_depth, w_image, w_depth, use_surdepth ? 1 : 0, dL_dcolor, dL_ddepth_sil, loss_terms);
        GSB_LAUNCH_CHECK();
    }
    return GSB_OK;
}

}  // extern "C"
