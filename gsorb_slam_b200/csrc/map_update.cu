// map_update.cu -- the per-Gaussian tail of one mapping iteration in ONE launch, sm_100a.
//
// Render::RenderForFrame's optimisation step (src/Render.cc:420-476) ends, per Gaussian, with four streaming passes in the
// reference and in this library's unfused path:
//   BACKWARD::preprocess  (backward.cu:560-621)          blend-backward sums -> gradients of the ACTIVATED parameters
//   autograd of the prologue (src/Render.cc:750-759)      -> gradients of the raw parameters (+ dL/dTcw)
//   the scale regularisers  (src/Render.cc:462-469)       += on the log-scale gradients
//   torch::optim::Adam      (src/Gaussian.cc:131-175)     parameters, exp_avg, exp_avg_sq
// Each of them is bound by HBM (56 B of gradients per Gaussian written and read back between every pair).  Here a thread owns one
// Gaussian from the accumulators to the updated parameters: 64 B accumulators + 16 B record + 4 B radius + 3 x 56 B parameter /
// moment rows in, 3 x 56 B out, nothing in between ever leaves the registers.  All loads are issued before the first use (the
// kernel is sized for bytes in flight, not for occupancy: 2 CTAs of 256 threads per SM).
// A forward that overflowed its binning blob blended truncated lists: the host renders that frame again with a larger blob
// (mapping.py), so the kernel must not apply an update from it -- it reads the overflow latch and leaves everything untouched.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include "common.cuh"
#include "gauss_bwd.cuh"
#include "map_math.cuh"

namespace gsb {

constexpr int MU_THREADS = 256;
constexpr int MU_GROUPS = 5;   // means 3 | rgb 3 | logit opacity 1 | log scales 3 | unnormalised quaternions 4
__host__ __device__ constexpr int mu_width(int g) { return g == 0 ? 3 : g == 1 ? 3 : g == 2 ? 1 : g == 3 ? 3 : 4; }
constexpr int MU_ROW = 14;
constexpr int MU_MINB = 3;        // resident CTAs per SM, scalar variant (80 registers)
constexpr int MU_BULK_MINB = 3;   // bulk variant (80 registers, 42 KB of shared memory per CTA; measured 93 / 82 / 87 / 108 us at 2 / 3 / 4 / 5)

struct MapUpdateParams {
    FwdParams f;
    const int* radii;
    const float* acc;
    const SplatRec* rec;
    const GeomHeader* hdr;
    const float* Tcw;
    float* par[MU_GROUPS];
    float* m[MU_GROUPS];
    float* v[MU_GROUPS];
    float* grad[MU_GROUPS];    // GRADS only
    float step_size[MU_GROUPS];
    float omb1, beta2, omb2, eps, inv_sqrt_bc2;
    float max_scalar, w_scalar, w_long;
    const float* reg_acc;      // {C, reg_scalar, sum(max - min)} from scale_reg_sum_kernel, or NULL: regularisers off
    float* reg_terms;
    float* dTcw;
    int z_attached;
};

__device__ __forceinline__ void report_reg_terms(const MapUpdateParams& q)   // what scale_reg_apply_kernel reports
{
    if (!q.reg_acc || !q.reg_terms) return;
    q.reg_terms[0] = q.reg_acc[1];
    q.reg_terms[1] = q.reg_acc[2] / q.reg_acc[0];
    q.reg_terms[2] = q.reg_acc[0];
    q.reg_terms[3] = 0.f;
}

// One row from the blend-backward sums and the raw parameters pv[14] to the raw-parameter gradients gr[14] (regularisers included)
// and the row's 12 terms of dL/dTcw.
__device__ __forceinline__ void map_row_gradients(const MapUpdateParams& q, size_t i, int radius, const float4& a0, const float4& a1,
                                                  const float4& a2, const float4& a3, const float4& rb, const float* pv, float* gr,
                                                  float* part)
{
    const FwdParams& p = q.f;
    const float reg_C = q.reg_acc ? q.reg_acc[0] : 0.f;
    // ---- the prologue's activations, recomputed (same device functions as prologue_kernel: same bits) ----
    const float wx = pv[0], wy = pv[1], wz = pv[2];
    const float mx = to_camera(q.Tcw, 0, wx, wy, wz), my = to_camera(q.Tcw, 1, wx, wy, wz), mz = to_camera(q.Tcw, 2, wx, wy, wz);
    const float sig = sigmoid_act(pv[6]);
    const float sc[3] = {expf(pv[7]), expf(pv[8]), expf(pv[9])};
    const float nrm = quat_norm(pv[10], pv[11], pv[12], pv[13]);
    // (one reciprocal instead of the prologue's four divisions: the rotation the gradient is evaluated at may differ from the
    // forward's by an ulp, which is far below the gradient tolerance; the forward itself never sees these values)
    const float inv_nrm = 1.0f / nrm;
    const float4 qn = make_float4(pv[10] * inv_nrm, pv[11] * inv_nrm, pv[12] * inv_nrm, pv[13] * inv_nrm);
    // ---- rasterizer backward of the row ----
    const bool rendered = radius > 0;
    float a[9];
    gauss_moments_to_2d(p, rendered, a0, a1, a2, a3, rb, a);
    float dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float dmx = 0.f, dmy = 0.f, dmz = 0.f;
    float dscale[3] = {0.f, 0.f, 0.f}, drot[4] = {0.f, 0.f, 0.f, 0.f};
    if (rendered) {
        gauss_backward_chain<false>(p, i, a, mx, my, mz, qn, sc[0], sc[1], sc[2], nullptr, 0u, dcov, dmx, dmy, dmz, dscale, drot);
        if (q.z_attached) dmz += a2.z;
    }
    // ---- chain rule of the prologue (prologue_backward_kernel) ----
    const float* T = q.Tcw;
#pragma unroll
    for (int c = 0; c < 3; c++) gr[c] = T[c] * dmx + T[4 + c] * dmy + T[8 + c] * dmz;
    part[0] = dmx * wx; part[1] = dmx * wy; part[2] = dmx * wz; part[3] = dmx;
    part[4] = dmy * wx; part[5] = dmy * wy; part[6] = dmy * wz; part[7] = dmy;
    part[8] = dmz * wx; part[9] = dmz * wy; part[10] = dmz * wz; part[11] = dmz;
    gr[3] = a[6]; gr[4] = a[7]; gr[5] = a[8];
    gr[6] = a[5] * sig * (1.0f - sig);
#pragma unroll
    for (int k = 0; k < 3; k++) gr[7 + k] = dscale[k] * sc[k];
    {
        const float dot = qn.x * drot[0] + qn.y * drot[1] + qn.z * drot[2] + qn.w * drot[3];
        gr[10] = (drot[0] - qn.x * dot) * inv_nrm;
        gr[11] = (drot[1] - qn.y * dot) * inv_nrm;
        gr[12] = (drot[2] - qn.z * dot) * inv_nrm;
        gr[13] = (drot[3] - qn.w * dot) * inv_nrm;
    }
    // ---- scale regularisers (scale_reg_apply_kernel) ----
    if (reg_C > 0.f) {
        const float wl = q.w_long / reg_C;
        const float n = (sc[0] > q.max_scalar ? 1.f : 0.f) + (sc[1] > q.max_scalar ? 1.f : 0.f) + (sc[2] > q.max_scalar ? 1.f : 0.f);
        if (n > 0.f) {
            // first maximum / first minimum, as torch.max / torch.min report them (selects: the arrays stay in registers)
            const bool max1 = sc[1] > sc[0], max2 = sc[2] > (max1 ? sc[1] : sc[0]);
            const bool min1 = sc[1] < sc[0], min2 = sc[2] < (min1 ? sc[1] : sc[0]);
            const int imax = max2 ? 2 : max1 ? 1 : 0, imin = min2 ? 2 : min1 ? 1 : 0;
            const float up = n * (q.w_scalar + wl), dn = n * wl;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                if (k == imax) gr[7 + k] += up * sc[k];
                if (k == imin) gr[7 + k] -= dn * sc[k];
            }
        }
    }
}

// Scalar variant (any alignment, any arena capacity): every load of the row is issued up front from the thread's own registers.
template <bool GRADS, int MINB>
__global__ void __launch_bounds__(MU_THREADS, MINB)
map_update_kernel(MapUpdateParams q)
{
    const FwdParams& p = q.f;
    if (q.hdr->overflow) return;   // uniform over the grid
    const int idx = blockIdx.x * MU_THREADS + threadIdx.x;
    float part[12];
#pragma unroll
    for (int k = 0; k < 12; k++) part[k] = 0.f;
    if (idx == 0) report_reg_terms(q);
    if (idx < p.P) {
        const size_t i = (size_t)idx;
        const int radius = q.radii[idx];
        const float4* ap = reinterpret_cast<const float4*>(q.acc + i * ACC_FLOATS);
        const float4 a0 = ap[0], a1 = ap[1], a2 = ap[2], a3 = ap[3];
        const float4 rb = q.rec[i].b;
        float pv[MU_ROW], mv[MU_ROW], vv[MU_ROW];
        {
            int k = 0;
#pragma unroll
            for (int g = 0; g < MU_GROUPS; g++)
#pragma unroll
                for (int c = 0; c < mu_width(g); c++, k++) {
                    const size_t e = (size_t)mu_width(g) * i + c;
                    pv[k] = q.par[g][e];
                    mv[k] = q.m[g][e];
                    vv[k] = q.v[g][e];
                }
        }
        float gr[MU_ROW];
        map_row_gradients(q, i, radius, a0, a1, a2, a3, rb, pv, gr, part);
        {
            int k = 0;
#pragma unroll
            for (int g = 0; g < MU_GROUPS; g++)
#pragma unroll
                for (int c = 0; c < mu_width(g); c++, k++) {
                    const size_t e = (size_t)mu_width(g) * i + c;
                    if (GRADS) q.grad[g][e] = gr[k];
                    adam_update_fast(pv[k], gr[k], mv[k], vv[k], q.step_size[g], q.omb1, q.beta2, q.omb2, q.eps, q.inv_sqrt_bc2);
                    q.m[g][e] = mv[k];
                    q.v[g][e] = vv[k];
                    q.par[g][e] = pv[k];
                }
        }
    }
    if (q.dTcw) reduce12_scatter_and_add<MU_THREADS>(part, q.dTcw);
}

// Bulk variant (every arena pointer 16-byte aligned): a CTA owns 256 consecutive rows, i.e. ONE contiguous chunk of each of the 15
// arrays (5 groups x {parameter, exp_avg, exp_avg_sq}).  The TMA unit's 1-D bulk copy (cp.async.bulk, UBLKCP) brings the 15 chunks
// into shared memory with full-line requests, one mbarrier transaction count for all of them, while the threads fetch their
// accumulators and records; every thread updates its row in place in shared memory and the 15 chunks go back by bulk stores.
// The per-thread form above moves the same bytes with 4-byte accesses of stride 12 / 16 (three requests per line and
// partially written sectors): 105 us against the 70 us the bytes cost.  The last, partial CTA copies its rows with plain loads.
__host__ __device__ constexpr int mu_group_off(int g) { return (g == 0 ? 0 : g == 1 ? 3 : g == 2 ? 6 : g == 3 ? 7 : 10) * MU_THREADS; }   // floats, inside one set
constexpr int MU_SET = MU_ROW * MU_THREADS;   // floats per set (parameters | exp_avg | exp_avg_sq)
static_assert(mu_group_off(4) + 4 * MU_THREADS == MU_SET, "set layout");

template <bool GRADS, int MINB>
__global__ void __launch_bounds__(MU_THREADS, MINB)
map_update_bulk_kernel(MapUpdateParams q)
{
    extern __shared__ __align__(128) float s_rows[];   // [3][MU_SET]
    __shared__ __align__(8) uint64_t s_bar;
    const FwdParams& p = q.f;
    if (q.hdr->overflow) return;   // uniform over the grid
    const int row0 = blockIdx.x * MU_THREADS, idx = row0 + threadIdx.x;
    const int n = min(MU_THREADS, p.P - row0);
    const bool full = n == MU_THREADS;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar);
    float* const* sets[3] = {q.par, q.m, q.v};
    if (full) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(3 * MU_SET * sizeof(float))) : "memory");
#pragma unroll
            for (int s = 0; s < 3; s++)
#pragma unroll
                for (int g = 0; g < MU_GROUPS; g++) {
                    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_rows + s * MU_SET + mu_group_off(g));
                    const float* src = sets[s][g] + (size_t)mu_width(g) * row0;
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(dst), "l"(src), "r"((uint32_t)(mu_width(g) * MU_THREADS * sizeof(float))), "r"(bar) : "memory");
                }
        }
    } else {
#pragma unroll
        for (int s = 0; s < 3; s++)
#pragma unroll
            for (int g = 0; g < MU_GROUPS; g++)
                for (int e = threadIdx.x; e < mu_width(g) * n; e += MU_THREADS)
                    s_rows[s * MU_SET + mu_group_off(g) + e] = sets[s][g][(size_t)mu_width(g) * row0 + e];
    }
    float part[12];
#pragma unroll
    for (int k = 0; k < 12; k++) part[k] = 0.f;
    if (idx == 0) report_reg_terms(q);
    const bool live = idx < p.P;
    // the thread's own loads travel while the bulk copies do
    int radius = 0;
    float4 a0, a1, a2, a3, rb;
    a0 = a1 = a2 = a3 = rb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
        radius = q.radii[idx];
        const float4* ap = reinterpret_cast<const float4*>(q.acc + (size_t)idx * ACC_FLOATS);
        a0 = ap[0]; a1 = ap[1]; a2 = ap[2]; a3 = ap[3];
        rb = q.rec[idx].b;
    }
    __syncthreads();   // the barrier is initialised (full) / the rows are in shared memory (partial)
    if (full) {
        uint32_t done;
        do {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(bar), "r"(0u) : "memory");
        } while (!done);
    }
    if (live) {
        const int t = threadIdx.x;
        float pv[MU_ROW], gr[MU_ROW];
        {
            int k = 0;
#pragma unroll
            for (int g = 0; g < MU_GROUPS; g++)
#pragma unroll
                for (int c = 0; c < mu_width(g); c++, k++) pv[k] = s_rows[mu_group_off(g) + mu_width(g) * t + c];
        }
        map_row_gradients(q, (size_t)idx, radius, a0, a1, a2, a3, rb, pv, gr, part);
        {
            int k = 0;
#pragma unroll
            for (int g = 0; g < MU_GROUPS; g++)
#pragma unroll
                for (int c = 0; c < mu_width(g); c++, k++) {
                    const int e = mu_group_off(g) + mu_width(g) * t + c;
                    if (GRADS) q.grad[g][(size_t)mu_width(g) * idx + c] = gr[k];
                    float mi = s_rows[MU_SET + e], vi = s_rows[2 * MU_SET + e];
                    adam_update_fast(pv[k], gr[k], mi, vi, q.step_size[g], q.omb1, q.beta2, q.omb2, q.eps, q.inv_sqrt_bc2);
                    s_rows[e] = pv[k];
                    s_rows[MU_SET + e] = mi;
                    s_rows[2 * MU_SET + e] = vi;
                }
        }
    }
    if (full) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the bulk stores
    __syncthreads();
    if (full) {
        if (threadIdx.x == 0) {
#pragma unroll
            for (int s = 0; s < 3; s++)
#pragma unroll
                for (int g = 0; g < MU_GROUPS; g++) {
                    const uint32_t src = (uint32_t)__cvta_generic_to_shared(s_rows + s * MU_SET + mu_group_off(g));
                    float* dst = sets[s][g] + (size_t)mu_width(g) * row0;
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                 ::"l"(dst), "r"(src), "r"((uint32_t)(mu_width(g) * MU_THREADS * sizeof(float))) : "memory");
                }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    } else {
#pragma unroll
        for (int s = 0; s < 3; s++)
#pragma unroll
            for (int g = 0; g < MU_GROUPS; g++)
                for (int e = threadIdx.x; e < mu_width(g) * n; e += MU_THREADS)
                    sets[s][g][(size_t)mu_width(g) * row0 + e] = s_rows[s * MU_SET + mu_group_off(g) + e];
    }
    if (q.dTcw) reduce12_scatter_and_add<MU_THREADS>(part, q.dTcw);
    if (full && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory stays alive until it is read
}

// ---- tracking: only the camera pose is optimised (Render::RenderStartTraking, src/Render.cc:1052-1127) -------------------------
// dL/dTcw[0:3,:] = sum_i dL/dmean_cam_i [mean_world_i; 1]^T needs the per-Gaussian backward only as far as dL/dmean_cam: no gradient
// array is written and the scale / rotation branch of the chain is dead code here.  Reads the activated arrays the forward saw.
struct PoseGradParams {
    FwdParams f;
    const int* radii;
    const float* acc;
    const SplatRec* rec;
    const uint32_t* tiles_touched;   // tile-row band only (else NULL): 0 = this band wrote no record for the Gaussian
    const float* means_world;
    float* dTcw;
    int z_attached;
};

__global__ void __launch_bounds__(MU_THREADS, 4)
pose_gradient_kernel(PoseGradParams q)
{
    const FwdParams& p = q.f;
    const int idx = blockIdx.x * MU_THREADS + threadIdx.x;
    float part[12];
#pragma unroll
    for (int k = 0; k < 12; k++) part[k] = 0.f;
    if (idx < p.P) {
        const size_t i = (size_t)idx;
        int radius = q.radii[idx];
        if (q.tiles_touched && q.tiles_touched[idx] == 0) radius = 0;
        const float4* ap = reinterpret_cast<const float4*>(q.acc + i * ACC_FLOATS);
        const float4 a0 = ap[0], a1 = ap[1], a2 = ap[2], a3 = ap[3];
        const float4 rb = q.rec[i].b;
        const float mx = p.means3D[3 * i], my = p.means3D[3 * i + 1], mz = p.means3D[3 * i + 2];
        const float4 qv = *reinterpret_cast<const float4*>(p.rotations + 4 * i);
        const float s0 = p.scales[3 * i], s1 = p.scales[3 * i + 1], s2 = p.scales[3 * i + 2];
        const float wx = q.means_world[3 * i], wy = q.means_world[3 * i + 1], wz = q.means_world[3 * i + 2];
        if (radius > 0) {
            float a[9], dcov[6], dscale[3], drot[4];
            float dmx = 0.f, dmy = 0.f, dmz = 0.f;
            gauss_moments_to_2d(p, true, a0, a1, a2, a3, rb, a);
            gauss_backward_chain<false>(p, i, a, mx, my, mz, qv, s0, s1, s2, nullptr, 0u, dcov, dmx, dmy, dmz, dscale, drot);
            if (q.z_attached) dmz += a2.z;
            part[0] = dmx * wx; part[1] = dmx * wy; part[2] = dmx * wz; part[3] = dmx;
            part[4] = dmy * wx; part[5] = dmy * wy; part[6] = dmy * wz; part[7] = dmy;
            part[8] = dmz * wx; part[9] = dmz * wy; part[10] = dmz * wz; part[11] = dmz;
        }
    }
    reduce12_scatter_and_add<MU_THREADS>(part, q.dTcw);
}

int launch_pose_gradient(const FwdParams& p, const char* geom, const GeomLayout& GL, const int* radii, int z_attached,
                         const float* means_world, float* dTcw, cudaStream_t s)
{
    GSB_CUDA_CHECK(cudaMemsetAsync(dTcw, 0, 12 * sizeof(float), s));
    if (p.P <= 0) return GSB_OK;
    PoseGradParams q;
    q.f = p;
    q.radii = radii ? radii : reinterpret_cast<const int*>(geom + GL.radii);
    q.acc = reinterpret_cast<const float*>(geom + GL.acc);
    q.rec = reinterpret_cast<const SplatRec*>(geom + GL.rec);
    q.tiles_touched = (p.band_y0 > 0 || p.band_y1 < p.tiles_y) ? reinterpret_cast<const uint32_t*>(geom + GL.tiles_touched) : nullptr;
    q.means_world = means_world;
    q.dTcw = dTcw;
    q.z_attached = z_attached;
    StageTimer _t(ST_GAUSS_BWD, s);
    pose_gradient_kernel<<<(p.P + MU_THREADS - 1) / MU_THREADS, MU_THREADS, 0, s>>>(q);
    GSB_LAUNCH_CHECK();
    return GSB_OK;
}

int launch_map_update(const FwdParams& p, const char* geom, const GeomLayout& GL, const int* radii, int z_attached,
                      const gsb_map_update& u, cudaStream_t s)
{
    if (u.dL_dTcw) GSB_CUDA_CHECK(cudaMemsetAsync(u.dL_dTcw, 0, 12 * sizeof(float), s));
    if (p.P <= 0) return GSB_OK;
    MapUpdateParams q;
    q.f = p;
    q.radii = radii ? radii : reinterpret_cast<const int*>(geom + GL.radii);
    q.acc = reinterpret_cast<const float*>(geom + GL.acc);
    q.rec = reinterpret_cast<const SplatRec*>(geom + GL.rec);
    q.hdr = reinterpret_cast<const GeomHeader*>(geom + GL.header);
    q.Tcw = u.Tcw;
    // torch keeps the hyper-parameters as doubles and rounds each derived scalar once (launch_adam_groups)
    const double bc1 = 1.0 - pow(u.beta1, (double)u.step), bc2 = 1.0 - pow(u.beta2, (double)u.step);
    bool grads = false;
    for (int g = 0; g < MU_GROUPS; g++) {
        q.par[g] = u.params[g];
        q.m[g] = u.exp_avg[g];
        q.v[g] = u.exp_avg_sq[g];
        q.grad[g] = u.grads[g];
        grads = grads || u.grads[g];
        q.step_size[g] = (float)((double)u.lr[g] / bc1);
    }
    q.omb1 = (float)(1.0 - u.beta1);
    q.beta2 = (float)u.beta2;
    q.omb2 = (float)(1.0 - u.beta2);
    q.eps = (float)u.eps;
    q.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    q.max_scalar = u.max_scalar; q.w_scalar = u.w_scalar; q.w_long = u.w_long;
    q.reg_acc = nullptr;
    q.reg_terms = u.reg_terms;
    if (u.max_scalar > 0.f) {   // the reduction pass of the regularisers; its apply pass is folded into the kernel below
        if (int rc = launch_scale_regulariser_sum(p.P, u.params[3], u.max_scalar, u.reg_terms + 4, s)) return rc;
        q.reg_acc = u.reg_terms + 4;
    }
    q.dTcw = u.dL_dTcw;
    q.z_attached = z_attached;
    StageTimer _t(ST_GAUSS_BWD, s);
    const int grid = (p.P + MU_THREADS - 1) / MU_THREADS;
    uintptr_t bits = 0;
    for (int g = 0; g < MU_GROUPS; g++)
        bits |= reinterpret_cast<uintptr_t>(q.par[g]) | reinterpret_cast<uintptr_t>(q.m[g]) | reinterpret_cast<uintptr_t>(q.v[g]);
    bool bulk = (bits & 15) == 0;
    const size_t smem = 3 * MU_SET * sizeof(float);
#ifdef GSB_TUNING
    static const char* mode = getenv("GSB_MAP_UPDATE");            // scalar: the per-thread variant whatever the alignment
    static const int minb = [] { const char* e = getenv("GSB_MAP_UPDATE_MINB"); return e ? atoi(e) : 0; }();
    if (mode && !strcmp(mode, "scalar")) bulk = false;
    if (!grads && bulk && minb == 2) map_update_bulk_kernel<false, 2><<<grid, MU_THREADS, smem, s>>>(q);
    else if (!grads && bulk && minb == 3) map_update_bulk_kernel<false, 3><<<grid, MU_THREADS, smem, s>>>(q);
    else if (!grads && bulk && minb == 5) map_update_bulk_kernel<false, 5><<<grid, MU_THREADS, smem, s>>>(q);
    else
#endif
    if (bulk && grads) map_update_bulk_kernel<true, MU_BULK_MINB><<<grid, MU_THREADS, smem, s>>>(q);
    else if (bulk) map_update_bulk_kernel<false, MU_BULK_MINB><<<grid, MU_THREADS, smem, s>>>(q);
    else if (grads) map_update_kernel<true, MU_MINB><<<grid, MU_THREADS, 0, s>>>(q);
    else map_update_kernel<false, MU_MINB><<<grid, MU_THREADS, 0, s>>>(q);
    GSB_LAUNCH_CHECK();
    return GSB_OK;
}

}  // namespace gsb
