// gauss_bwd.cu -- per-Gaussian backward (K8 + K9 fused) for sm_100a.
//
// Replaces BACKWARD::preprocess (backward.cu:560-621) = computeCov2DCUDA (:144-274) +
// preprocessCUDA<3> (:346-396) + computeCov3D backward (:278-341) + SH backward (:20-139).
// One launch instead of two; the 3D covariance is recomputed from scale/rotation instead of
// being stored by the forward pass (24 B/Gaussian less state), and every gradient array is
// fully written (zeros for Gaussians that were not rendered), so the caller does not have
// to zero-fill 108 B/Gaussian first (src/Rasterizer.cu:253-261).
#include <cstdlib>
#include "common.cuh"
#include "gauss_bwd.cuh"

namespace gsb {

constexpr int GB_THREADS = 256;

struct GaussBwdParams {
    FwdParams f;
    const int* radii;
    const uint32_t* tiles_touched;   // tiles of THIS band the Gaussian joined (preprocess): 0 = nothing to back-propagate
    const float* acc;          // [P][16] packed blend-backward sums (raw moments, GradAcc in common.cuh)
    const SplatRec* rec;       // conic + opacity for the moment -> gradient maps
    const uint8_t* clamped;    // SH clamp bits
    gsb_grad_outputs g;
    float* dL_dzcolor;         // [P] or NULL: gradient of the depth pass' z_cam colour (fused 5-channel pass)
    int z_attached;            // add that gradient to dL_dmean3D.z (the colour is a function of the mean: mapping mode)
};

template <int MINB, bool BAND>
__global__ void __launch_bounds__(GB_THREADS, MINB)
gauss_backward_kernel(GaussBwdParams q)
{
    const int idx = blockIdx.x * GB_THREADS + threadIdx.x;
    const FwdParams& p = q.f;
    if (idx >= p.P) return;
    const size_t i = (size_t)idx;
    const gsb_grad_outputs& g = q.g;
    // every load that depends only on idx is issued up front, whether or not the Gaussian was rendered: the
    // kernel is bound by memory latency, and a dependent chain radii -> accumulators -> parameters triples it
    // (tile-row shard: most Gaussians lie outside this rank's band -- there the dependent load pays: radius and band-local tile count
    // first, everything else only for the Gaussians the band rendered; their gradients are written as zeros)
    // BAND = false (the whole image): every load that depends only on idx is issued up front, as before
    int radius_ld = q.radii[idx];
    if (BAND && radius_ld > 0 && q.tiles_touched[idx] == 0) radius_ld = 0;
    const bool fetch = !BAND || radius_ld > 0;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* ap = reinterpret_cast<const float4*>(q.acc + i * ACC_FLOATS);
    const float4 a0 = fetch ? ap[0] : z4, a1 = fetch ? ap[1] : z4, a2 = fetch ? ap[2] : z4, a3 = fetch ? ap[3] : z4;
    const float4 rb = fetch ? q.rec[i].b : z4;  // conic.x, conic.y, conic.z, opacity (stale bytes when not rendered: never used then)
    float mx = 0.f, my = 0.f, mz = 0.f;
    float4 qv = make_float4(0.f, 0.f, 0.f, 0.f);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    if (fetch) {
        mx = p.means3D[3 * i]; my = p.means3D[3 * i + 1]; mz = p.means3D[3 * i + 2];
        if (!p.cov3D_precomp) {
            qv = *reinterpret_cast<const float4*>(p.rotations + 4 * i);
            s0 = p.scales[3 * i]; s1 = p.scales[3 * i + 1]; s2 = p.scales[3 * i + 2];
        }
    }
    const bool rendered = radius_ld > 0;
    float a[9];
    gauss_moments_to_2d(p, rendered, a0, a1, a2, a3, rb, a);
    if (g.dL_dmean2D) { g.dL_dmean2D[3 * i] = a[0]; g.dL_dmean2D[3 * i + 1] = a[1]; g.dL_dmean2D[3 * i + 2] = 0.f; }
    if (g.dL_dconic) { g.dL_dconic[4 * i] = a[2]; g.dL_dconic[4 * i + 1] = a[3]; g.dL_dconic[4 * i + 2] = 0.f; g.dL_dconic[4 * i + 3] = a[4]; }
    if (g.dL_dopacity) g.dL_dopacity[i] = a[5];
    if (q.dL_dzcolor) q.dL_dzcolor[i] = rendered ? a2.z : 0.f;
    if (g.dL_dcolor) { g.dL_dcolor[3 * i] = a[6]; g.dL_dcolor[3 * i + 1] = a[7]; g.dL_dcolor[3 * i + 2] = a[8]; }

    float dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float dmx = 0.f, dmy = 0.f, dmz = 0.f;
    float dscale[3] = {0.f, 0.f, 0.f}, drot[4] = {0.f, 0.f, 0.f, 0.f};
    if (rendered) {
        gauss_backward_chain<true>(p, i, a, mx, my, mz, qv, s0, s1, s2, g.dL_dsh, p.shs ? q.clamped[idx] : 0u, dcov, dmx, dmy, dmz, dscale, drot);
    } else if (p.shs && g.dL_dsh) {
        for (int k = 0; k < p.M * 3; k++) g.dL_dsh[i * p.M * 3 + k] = 0.f;
    }
    if (q.z_attached && rendered) dmz += a2.z;
    if (g.dL_dmean3D) { g.dL_dmean3D[3 * i] = dmx; g.dL_dmean3D[3 * i + 1] = dmy; g.dL_dmean3D[3 * i + 2] = dmz; }
    if (g.dL_dcov3D) {
#pragma unroll
        for (int k = 0; k < 6; k++) g.dL_dcov3D[6 * i + k] = dcov[k];
    }
    if (g.dL_dscale) { g.dL_dscale[3 * i] = dscale[0]; g.dL_dscale[3 * i + 1] = dscale[1]; g.dL_dscale[3 * i + 2] = dscale[2]; }
    if (g.dL_drot) { g.dL_drot[4 * i] = drot[0]; g.dL_drot[4 * i + 1] = drot[1]; g.dL_drot[4 * i + 2] = drot[2]; g.dL_drot[4 * i + 3] = drot[3]; }
}

int launch_gauss_backward(const FwdParams& p, const char* geom, const GeomLayout& GL, const int* radii,
                          const gsb_grad_outputs& g, float* dL_dzcolor, int z_attached, cudaStream_t s)
{
    if (p.P <= 0) return GSB_OK;
    GaussBwdParams q;
    q.f = p;
    q.radii = radii ? radii : reinterpret_cast<const int*>(geom + GL.radii);
    q.tiles_touched = reinterpret_cast<const uint32_t*>(geom + GL.tiles_touched);
    q.acc = reinterpret_cast<const float*>(geom + GL.acc);
    q.rec = reinterpret_cast<const SplatRec*>(geom + GL.rec);
    q.clamped = reinterpret_cast<const uint8_t*>(geom + GL.clamped);
    q.g = g;
    q.dL_dzcolor = dL_dzcolor;
    q.z_attached = z_attached;
    {
        StageTimer _t(ST_GAUSS_BWD, s);
        // 4 resident CTAs (64 registers, small spill) hide more memory latency than 3 (80 registers): measured, round 1
#ifdef GSB_TUNING
        static const int minb = [] { const char* e = getenv("GSB_GAUSS_BWD_MINB"); return e ? atoi(e) : 4; }();
        if (minb == 3) gauss_backward_kernel<3, false><<<(p.P + GB_THREADS - 1) / GB_THREADS, GB_THREADS, 0, s>>>(q);
        else if (minb == 5) gauss_backward_kernel<5, false><<<(p.P + GB_THREADS - 1) / GB_THREADS, GB_THREADS, 0, s>>>(q);
        else
#endif
        if (p.band_y0 > 0 || p.band_y1 < p.tiles_y) gauss_backward_kernel<4, true><<<(p.P + GB_THREADS - 1) / GB_THREADS, GB_THREADS, 0, s>>>(q);
        else gauss_backward_kernel<4, false><<<(p.P + GB_THREADS - 1) / GB_THREADS, GB_THREADS, 0, s>>>(q);
        GSB_LAUNCH_CHECK();
    }
    return GSB_OK;
}

}  // namespace gsb
