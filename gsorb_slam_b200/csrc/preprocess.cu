// preprocess.cu -- per-Gaussian projection + per-tile instance counting (K1), with the tile-count scan
// (K2, scan.cuh) run by the last CTA to finish; radii-only filter (K10) and frustum mask (K11); sm_100a.
//
// Replaces (reference file:line, behaviour only -- nothing is copied):
//   FORWARD::preprocess / preprocessCUDA      forward.cu:155-256
//   computeCov3D / computeCov2D               forward.cu:118-152 / :74-113
//   in_frustum, getRect, ndc2Pix              auxiliary.h:139-163, :46-56, :41-44
//   cub::DeviceScan::InclusiveSum + D2H copy  rasterizer_impl.cu:280-285
//   preprocessfilterCUDA / visible_filter     forward.cu:404-473, rasterizer_impl.cu:348-401
//   checkFrustum / markVisible                rasterizer_impl.cu:55-67, :142-154
//
// Layout: the reference scatters the projected state over seven SoA arrays (79 B/Gaussian);
// here one 48-byte SplatRec per Gaussian carries everything the blend kernels gather, and
// the AoS float3 inputs are staged through shared memory with 128-bit coalesced loads.
// Binning starts here: every (Gaussian, tile) instance is counted with an atomic whose return
// value is the instance's slot inside the tile's segment (kept for `duplicate`, binning.cu), and
// the tile-row band of a sharded frame (gsb_raster_args::tile_row_begin / _end) is applied.
#include <cstdlib>
#include "common.cuh"
#include "scan.cuh"

namespace gsb {

constexpr int PRE_THREADS = 256;

// Stage a block's [PRE_THREADS,3] slice of an AoS float3 array into shared memory with
// LDG.128 (3072 B per block = 192 float4), falling back to scalar loads when the base
// pointer is not 16-byte aligned or the slice is ragged.
__device__ __forceinline__ void stage_float3(const float* __restrict__ g, float* s, int P, int base, bool aligned)
{
    const int n = min(PRE_THREADS, P - base) * 3;  // floats in this slice
    const float* src = g + (size_t)base * 3;
    if (aligned && n == PRE_THREADS * 3) {
        if (threadIdx.x < PRE_THREADS * 3 / 4)
            reinterpret_cast<float4*>(s)[threadIdx.x] = __ldg(reinterpret_cast<const float4*>(src) + threadIdx.x);
    } else {
        for (int i = threadIdx.x; i < n; i += PRE_THREADS) s[i] = __ldg(src + i);
    }
}

struct Projected {
    bool visible;
    float depth, px, py, conx, cony, conz;
    float cov_xx, cov_yy, det;
    int radius;
    uint32_t minx, miny, maxx, maxy;
};

// forward.cu:155-256 up to the tile rectangle; op order as the reference compiles it.
__device__ __forceinline__ Projected project_gaussian(const FwdParams& p, float mx, float my, float mz,
                                                      const float* cov3D, int W, int H, int gx, int gy)
{
    Projected o;
    o.visible = false;
    o.radius = 0;
    const float* V = p.viewmatrix;
    const float* PM = p.projmatrix;
    const float pvz = xform_row(V, 2, mx, my, mz);
    if (!(pvz > 0.2f)) return o;  // auxiliary.h:154 (NaN also culls: "z <= 0.2" is false for NaN in the
                                  // reference, but such a Gaussian dies at det/radius; keep it out)
    const float hx = xform_row(PM, 0, mx, my, mz);
    const float hy = xform_row(PM, 1, mx, my, mz);
    const float hw = xform_row(PM, 3, mx, my, mz);
    const float p_w = __fdiv_rn(1.0f, __fadd_rn(hw, 0.0000001f));
    const float projx = __fmul_rn(hx, p_w), projy = __fmul_rn(hy, p_w);
    float cov[3];
    compute_cov2d(mx, my, mz, p.focal_x, p.focal_y, p.tan_fovx, p.tan_fovy, cov3D, V, cov, nullptr);
    const float det = fmaf(cov[0], cov[2], -__fmul_rn(cov[1], cov[1]));
    if (det == 0.0f) return o;
    const float det_inv = __fdiv_rn(1.f, det);
    o.conx = __fmul_rn(cov[2], det_inv);
    o.cony = __fmul_rn(cov[1], -det_inv);
    o.conz = __fmul_rn(cov[0], det_inv);
    const float mid = __fmul_rn(__fadd_rn(cov[0], cov[2]), 0.5f);
    const float disc = fmaxf(fmaf(mid, mid, -det), 0.1f);
    const float sq = __fsqrt_rn(disc);
    const float lambda1 = __fadd_rn(mid, sq), lambda2 = __fsub_rn(mid, sq);
    const float my_radius = ceilf(__fmul_rn(__fsqrt_rn(fmaxf(lambda1, lambda2)), 3.f));
    o.px = ndc2pix(projx, W);
    o.py = ndc2pix(projy, H);
    o.radius = (int)my_radius;
    get_rect(o.px, o.py, o.radius, gx, gy, o.minx, o.miny, o.maxx, o.maxy);
    if ((o.maxx - o.minx) * (o.maxy - o.miny) == 0) {
        o.radius = 0;
        return o;
    }
    o.depth = pvz;
    o.cov_xx = cov[0];
    o.cov_yy = cov[2];
    o.det = det;
    o.visible = true;
    return o;
}

// forward.cu:20-71 computeColorFromSH
__device__ __forceinline__ void sh_to_rgb(int deg, const float* __restrict__ sh, float mx, float my, float mz,
                                          const float* __restrict__ campos, float* rgb, uint32_t& clamped)
{
    float dx = mx - campos[0], dy = my - campos[1], dz = mz - campos[2];
    const float len = sqrtf(dx * dx + dy * dy + dz * dz);
    const float x = dx / len, y = dy / len, z = dz / len;
    clamped = 0;
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        auto S = [&](int k) { return sh[k * 3 + ch]; };
        float res = SH_C0 * S(0);
        if (deg > 0) {
            res = res - SH_C1 * y * S(1) + SH_C1 * z * S(2) - SH_C1 * x * S(3);
            if (deg > 1) {
                const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                res = res + SH_C2[0] * xy * S(4) + SH_C2[1] * yz * S(5) + SH_C2[2] * (2.0f * zz - xx - yy) * S(6) +
                      SH_C2[3] * xz * S(7) + SH_C2[4] * (xx - yy) * S(8);
                if (deg > 2) {
                    res = res + SH_C3[0] * y * (3.0f * xx - yy) * S(9) + SH_C3[1] * xy * z * S(10) +
                          SH_C3[2] * y * (4.0f * zz - xx - yy) * S(11) +
                          SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * S(12) +
                          SH_C3[4] * x * (4.0f * zz - xx - yy) * S(13) + SH_C3[5] * z * (xx - yy) * S(14) +
                          SH_C3[6] * x * (xx - 3.0f * yy) * S(15);
                }
            }
        }
        res += 0.5f;
        if (res < 0.f) clamped |= 1u << ch;
        rgb[ch] = fmaxf(res, 0.0f);
    }
}

// One thread per Gaussian, 256 per CTA.  Emits the packed record, radii, tiles_touched and
// the CTA's tile-count sum (first level of the two-level scan).
template <int MINB, bool BAND>
__global__ void __launch_bounds__(PRE_THREADS, MINB)
preprocess_kernel(FwdParams p, SplatRec* __restrict__ rec, int* __restrict__ radii_blob, int* __restrict__ radii_out,
                  uint32_t* __restrict__ tiles_touched, uint64_t* __restrict__ buckets, uint32_t* __restrict__ tile_count,
                  uint8_t* __restrict__ clamped_out, int aligned_means, int aligned_scales, int aligned_colors,
                  uint2* __restrict__ ranges, uint32_t* __restrict__ cursor, GeomHeader* __restrict__ hdr, uint32_t capacity)
{
    __shared__ __align__(16) float s_mean[PRE_THREADS * 3];
    __shared__ __align__(16) float s_scale[PRE_THREADS * 3];
    __shared__ __align__(16) float s_col[PRE_THREADS * 3];
    const int base = blockIdx.x * PRE_THREADS;
    const int idx = base + threadIdx.x;
    stage_float3(p.means3D, s_mean, p.P, base, aligned_means);
    if (p.scales) stage_float3(p.scales, s_scale, p.P, base, aligned_scales);
    if (p.colors_precomp) stage_float3(p.colors_precomp, s_col, p.P, base, aligned_colors);
    // per-Gaussian loads that depend only on idx are issued before the staging barrier (the kernel is bound by memory latency)
    float4 q_ld = make_float4(0.f, 0.f, 0.f, 1.f);
    float opacity_ld = 0.f;
    if (idx < p.P) {
        if (!p.cov3D_precomp) q_ld = __ldg(reinterpret_cast<const float4*>(p.rotations) + idx);
        opacity_ld = __ldg(p.opacities + idx);
    }
    __syncthreads();

    uint32_t touched = 0;
    if (idx < p.P) {
        const float mx = s_mean[3 * threadIdx.x], my = s_mean[3 * threadIdx.x + 1], mz = s_mean[3 * threadIdx.x + 2];
        float cov3D[6];
        if (p.cov3D_precomp) {
            const float2* c2 = reinterpret_cast<const float2*>(p.cov3D_precomp + 6 * (size_t)idx);
            const float2 a = __ldg(c2), b = __ldg(c2 + 1), c = __ldg(c2 + 2);
            cov3D[0] = a.x; cov3D[1] = a.y; cov3D[2] = b.x; cov3D[3] = b.y; cov3D[4] = c.x; cov3D[5] = c.y;
        } else {
            const float4 q = q_ld;
            compute_cov3d(s_scale[3 * threadIdx.x], s_scale[3 * threadIdx.x + 1], s_scale[3 * threadIdx.x + 2],
                          p.scale_modifier, q.x, q.y, q.z, q.w, cov3D);
        }
        const Projected o = project_gaussian(p, mx, my, mz, cov3D, p.W, p.H, p.tiles_x, p.tiles_y);
        int radius = 0;
        if (o.visible) {
            radius = o.radius;
            // tile-row shard: only the rows of this rank's band are binned (radii stay those of the whole image)
            const uint32_t by0 = max(o.miny, (uint32_t)p.band_y0), by1 = min(o.maxy, (uint32_t)p.band_y1);
            const uint32_t rows = by1 > by0 ? by1 - by0 : 0u;
            touched = rows * (o.maxx - o.minx);
        }
        // tile-row shard: a Gaussian that is visible but touches no tile of this rank's band keeps its radius and writes nothing
        // else -- no record, no colour, no threshold search (tiles_touched = 0 tells the per-Gaussian backward to skip it)
        if (o.visible && (!BAND || touched > 0)) {   // BAND = false (the whole image): a visible Gaussian always touches a tile
            const uint32_t by0 = max(o.miny, (uint32_t)p.band_y0), by1 = min(o.maxy, (uint32_t)p.band_y1);
            // Binning: one counting atomic per (Gaussian, tile) instance; the slot it returns is the record's place in the tile's
            // bucket, so the (depth bits << 32 | id) record is written right here and the tile sort picks it up -- no second pass over
            // the Gaussians.  Slots past the bucket's capacity are only counted: the scan sees the longest list and the per-tile
            // kernels fall back to the `duplicate` pass for such a frame (binning.cu).
            const uint64_t record = ((uint64_t)__float_as_uint(o.depth) << 32) | (uint32_t)idx;
            if (touched <= 4) {   // nearly all of them at SLAM splat sizes: the atomics are issued back to back, then the stores
                uint32_t slot[4], tl[4];
                uint32_t tx = o.minx, ty = by0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    slot[k] = 0xffffffffu;
                    tl[k] = ty * (uint32_t)p.tiles_x + tx;
                    if ((uint32_t)k < touched) {
                        slot[k] = atomicAdd(&tile_count[(size_t)tl[k] * TILE_CTR_STRIDE], 1u);
                        if (++tx == o.maxx) { tx = o.minx; ty++; }
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (slot[k] < (uint32_t)BUCKET_CAP) buckets[(size_t)tl[k] * BUCKET_CAP + slot[k]] = record;
            } else {
                for (uint32_t ty = by0; ty < by1; ty++)
                    for (uint32_t tx = o.minx; tx < o.maxx; tx++) {
                        const uint32_t tile = ty * (uint32_t)p.tiles_x + tx;
                        const uint32_t slot = atomicAdd(&tile_count[(size_t)tile * TILE_CTR_STRIDE], 1u);
                        if (slot < (uint32_t)BUCKET_CAP) buckets[(size_t)tile * BUCKET_CAP + slot] = record;
                    }
            }
            float rgb[3];
            uint32_t cl = 0;
            if (p.colors_precomp) {
                rgb[0] = s_col[3 * threadIdx.x]; rgb[1] = s_col[3 * threadIdx.x + 1]; rgb[2] = s_col[3 * threadIdx.x + 2];
            } else {
                sh_to_rgb(p.D, p.shs + (size_t)idx * p.M * 3, mx, my, mz, p.cam_pos, rgb, cl);
                clamped_out[idx] = (uint8_t)cl;
            }
            const float opacity = opacity_ld;
            // Cull data.  The footprint extents are conservative (they never change a result: the exact tests of forward.cu:346-356 are
            // still applied to everything that survives).  The power threshold is EXACT: thr = the smallest float p <= 0 with
            // fmul_rn(opacity, expf(p)) >= 1/255, i.e. the reference's `alpha < 1/255 -> continue` (forward.cu:357-359) restated in
            // power space with the very expf the blend kernels evaluate.  alpha(p) grows by >= 4 ulp per ulp of p near the threshold
            // and expf is good to 2 ulp, so the boundary is found by stepping a few ulps from -ln(255 * opacity); the four values
            // below it are checked as well, and should one of them pass (expf not monotonic there) the threshold moves down to it.
            // The forward kernel keeps the alpha test (thr only saves it the exponential); the backward kernel decides by thr alone.
            const float lim = logf(255.0f * opacity);             // > 0 iff the splat can ever contribute
            float thr, ex, ey;
            if (!(lim > 0.f)) {
                thr = 1.0f;  // power <= 0 < thr always: never contributes
                ex = ey = 0.f;
            } else {
                const float t2 = 2.0f * (lim * 1.002f + 0.01f);   // 2 * (-conservative threshold)
                auto passes = [&](uint32_t bits) { return !(__fmul_rn(opacity, expf(__uint_as_float(bits))) < 1.0f / 255.0f); };
                uint32_t pb = __float_as_uint(-lim);              // negative floats: bits + 1 = one ulp further from zero
                int steps = 0;
                if (passes(pb)) {
                    while (steps++ < 24 && passes(pb + 1)) pb++;
                } else {
                    do pb--; while (steps++ < 24 && !passes(pb) && (pb << 1) != 0u);
                }
                for (uint32_t k = 2; k <= 5 && steps <= 24; k++)
                    if (passes(pb + k)) { pb += k; k = 1; steps++; }
                if (steps > 24) {
                    // the ulp walk did not settle (a threshold within a few 1e-7 of zero: opacity barely above 1/255): bisect the bit
                    // patterns between -0.0 (passes: lim > 0) and the conservative bound (fails) -- the backward kernel has no alpha test
                    // to fall back on, so the threshold must be exact here too
                    uint32_t lo = 0x80000000u, hi = __float_as_uint(-0.5f * t2);
                    if (!passes(lo)) lo = hi = __float_as_uint(1.0f);          // cannot contribute at all
                    else if (passes(hi)) lo = hi;                               // (not expected) keep the conservative bound
                    while (hi - lo > 1u) {
                        const uint32_t mid = lo + ((hi - lo) >> 1);
                        if (passes(mid)) lo = mid; else hi = mid;
                    }
                    pb = lo;
                }
                thr = __uint_as_float(pb);
                // bbox of {d : d^T Q d <= t2} is sqrt(t2 * (Q^-1)_ii); Q^-1 = cov2D up to fp32 rounding
                // of the conic, hence the 2 % + 0.05 px slack.
                ex = sqrtf(t2 * o.cov_xx) * 1.02f + 0.05f;
                ey = sqrtf(t2 * o.cov_yy) * 1.02f + 0.05f;
                // No culling for anything numerically suspicious: a non-positive-definite 2D
                // covariance (possible with caller-supplied cov3D) has an unbounded footprint, and
                // for splats with sigma > 100 px in both axes the fp32 determinant (hence the conic)
                // can be off by more than the slack.
                const bool sane = o.det > 0.f && o.cov_xx > 0.f && o.cov_yy > 0.f && fminf(o.cov_xx, o.cov_yy) < 1.0e4f;
                if (!sane || !(ex < 60000.f)) ex = 60000.f;  // also catches NaN
                if (!sane || !(ey < 60000.f)) ey = 60000.f;
            }
            const __half2 ext = __halves2half2(__float2half_ru(ex), __float2half_ru(ey));
            SplatRec r;
            r.a = make_float4(o.px, o.py, __uint_as_float(*reinterpret_cast<const uint32_t*>(&ext)), thr);
            r.b = make_float4(o.conx, o.cony, o.conz, opacity);
            r.c = make_float4(rgb[0], rgb[1], rgb[2], o.depth);
            rec[idx] = r;
        }
        radii_blob[idx] = radius;
        if (radii_out) radii_out[idx] = radius;
        tiles_touched[idx] = touched;
    }
    // K2 without a launch: the last CTA to get here turns the tile counts into segments (scan.cuh).  The completion counter
    // lives one slot past the tile counters and is zeroed with them.
    __shared__ bool s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        // No fence here: every counting atomic of this CTA has RETURNED its slot to the issuing thread (the bucket store's address
        // depends on it), i.e. it has been performed at the L2, before that thread reached the barrier above -- and the counts are
        // all the scan reads.  (A __threadfence() here also waited for the CTA's scattered bucket stores: 15 us per frame.)
        const int tiles = p.tiles_x * p.tiles_y;
        s_last = atomicAdd(&tile_count[(size_t)tiles * TILE_CTR_STRIDE], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        tile_scan_body<PRE_THREADS, 6>(tile_count, ranges, cursor, p.tiles_x * p.tiles_y, hdr, capacity, p.P);
    }
}

int launch_preprocess(const FwdParams& p, char* geom, const GeomLayout& GL, char* image, const ImageLayout& IL,
                      int* radii_out, uint32_t capacity, cudaStream_t s)
{
    if (p.P <= 0) return GSB_OK;
    auto al = [](const void* q) { return q && (reinterpret_cast<uintptr_t>(q) & 15) == 0 ? 1 : 0; };
    {
        StageTimer _t(ST_PREPROCESS, s);
        // resident CTAs per SM the compiler must allow (the kernel is latency bound: occupancy against registers).  Measured with the
        // bucket stores in the kernel: 8 (32 registers, 44 B of spills) 67.6 us, 6 (40 registers) 64.0 us, 5 (48) 65.9 us.  (Deferring
        // the bucket stores to the end of the thread, under the atomics' round trip, costs registers: 66.7 us at 6.)
#ifdef GSB_TUNING
        static const int minb = [] { const char* e = getenv("GSB_PREPROCESS_MINB"); return e ? atoi(e) : 6; }();
#endif
#define GSB_PRE_LAUNCH_BD(MB, BD)                                                                                         \
    preprocess_kernel<MB, BD><<<GL.num_blocks, PRE_THREADS, 0, s>>>(                                                          \
        p, reinterpret_cast<SplatRec*>(geom + GL.rec), reinterpret_cast<int*>(geom + GL.radii), radii_out,                \
        reinterpret_cast<uint32_t*>(geom + GL.tiles_touched), reinterpret_cast<uint64_t*>(image + IL.buckets),             \
        reinterpret_cast<uint32_t*>(image + IL.tile_count), reinterpret_cast<uint8_t*>(geom + GL.clamped),                \
        al(p.means3D), al(p.scales), al(p.colors_precomp), reinterpret_cast<uint2*>(image + IL.ranges),                 \
        reinterpret_cast<uint32_t*>(image + IL.tile_cursor), reinterpret_cast<GeomHeader*>(geom + GL.header), capacity)
        const bool band_mode = p.band_y0 > 0 || p.band_y1 < p.tiles_y;   // tile-row shard: Gaussians outside the band write no record
#define GSB_PRE_LAUNCH(MB) do { if (band_mode) GSB_PRE_LAUNCH_BD(MB, true); else GSB_PRE_LAUNCH_BD(MB, false); } while (0)
#ifdef GSB_TUNING
        if (minb == 8) GSB_PRE_LAUNCH(8); else if (minb == 5) GSB_PRE_LAUNCH(5); else
#endif
        GSB_PRE_LAUNCH(6);
#undef GSB_PRE_LAUNCH
#undef GSB_PRE_LAUNCH_BD
        GSB_LAUNCH_CHECK();
    }
    return GSB_OK;
}

// ---- visible_filter: radii only (forward.cu:404-473) ---------------------------------------
__global__ void __launch_bounds__(PRE_THREADS)
visible_filter_kernel(FwdParams p, int* __restrict__ radii)
{
    const int idx = blockIdx.x * PRE_THREADS + threadIdx.x;
    if (idx >= p.P) return;
    const float mx = __ldg(p.means3D + 3 * (size_t)idx), my = __ldg(p.means3D + 3 * (size_t)idx + 1),
                mz = __ldg(p.means3D + 3 * (size_t)idx + 2);
    float cov3D[6];
    if (p.cov3D_precomp) {
#pragma unroll
        for (int k = 0; k < 6; k++) cov3D[k] = __ldg(p.cov3D_precomp + 6 * (size_t)idx + k);
    } else {
        const float4 q = __ldg(reinterpret_cast<const float4*>(p.rotations) + idx);
        compute_cov3d(__ldg(p.scales + 3 * (size_t)idx), __ldg(p.scales + 3 * (size_t)idx + 1),
                      __ldg(p.scales + 3 * (size_t)idx + 2), p.scale_modifier, q.x, q.y, q.z, q.w, cov3D);
    }
    const Projected o = project_gaussian(p, mx, my, mz, cov3D, p.W, p.H, p.tiles_x, p.tiles_y);
    radii[idx] = o.visible ? o.radius : 0;
}

int launch_visible_filter(const FwdParams& p, int* radii, cudaStream_t s)
{
    if (p.P <= 0) return GSB_OK;
    {
        StageTimer _t(ST_OTHER, s);
        visible_filter_kernel<<<(p.P + PRE_THREADS - 1) / PRE_THREADS, PRE_THREADS, 0, s>>>(p, radii);
        GSB_LAUNCH_CHECK();
    }
    return GSB_OK;
}

// ---- mark_visible (rasterizer_impl.cu:55-67) ------------------------------------------------
__global__ void __launch_bounds__(PRE_THREADS)
mark_visible_kernel(int P, const float* __restrict__ means3D, const float* __restrict__ view, uint8_t* __restrict__ present)
{
    const int idx = blockIdx.x * PRE_THREADS + threadIdx.x;
    if (idx >= P) return;
    const float z = xform_row(view, 2, __ldg(means3D + 3 * (size_t)idx), __ldg(means3D + 3 * (size_t)idx + 1),
                              __ldg(means3D + 3 * (size_t)idx + 2));
    present[idx] = z > 0.2f ? 1 : 0;
}

int launch_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present, cudaStream_t s)
{
    if (P <= 0) return GSB_OK;
    {
        StageTimer _t(ST_OTHER, s);
        mark_visible_kernel<<<(P + PRE_THREADS - 1) / PRE_THREADS, PRE_THREADS, 0, s>>>(P, means3D, viewmatrix, present);
        GSB_LAUNCH_CHECK();
    }
    return GSB_OK;
}

}  // namespace gsb
