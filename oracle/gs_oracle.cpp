// oracle/gs_oracle.cpp -- TEST INFRASTRUCTURE (the parity checker), NOT PRODUCT CODE.
//
// A plain-C++ CPU restatement of the reference's differentiable Gaussian rasterizer
// (GSORB-SLAM, Thirdparty/diff_gaussian_rasterization/cuda_rasterizer) and of its
// simple_knn.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
// load this library; the product (gsorb_slam_b200/csrc) never links or calls it.
//
// Parity status: PINNED against outputs of the reference's own CUDA kernels
// (oracle/_ref/libgsref.so, built in place from /root/reference and executed on a B200;
// fixtures under tests/golden/, generator tests/golden/make_golden.py).  The reference
// ships no golden vectors of its own.
//
// Every function cites the reference file:line it restates.  Abbreviations:
//   fwd.cu  = Thirdparty/diff_gaussian_rasterization/cuda_rasterizer/forward.cu
//   bwd.cu  = .../backward.cu      impl.cu = .../rasterizer_impl.cu     aux.h = .../auxiliary.h
//
// Floating-point contract.  The forward geometry (projection, covariance, conic, radius,
// tile rectangle) feeds INTEGER decisions (ceil, trunc, ==0), so it is restated with the
// exact fp32 operation order -- including which products are fused into FMAs -- that
// nvcc 12.9 emits for the reference sources (read from the SASS of oracle/_ref).  This
// file is compiled with -ffp-contract=off; every fused multiply-add below is an explicit
// fmaf().  IEEE division / sqrt / reciprocal are correctly rounded on both sides, so those
// quantities are expected to be BIT-IDENTICAL to the reference kernels.  expf() differs
// between libm and CUDA (MUFU.EX2 based), so blended colours / gradients carry a tolerance.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr int BLOCK_X = 16;  // config.h:16
constexpr int BLOCK_Y = 16;  // config.h:17
constexpr int NCH = 3;       // config.h:15

// nvcc's contraction of  a0*b0 + a1*b1 + a2*b2  (left-assoc):  the MIDDLE product is a
// plain multiply, the first is fused onto it, the third fused last (seen throughout the
// reference SASS: transformPoint4x*, glm mat3*mat3).
inline float nv3(float a0, float b0, float a1, float b1, float a2, float b2)
{
    return fmaf(a2, b2, fmaf(a0, b0, a1 * b1));
}

// aux.h:58-66 transformPoint4x3 / :68-77 transformPoint4x4 (matrix read column-major).
inline float xform_row(const float* m, int r, float x, float y, float z)
{
    return nv3(m[r], x, m[4 + r], y, m[8 + r], z) + m[12 + r];
}

// aux.h:41-44 ndc2Pix -- evaluated in double (the literals 1.0/0.5 are doubles), one rounding.
inline float ndc2pix(float v, int S)
{
    return (float)(std::fma((double)v + 1.0, (double)S, -1.0) * 0.5);  // DADD, DFMA, DMUL, F2F
}

// aux.h:46-56 getRect
inline void get_rect(float px, float py, int max_radius, int gx, int gy,
                     uint32_t& minx, uint32_t& miny, uint32_t& maxx, uint32_t& maxy)
{
    const float r = (float)max_radius;
    auto lo = [&](float p, int g) {
        int v = (int)((p - r) * 0.0625f);  // "/ BLOCK_X" with BLOCK_X == 16 (exact)
        return (uint32_t)std::min<uint32_t>((uint32_t)g, (uint32_t)std::max(0, v));
    };
    auto hi = [&](float p, int g) {
        float t = p + r;
        t = t + 16.0f;  // + BLOCK_X
        t = t - 1.0f;   // - 1
        int v = (int)(t * 0.0625f);
        return (uint32_t)std::min<uint32_t>((uint32_t)g, (uint32_t)std::max(0, v));
    };
    static_assert(BLOCK_X == 16 && BLOCK_Y == 16, "rect arithmetic assumes 16x16 tiles");
    minx = lo(px, gx);
    miny = lo(py, gy);
    maxx = hi(px, gx);
    maxy = hi(py, gy);
}

// float -> int truncation with CUDA F2I semantics for out-of-range / NaN inputs is not
// reproduced; inputs in the tests stay in range.

// fwd.cu:118-152 computeCov3D (quaternion used as given, NOT normalised: fwd.cu:127)
inline void compute_cov3d(const float* s3, float mod, const float* q4, float* cov3D)
{
    const float r = q4[0], x = q4[1], y = q4[2], z = q4[3];
    const float sx = mod * s3[0], sy = mod * s3[1], sz = mod * s3[2];
    // rotation entries, contraction pattern as emitted for fwd.cu:134-138
    const float xz = x * z, rx = r * x, rz = r * z, yy = y * y, zz = z * z;
    const float A = fmaf(r, y, xz);     // xz + ry
    const float B = fmaf(-r, y, xz);    // xz - ry
    const float C = fmaf(y, z, -rx);    // yz - rx
    const float Dd = fmaf(y, z, rx);    // yz + rx
    const float E = fmaf(x, y, -rz);    // xy - rz
    const float F = fmaf(x, y, rz);     // xy + rz
    const float G = fmaf(x, x, yy);     // xx + yy
    const float Hh = yy + zz;           // yy + zz  (two plain products, plain add)
    const float I = fmaf(x, x, zz);     // xx + zz
    // glm column-major R[c][r]
    const float R00 = 1.f - (Hh + Hh), R01 = E + E, R02 = A + A;
    const float R10 = F + F, R11 = 1.f - (I + I), R12 = C + C;
    const float R20 = B + B, R21 = Dd + Dd, R22 = 1.f - (G + G);
    // M = S * R  (fwd.cu:140): M[c][r] = s_r * R[c][r]
    const float M00 = sx * R00, M01 = sy * R01, M02 = sz * R02;
    const float M10 = sx * R10, M11 = sy * R11, M12 = sz * R12;
    const float M20 = sx * R20, M21 = sy * R21, M22 = sz * R22;
    // Sigma = transpose(M) * M (fwd.cu:143): Sigma[c][r] = sum_k M[r][k] * M[c][k]
    cov3D[0] = nv3(M00, M00, M01, M01, M02, M02);
    cov3D[1] = nv3(M00, M10, M01, M11, M02, M12);
    cov3D[2] = nv3(M00, M20, M01, M21, M02, M22);
    cov3D[3] = nv3(M10, M10, M11, M11, M12, M12);
    cov3D[4] = nv3(M10, M20, M11, M21, M12, M22);
    cov3D[5] = nv3(M20, M20, M21, M21, M22, M22);
}

struct Cov2DInter {
    float tx, ty, tz;        // clamped t (fwd.cu:80-87)
    float txtz, tytz;        // unclamped ratios
    float T00, T01, T02, T10, T11, T12;  // T = W*J rows (glm T[0][r], T[1][r])
};

// fwd.cu:74-113 computeCov2D.  Returns (cov.x, cov.y, cov.z) INCLUDING the +0.3 low-pass.
inline void compute_cov2d(const float* mean, float focal_x, float focal_y, float tan_fovx,
                          float tan_fovy, const float* c, const float* v, float* cov,
                          Cov2DInter* inter)
{
    const float x = mean[0], y = mean[1], z = mean[2];
    float tx = xform_row(v, 0, x, y, z);
    float ty = xform_row(v, 1, x, y, z);
    const float tz = xform_row(v, 2, x, y, z);
    const float limx = tan_fovx * 1.3f, limy = tan_fovy * 1.3f;
    const float txtz = tx / tz, tytz = ty / tz;
    const float cx = std::min(std::max(txtz, -limx), limx);
    const float cy = std::min(std::max(tytz, -limy), limy);
    tx = cx * tz;
    ty = cy * tz;
    const float tz2 = tz * tz;
    const float J00 = focal_x / tz;
    const float J02 = ((-tx) * focal_x) / tz2;  // -(focal_x * t.x) / (t.z * t.z)
    const float J11 = focal_y / tz;
    const float J12 = ((-ty) * focal_y) / tz2;
    // T = W * J (fwd.cu:99); W[0]=(v0,v4,v8) W[1]=(v1,v5,v9) W[2]=(v2,v6,v10)
    const float W0[3] = {v[0], v[4], v[8]}, W1[3] = {v[1], v[5], v[9]}, W2[3] = {v[2], v[6], v[10]};
    float T0[3], T1[3];
    for (int r = 0; r < 3; r++) {
        T0[r] = fmaf(W2[r], J02, W0[r] * J00);
        T1[r] = fmaf(W2[r], J12, W1[r] * J11);
    }
    // A = transpose(T) * transpose(Vrk): A[c][r] = T[r][0]*Vrk[0][c] + T[r][1]*Vrk[1][c] + T[r][2]*Vrk[2][c]
    const float V0[3] = {c[0], c[1], c[2]}, V1[3] = {c[1], c[3], c[4]}, V2[3] = {c[2], c[4], c[5]};
    float A0[3], A1[3];  // A[c][0], A[c][1]
    for (int cc = 0; cc < 3; cc++) {
        A0[cc] = nv3(T0[0], V0[cc], T0[1], V1[cc], T0[2], V2[cc]);
        A1[cc] = nv3(T1[0], V0[cc], T1[1], V1[cc], T1[2], V2[cc]);
    }
    // cov = A * T: cov[c][r] = A[0][r]*T[c][0] + A[1][r]*T[c][1] + A[2][r]*T[c][2]
    const float c00 = nv3(A0[0], T0[0], A0[1], T0[1], A0[2], T0[2]);
    const float c01 = nv3(A1[0], T0[0], A1[1], T0[1], A1[2], T0[2]);
    const float c11 = nv3(A1[0], T1[0], A1[1], T1[1], A1[2], T1[2]);
    cov[0] = c00 + 0.3f;
    cov[1] = c01;
    cov[2] = c11 + 0.3f;
    if (inter) {
        inter->tx = tx; inter->ty = ty; inter->tz = tz; inter->txtz = txtz; inter->tytz = tytz;
        inter->T00 = T0[0]; inter->T01 = T0[1]; inter->T02 = T0[2];
        inter->T10 = T1[0]; inter->T11 = T1[1]; inter->T12 = T1[2];
    }
}

// SH basis constants, aux.h:20-38
const float SH_C0 = 0.28209479177387814f;
const float SH_C1 = 0.4886025119029199f;
const float SH_C2[] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                       -1.0925484305920792f, 0.5462742152960396f};
const float SH_C3[] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                       0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                       -0.5900435899266435f};

// fwd.cu:20-71 computeColorFromSH (tolerance path: not restated op-for-op)
inline void sh_to_rgb(int deg, int M, const float* mean, const float* campos, const float* sh,
                      float* rgb, uint8_t* clamped)
{
    float dx = mean[0] - campos[0], dy = mean[1] - campos[1], dz = mean[2] - campos[2];
    const float len = std::sqrt(dx * dx + dy * dy + dz * dz);
    const float x = dx / len, y = dy / len, z = dz / len;
    for (int ch = 0; ch < 3; ch++) {
        auto S = [&](int k) { return sh[k * 3 + ch]; };
        float res = SH_C0 * S(0);
        if (deg > 0) {
            res = res - SH_C1 * y * S(1) + SH_C1 * z * S(2) - SH_C1 * x * S(3);
            if (deg > 1) {
                const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                res = res + SH_C2[0] * xy * S(4) + SH_C2[1] * yz * S(5) +
                      SH_C2[2] * (2.0f * zz - xx - yy) * S(6) + SH_C2[3] * xz * S(7) +
                      SH_C2[4] * (xx - yy) * S(8);
                if (deg > 2) {
                    res = res + SH_C3[0] * y * (3.0f * xx - yy) * S(9) + SH_C3[1] * xy * z * S(10) +
                          SH_C3[2] * y * (4.0f * zz - xx - yy) * S(11) +
                          SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * S(12) +
                          SH_C3[4] * x * (4.0f * zz - xx - yy) * S(13) +
                          SH_C3[5] * z * (xx - yy) * S(14) + SH_C3[6] * x * (xx - 3.0f * yy) * S(15);
                }
            }
        }
        res += 0.5f;
        clamped[ch] = res < 0;
        rgb[ch] = std::max(res, 0.0f);
    }
    (void)M;
}

struct State {
    int P = 0, W = 0, H = 0, gx = 0, gy = 0, D = 0, M = 0;
    bool colors_precomp = false, cov_precomp = false;
    std::vector<float> depths, means2D, cov3D, conic_opacity, rgb;
    std::vector<uint8_t> clamped;
    std::vector<int> radii;
    std::vector<uint32_t> tiles_touched, point_offsets;
    std::vector<uint64_t> keys;          // sorted
    std::vector<uint32_t> point_list;    // sorted
    std::vector<uint32_t> ranges;        // 2 per tile
    std::vector<float> final_T;
    std::vector<uint32_t> n_contrib;
    long long num_rendered = 0;
};

}  // namespace

extern "C" {

// ---------------------------------------------------------------------------------------
// Forward: impl.cu:199-345 (orchestration), fwd.cu:155-256 (preprocess), impl.cu:71-139
// (keys + ranges), fwd.cu:261-401 (blend).
// Returns an opaque state handle (the analogue of geom/binning/img buffers).
// ---------------------------------------------------------------------------------------
void* gso_forward(int P, int D, int M, const float* background, int width, int height,
                  const float* means3D, const float* shs, const float* colors_precomp,
                  const float* opacities, const float* scales, float scale_modifier,
                  const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                  const float* projmatrix, const float* cam_pos, float tan_fovx, float tan_fovy,
                  float* out_color, float* out_depth, int* radii_out, int* num_rendered_out)
{
    State* st = new State();
    st->P = P; st->W = width; st->H = height; st->D = D; st->M = M;
    st->gx = (width + BLOCK_X - 1) / BLOCK_X;
    st->gy = (height + BLOCK_Y - 1) / BLOCK_Y;
    st->colors_precomp = colors_precomp != nullptr;
    st->cov_precomp = cov3D_precomp != nullptr;
    const float focal_y = height / (2.0f * tan_fovy);  // impl.cu:224-225
    const float focal_x = width / (2.0f * tan_fovx);
    st->depths.assign(P, 0.f); st->means2D.assign(2 * (size_t)P, 0.f);
    st->cov3D.assign(6 * (size_t)P, 0.f); st->conic_opacity.assign(4 * (size_t)P, 0.f);
    st->rgb.assign(3 * (size_t)P, 0.f); st->clamped.assign(3 * (size_t)P, 0);
    st->radii.assign(P, 0); st->tiles_touched.assign(P, 0); st->point_offsets.assign(P, 0);
    const int gx = st->gx, gy = st->gy;

    // ---- K1 preprocess (fwd.cu:155-256) ----
#pragma omp parallel for schedule(static)
    for (int idx = 0; idx < P; idx++) {
        const float* p = means3D + 3 * (size_t)idx;
        // in_frustum (aux.h:139-163): only the z <= 0.2 test culls
        const float pvz = xform_row(viewmatrix, 2, p[0], p[1], p[2]);
        if (pvz <= 0.2f) continue;
        const float hx = xform_row(projmatrix, 0, p[0], p[1], p[2]);
        const float hy = xform_row(projmatrix, 1, p[0], p[1], p[2]);
        const float hw = xform_row(projmatrix, 3, p[0], p[1], p[2]);
        const float p_w = 1.0f / (hw + 0.0000001f);
        const float projx = hx * p_w, projy = hy * p_w;
        const float* cov3D;
        if (cov3D_precomp) {
            cov3D = cov3D_precomp + 6 * (size_t)idx;
        } else {
            compute_cov3d(scales + 3 * (size_t)idx, scale_modifier, rotations + 4 * (size_t)idx,
                          &st->cov3D[6 * (size_t)idx]);
            cov3D = &st->cov3D[6 * (size_t)idx];
        }
        float cov[3];
        compute_cov2d(p, focal_x, focal_y, tan_fovx, tan_fovy, cov3D, viewmatrix, cov, nullptr);
        const float det = fmaf(cov[0], cov[2], -(cov[1] * cov[1]));  // fwd.cu:219
        if (det == 0.0f) continue;
        const float det_inv = 1.f / det;
        const float conic[3] = {cov[2] * det_inv, cov[1] * (-det_inv), cov[0] * det_inv};
        const float mid = (cov[0] + cov[2]) * 0.5f;                 // fwd.cu:229
        const float disc = std::max(fmaf(mid, mid, -det), 0.1f);
        const float sq = std::sqrt(disc);
        const float lambda1 = mid + sq, lambda2 = mid - sq;
        const float my_radius = std::ceil(std::sqrt(std::max(lambda1, lambda2)) * 3.f);
        const float px = ndc2pix(projx, width), py = ndc2pix(projy, height);
        uint32_t minx, miny, maxx, maxy;
        get_rect(px, py, (int)my_radius, gx, gy, minx, miny, maxx, maxy);
        if ((maxx - minx) * (maxy - miny) == 0) continue;
        if (!colors_precomp)
            sh_to_rgb(D, M, p, cam_pos, shs + (size_t)idx * M * 3, &st->rgb[3 * (size_t)idx],
                      &st->clamped[3 * (size_t)idx]);
        st->depths[idx] = pvz;
        st->radii[idx] = (int)my_radius;
        st->means2D[2 * (size_t)idx] = px;
        st->means2D[2 * (size_t)idx + 1] = py;
        st->conic_opacity[4 * (size_t)idx + 0] = conic[0];
        st->conic_opacity[4 * (size_t)idx + 1] = conic[1];
        st->conic_opacity[4 * (size_t)idx + 2] = conic[2];
        st->conic_opacity[4 * (size_t)idx + 3] = opacities[idx];
        st->tiles_touched[idx] = (maxy - miny) * (maxx - minx);
    }
    if (radii_out) std::memcpy(radii_out, st->radii.data(), sizeof(int) * (size_t)P);

    // ---- K2 inclusive scan (impl.cu:280-285) ----
    uint64_t acc = 0;
    for (int i = 0; i < P; i++) { acc += st->tiles_touched[i]; st->point_offsets[i] = (uint32_t)acc; }
    const size_t R = (size_t)acc;
    st->num_rendered = (long long)R;
    if (num_rendered_out) *num_rendered_out = (int)R;

    // ---- K3 duplicateWithKeys (impl.cu:71-112): y-major / x-minor, ascending idx ----
    std::vector<uint64_t> keys(R);
    std::vector<uint32_t> vals(R);
#pragma omp parallel for schedule(static)
    for (int idx = 0; idx < P; idx++) {
        if (st->radii[idx] <= 0) continue;
        size_t off = idx == 0 ? 0 : st->point_offsets[idx - 1];
        uint32_t minx, miny, maxx, maxy;
        get_rect(st->means2D[2 * (size_t)idx], st->means2D[2 * (size_t)idx + 1], st->radii[idx],
                 gx, gy, minx, miny, maxx, maxy);
        uint32_t dbits;
        std::memcpy(&dbits, &st->depths[idx], 4);
        for (uint32_t y = miny; y < maxy; y++)
            for (uint32_t x = minx; x < maxx; x++) {
                keys[off] = ((uint64_t)(y * (uint32_t)gx + x) << 32) | dbits;
                vals[off] = (uint32_t)idx;
                off++;
            }
    }
    // ---- K4 stable sort by (tile, depth bits) (impl.cu:307-315; cub radix sort is stable) ----
    std::vector<uint32_t> order(R);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(),
                     [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
    st->keys.resize(R); st->point_list.resize(R);
    for (size_t i = 0; i < R; i++) { st->keys[i] = keys[order[i]]; st->point_list[i] = vals[order[i]]; }
    // ---- K5 identifyTileRanges (impl.cu:117-139), ranges zeroed first (impl.cu:317) ----
    st->ranges.assign(2 * (size_t)gx * gy, 0u);
    for (size_t i = 0; i < R; i++) {
        const uint32_t cur = (uint32_t)(st->keys[i] >> 32);
        if (i == 0) st->ranges[2 * cur] = 0;
        else {
            const uint32_t prev = (uint32_t)(st->keys[i - 1] >> 32);
            if (cur != prev) { st->ranges[2 * prev + 1] = (uint32_t)i; st->ranges[2 * cur] = (uint32_t)i; }
        }
        if (i == R - 1) st->ranges[2 * cur + 1] = (uint32_t)R;
    }

    // ---- K6 renderCUDA forward (fwd.cu:261-401) ----
    const float* feat = colors_precomp ? colors_precomp : st->rgb.data();
    st->final_T.assign((size_t)width * height, 0.f);
    st->n_contrib.assign((size_t)width * height, 0u);
    const size_t HW = (size_t)width * height;
#pragma omp parallel for schedule(dynamic, 1)
    for (int tile = 0; tile < gx * gy; tile++) {
        const int tx0 = (tile % gx) * BLOCK_X, ty0 = (tile / gx) * BLOCK_Y;
        const uint32_t r0 = st->ranges[2 * (size_t)tile], r1 = st->ranges[2 * (size_t)tile + 1];
        for (int py = ty0; py < std::min(ty0 + BLOCK_Y, height); py++)
            for (int px = tx0; px < std::min(tx0 + BLOCK_X, width); px++) {
                const float pxf = (float)px, pyf = (float)py;
                float T = 1.0f, C[NCH] = {0, 0, 0}, Dd = 0.0f;
                uint32_t contributor = 0, last_contributor = 0;
                for (uint32_t i = r0; i < r1; i++) {
                    contributor++;
                    const uint32_t g = st->point_list[i];
                    const float dx = st->means2D[2 * (size_t)g] - pxf;
                    const float dy = st->means2D[2 * (size_t)g + 1] - pyf;
                    const float* co = &st->conic_opacity[4 * (size_t)g];
                    // fwd.cu:346 as compiled: fma(fma(dx, cx*dx, (cz*dy)*dy), -0.5, -((cy*dx)*dy))
                    const float q = fmaf(dx, co[0] * dx, (co[2] * dy) * dy);
                    const float power = fmaf(q, -0.5f, -((co[1] * dx) * dy));
                    if (power > 0.0f) continue;
                    const float alpha = std::min(0.99f, co[3] * expf(power));
                    if (alpha < 1.0f / 255.0f) continue;
                    const float test_T = T * (1 - alpha);
                    if (test_T < 0.0001f) break;  // done = true; this splat is NOT blended
                    for (int ch = 0; ch < NCH; ch++)
                        C[ch] = fmaf(T, alpha * feat[(size_t)g * NCH + ch], C[ch]);
                    if (T > 0.5f) Dd = st->depths[g];  // median depth, fwd.cu:375-379
                    T = test_T;
                    last_contributor = contributor;
                }
                const size_t pix = (size_t)py * width + px;
                st->final_T[pix] = T;
                st->n_contrib[pix] = last_contributor;
                for (int ch = 0; ch < NCH; ch++)
                    out_color[ch * HW + pix] = fmaf(background[ch], T, C[ch]);
                out_depth[pix] = Dd;
            }
    }
    return st;
}

void gso_free(void* h) { delete (State*)h; }

// State accessors (sizes: P, 2P, 6P, 4P, 3P, P | R | 2*tiles | HW)
long long gso_num_rendered(void* h) { return ((State*)h)->num_rendered; }
void gso_get_geometry(void* h, float* depths, float* means2D, float* cov3D, float* conic_opacity,
                      float* rgb, uint32_t* tiles_touched)
{
    State* st = (State*)h;
    const size_t P = st->P;
    if (depths) std::memcpy(depths, st->depths.data(), 4 * P);
    if (means2D) std::memcpy(means2D, st->means2D.data(), 8 * P);
    if (cov3D) std::memcpy(cov3D, st->cov3D.data(), 24 * P);
    if (conic_opacity) std::memcpy(conic_opacity, st->conic_opacity.data(), 16 * P);
    if (rgb) std::memcpy(rgb, st->rgb.data(), 12 * P);
    if (tiles_touched) std::memcpy(tiles_touched, st->tiles_touched.data(), 4 * P);
}
void gso_get_binning(void* h, uint32_t* point_list, uint64_t* keys, uint32_t* ranges)
{
    State* st = (State*)h;
    if (point_list) std::memcpy(point_list, st->point_list.data(), 4 * st->point_list.size());
    if (keys) std::memcpy(keys, st->keys.data(), 8 * st->keys.size());
    if (ranges) std::memcpy(ranges, st->ranges.data(), 4 * st->ranges.size());
}
void gso_get_image_state(void* h, float* final_T, uint32_t* n_contrib)
{
    State* st = (State*)h;
    if (final_T) std::memcpy(final_T, st->final_T.data(), 4 * st->final_T.size());
    if (n_contrib) std::memcpy(n_contrib, st->n_contrib.data(), 4 * st->n_contrib.size());
}

// ---------------------------------------------------------------------------------------
// Backward: impl.cu:405-498; bwd.cu:399-557 (blend), :144-274 (cov2D), :346-396 + :278-341
// (projection, scale/rotation), :20-139 (SH).  Per-(pixel,splat) contributions are summed in
// double (the reference sums them with fp32 atomics in an unspecified order).
// Outputs follow src/Rasterizer.cu:253-261 shapes: dL_dmean2D[P,3], dL_dconic[P,4],
// dL_dopacity[P], dL_dcolor[P,3], dL_dmean3D[P,3], dL_dcov3D[P,6], dL_dsh[P,M,3],
// dL_dscale[P,3], dL_drot[P,4]; all fully written (zero where the reference leaves its
// zero-initialised buffers untouched).
// ---------------------------------------------------------------------------------------
void gso_backward(void* h, const float* background, const float* means3D, const float* shs,
                  const float* colors_precomp, const float* scales, float scale_modifier,
                  const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                  const float* projmatrix, const float* campos, float tan_fovx, float tan_fovy,
                  const float* dL_dpix, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                  float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
                  float* dL_dscale, float* dL_drot)
{
    State* st = (State*)h;
    const int P = st->P, W = st->W, H = st->H, gx = st->gx, gy = st->gy, M = st->M, D = st->D;
    const size_t HW = (size_t)W * H;
    const float focal_y = H / (2.0f * tan_fovy), focal_x = W / (2.0f * tan_fovx);
    const float* colors = colors_precomp ? colors_precomp : st->rgb.data();
    std::vector<double> a_mean2D(2 * (size_t)P, 0.0), a_conic(3 * (size_t)P, 0.0),
        a_opac(P, 0.0), a_col(3 * (size_t)P, 0.0);
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;  // bwd.cu:460-461

    // ---- K7 renderCUDA backward (bwd.cu:399-557) ----
#pragma omp parallel for schedule(dynamic, 1)
    for (int tile = 0; tile < gx * gy; tile++) {
        const int tx0 = (tile % gx) * BLOCK_X, ty0 = (tile / gx) * BLOCK_Y;
        const uint32_t r0 = st->ranges[2 * (size_t)tile], r1 = st->ranges[2 * (size_t)tile + 1];
        const int toDo = (int)(r1 - r0);
        for (int py = ty0; py < std::min(ty0 + BLOCK_Y, H); py++)
            for (int px = tx0; px < std::min(tx0 + BLOCK_X, W); px++) {
                const size_t pix = (size_t)py * W + px;
                const float pxf = (float)px, pyf = (float)py;
                const float T_final = st->final_T[pix];
                float T = T_final;
                const int last_contributor = (int)st->n_contrib[pix];
                float accum_rec[NCH] = {0, 0, 0}, last_color[NCH] = {0, 0, 0}, last_alpha = 0.f;
                float dpix[NCH];
                for (int ch = 0; ch < NCH; ch++) dpix[ch] = dL_dpix[ch * HW + pix];
                float bg_dot_dpixel = 0.f;
                for (int ch = 0; ch < NCH; ch++) bg_dot_dpixel += background[ch] * dpix[ch];
                for (int contributor = std::min(toDo, last_contributor) - 1; contributor >= 0; contributor--) {
                    const uint32_t g = st->point_list[r0 + contributor];
                    const float dx = st->means2D[2 * (size_t)g] - pxf;
                    const float dy = st->means2D[2 * (size_t)g + 1] - pyf;
                    const float* co = &st->conic_opacity[4 * (size_t)g];
                    const float q = fmaf(dx, co[0] * dx, (co[2] * dy) * dy);
                    const float power = fmaf(q, -0.5f, -((co[1] * dx) * dy));
                    if (power > 0.0f) continue;
                    const float G = expf(power);
                    const float alpha = std::min(0.99f, co[3] * G);
                    if (alpha < 1.0f / 255.0f) continue;
                    T = T / (1.f - alpha);
                    const float dchannel_dcolor = alpha * T;
                    float dL_dalpha = 0.0f;
                    for (int ch = 0; ch < NCH; ch++) {
                        const float c = colors[(size_t)g * NCH + ch];
                        accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
                        last_color[ch] = c;
                        dL_dalpha += (c - accum_rec[ch]) * dpix[ch];
                        const double v = (double)(dchannel_dcolor * dpix[ch]);
#pragma omp atomic
                        a_col[3 * (size_t)g + ch] += v;
                    }
                    dL_dalpha *= T;
                    last_alpha = alpha;
                    dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot_dpixel;
                    const float dL_dG = co[3] * dL_dalpha;
                    const float gdx = G * dx, gdy = G * dy;
                    const float dG_ddelx = -gdx * co[0] - gdy * co[1];
                    const float dG_ddely = -gdy * co[2] - gdx * co[1];
                    const double v0 = (double)(dL_dG * dG_ddelx * ddelx_dx);
                    const double v1 = (double)(dL_dG * dG_ddely * ddely_dy);
                    const double v2 = (double)(-0.5f * gdx * dx * dL_dG);
                    const double v3 = (double)(-0.5f * gdx * dy * dL_dG);
                    const double v4 = (double)(-0.5f * gdy * dy * dL_dG);
                    const double v5 = (double)(G * dL_dalpha);
#pragma omp atomic
                    a_mean2D[2 * (size_t)g] += v0;
#pragma omp atomic
                    a_mean2D[2 * (size_t)g + 1] += v1;
#pragma omp atomic
                    a_conic[3 * (size_t)g] += v2;
#pragma omp atomic
                    a_conic[3 * (size_t)g + 1] += v3;
#pragma omp atomic
                    a_conic[3 * (size_t)g + 2] += v4;
#pragma omp atomic
                    a_opac[g] += v5;
                }
            }
    }

    // ---- K8 + K9 per-Gaussian backward ----
    const float* v = viewmatrix;
    const float* proj = projmatrix;
#pragma omp parallel for schedule(static)
    for (int idx = 0; idx < P; idx++) {
        const size_t i = (size_t)idx;
        dL_dmean2D[3 * i] = (float)a_mean2D[2 * i];
        dL_dmean2D[3 * i + 1] = (float)a_mean2D[2 * i + 1];
        dL_dmean2D[3 * i + 2] = 0.f;
        dL_dconic[4 * i] = (float)a_conic[3 * i];
        dL_dconic[4 * i + 1] = (float)a_conic[3 * i + 1];
        dL_dconic[4 * i + 2] = 0.f;  // [1][0] slot never written (bwd.cu:549-551)
        dL_dconic[4 * i + 3] = (float)a_conic[3 * i + 2];
        dL_dopacity[i] = (float)a_opac[i];
        for (int ch = 0; ch < 3; ch++) dL_dcolor[3 * i + ch] = (float)a_col[3 * i + ch];
        for (int k = 0; k < 3; k++) dL_dmean3D[3 * i + k] = 0.f;
        for (int k = 0; k < 6; k++) dL_dcov3D[6 * i + k] = 0.f;
        if (dL_dscale) for (int k = 0; k < 3; k++) dL_dscale[3 * i + k] = 0.f;
        if (dL_drot) for (int k = 0; k < 4; k++) dL_drot[4 * i + k] = 0.f;
        if (dL_dsh) for (int k = 0; k < 3 * M; k++) dL_dsh[i * 3 * M + k] = 0.f;
        if (!(st->radii[idx] > 0)) continue;  // bwd.cu:156, :367

        // ---- computeCov2DCUDA (bwd.cu:144-274) ----
        const float* cov3D = cov3D_precomp ? cov3D_precomp + 6 * i : &st->cov3D[6 * i];
        const float* mean = means3D + 3 * i;
        float cov[3];
        Cov2DInter in;
        compute_cov2d(mean, focal_x, focal_y, tan_fovx, tan_fovy, cov3D, v, cov, &in);
        const float limx = 1.3f * tan_fovx, limy = 1.3f * tan_fovy;
        const float x_grad_mul = (in.txtz < -limx || in.txtz > limx) ? 0.f : 1.f;
        const float y_grad_mul = (in.tytz < -limy || in.tytz > limy) ? 0.f : 1.f;
        const float a = cov[0], b = cov[1], c = cov[2];
        const float dcx = dL_dconic[4 * i], dcy = dL_dconic[4 * i + 1], dcz = dL_dconic[4 * i + 3];
        const float denom = a * c - b * b;
        float dL_da = 0, dL_db = 0, dL_dc = 0;
        const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
        // glm T[c][r]: T[0][k] = T0k, T[1][k] = T1k  (bwd.cu:192)
        const float T00 = in.T00, T01 = in.T01, T02 = in.T02, T10 = in.T10, T11 = in.T11, T12 = in.T12;
        if (denom2inv != 0) {
            dL_da = denom2inv * (-c * c * dcx + 2 * b * c * dcy + (denom - a * c) * dcz);
            dL_dc = denom2inv * (-a * a * dcz + 2 * a * b * dcy + (denom - a * c) * dcx);
            dL_db = denom2inv * 2 * (b * c * dcx - (denom + 2 * b * b) * dcy + a * b * dcz);
            dL_dcov3D[6 * i + 0] = (T00 * T00 * dL_da + T00 * T10 * dL_db + T10 * T10 * dL_dc);
            dL_dcov3D[6 * i + 3] = (T01 * T01 * dL_da + T01 * T11 * dL_db + T11 * T11 * dL_dc);
            dL_dcov3D[6 * i + 5] = (T02 * T02 * dL_da + T02 * T12 * dL_db + T12 * T12 * dL_dc);
            dL_dcov3D[6 * i + 1] = 2 * T00 * T01 * dL_da + (T00 * T11 + T01 * T10) * dL_db + 2 * T10 * T11 * dL_dc;
            dL_dcov3D[6 * i + 2] = 2 * T00 * T02 * dL_da + (T00 * T12 + T02 * T10) * dL_db + 2 * T10 * T12 * dL_dc;
            dL_dcov3D[6 * i + 4] = 2 * T02 * T01 * dL_da + (T01 * T12 + T02 * T11) * dL_db + 2 * T11 * T12 * dL_dc;
        }
        // Vrk[c][r] symmetric
        const float V[3][3] = {{cov3D[0], cov3D[1], cov3D[2]}, {cov3D[1], cov3D[3], cov3D[4]}, {cov3D[2], cov3D[4], cov3D[5]}};
        const float dL_dT00 = 2 * (T00 * V[0][0] + T01 * V[0][1] + T02 * V[0][2]) * dL_da + (T10 * V[0][0] + T11 * V[0][1] + T12 * V[0][2]) * dL_db;
        const float dL_dT01 = 2 * (T00 * V[1][0] + T01 * V[1][1] + T02 * V[1][2]) * dL_da + (T10 * V[1][0] + T11 * V[1][1] + T12 * V[1][2]) * dL_db;
        const float dL_dT02 = 2 * (T00 * V[2][0] + T01 * V[2][1] + T02 * V[2][2]) * dL_da + (T10 * V[2][0] + T11 * V[2][1] + T12 * V[2][2]) * dL_db;
        const float dL_dT10 = 2 * (T10 * V[0][0] + T11 * V[0][1] + T12 * V[0][2]) * dL_dc + (T00 * V[0][0] + T01 * V[0][1] + T02 * V[0][2]) * dL_db;
        const float dL_dT11 = 2 * (T10 * V[1][0] + T11 * V[1][1] + T12 * V[1][2]) * dL_dc + (T00 * V[1][0] + T01 * V[1][1] + T02 * V[1][2]) * dL_db;
        const float dL_dT12 = 2 * (T10 * V[2][0] + T11 * V[2][1] + T12 * V[2][2]) * dL_dc + (T00 * V[2][0] + T01 * V[2][1] + T02 * V[2][2]) * dL_db;
        // W[c][r]: W[0]=(v0,v4,v8), W[1]=(v1,v5,v9), W[2]=(v2,v6,v10)   (bwd.cu:182-185)
        const float dL_dJ00 = v[0] * dL_dT00 + v[4] * dL_dT01 + v[8] * dL_dT02;
        const float dL_dJ02 = v[2] * dL_dT00 + v[6] * dL_dT01 + v[10] * dL_dT02;
        const float dL_dJ11 = v[1] * dL_dT10 + v[5] * dL_dT11 + v[9] * dL_dT12;
        const float dL_dJ12 = v[2] * dL_dT10 + v[6] * dL_dT11 + v[10] * dL_dT12;
        const float tz = 1.f / in.tz, tz2 = tz * tz, tz3 = tz2 * tz;
        const float dL_dtx = x_grad_mul * -focal_x * tz2 * dL_dJ02;
        const float dL_dty = y_grad_mul * -focal_y * tz2 * dL_dJ12;
        const float dL_dtz = -focal_x * tz2 * dL_dJ00 - focal_y * tz2 * dL_dJ11 +
                             (2 * focal_x * in.tx) * tz3 * dL_dJ02 + (2 * focal_y * in.ty) * tz3 * dL_dJ12;
        // transformVec4x3Transpose (aux.h:89-97); ASSIGNMENT (bwd.cu:273)
        float dmx = v[0] * dL_dtx + v[1] * dL_dty + v[2] * dL_dtz;
        float dmy = v[4] * dL_dtx + v[5] * dL_dty + v[6] * dL_dtz;
        float dmz = v[8] * dL_dtx + v[9] * dL_dty + v[10] * dL_dtz;

        // ---- preprocessCUDA backward (bwd.cu:346-396) ----
        const float mx = mean[0], my = mean[1], mz = mean[2];
        const float m_homw = proj[3] * mx + proj[7] * my + proj[11] * mz + proj[15];
        const float m_w = 1.0f / (m_homw + 0.0000001f);
        const float mul1 = (proj[0] * mx + proj[4] * my + proj[8] * mz + proj[12]) * m_w * m_w;
        const float mul2 = (proj[1] * mx + proj[5] * my + proj[9] * mz + proj[13]) * m_w * m_w;
        const float g2x = dL_dmean2D[3 * i], g2y = dL_dmean2D[3 * i + 1];
        dmx += (proj[0] * m_w - proj[3] * mul1) * g2x + (proj[1] * m_w - proj[3] * mul2) * g2y;
        dmy += (proj[4] * m_w - proj[7] * mul1) * g2x + (proj[5] * m_w - proj[7] * mul2) * g2y;
        dmz += (proj[8] * m_w - proj[11] * mul1) * g2x + (proj[9] * m_w - proj[11] * mul2) * g2y;

        // ---- SH backward (bwd.cu:20-139) ----
        if (shs) {
            const float* sh = shs + i * M * 3;
            float* dsh = dL_dsh + i * M * 3;
            const float ox = mx - campos[0], oy = my - campos[1], oz = mz - campos[2];
            const float len = std::sqrt(ox * ox + oy * oy + oz * oz);
            const float x = ox / len, y = oy / len, z = oz / len;
            float dRGB[3];
            for (int ch = 0; ch < 3; ch++) dRGB[ch] = dL_dcolor[3 * i + ch] * (st->clamped[3 * i + ch] ? 0.f : 1.f);
            float dRGBdx[3] = {0, 0, 0}, dRGBdy[3] = {0, 0, 0}, dRGBdz[3] = {0, 0, 0};
            auto S = [&](int k, int ch) { return sh[k * 3 + ch]; };
            auto setd = [&](int k, float w) { for (int ch = 0; ch < 3; ch++) dsh[k * 3 + ch] = w * dRGB[ch]; };
            setd(0, SH_C0);
            if (D > 0) {
                setd(1, -SH_C1 * y); setd(2, SH_C1 * z); setd(3, -SH_C1 * x);
                for (int ch = 0; ch < 3; ch++) {
                    dRGBdx[ch] = -SH_C1 * S(3, ch); dRGBdy[ch] = -SH_C1 * S(1, ch); dRGBdz[ch] = SH_C1 * S(2, ch);
                }
                if (D > 1) {
                    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                    setd(4, SH_C2[0] * xy); setd(5, SH_C2[1] * yz); setd(6, SH_C2[2] * (2.f * zz - xx - yy));
                    setd(7, SH_C2[3] * xz); setd(8, SH_C2[4] * (xx - yy));
                    for (int ch = 0; ch < 3; ch++) {
                        dRGBdx[ch] += SH_C2[0] * y * S(4, ch) + SH_C2[2] * 2.f * -x * S(6, ch) + SH_C2[3] * z * S(7, ch) + SH_C2[4] * 2.f * x * S(8, ch);
                        dRGBdy[ch] += SH_C2[0] * x * S(4, ch) + SH_C2[1] * z * S(5, ch) + SH_C2[2] * 2.f * -y * S(6, ch) + SH_C2[4] * 2.f * -y * S(8, ch);
                        dRGBdz[ch] += SH_C2[1] * y * S(5, ch) + SH_C2[2] * 2.f * 2.f * z * S(6, ch) + SH_C2[3] * x * S(7, ch);
                    }
                    if (D > 2) {
                        setd(9, SH_C3[0] * y * (3.f * xx - yy)); setd(10, SH_C3[1] * xy * z);
                        setd(11, SH_C3[2] * y * (4.f * zz - xx - yy));
                        setd(12, SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy));
                        setd(13, SH_C3[4] * x * (4.f * zz - xx - yy)); setd(14, SH_C3[5] * z * (xx - yy));
                        setd(15, SH_C3[6] * x * (xx - 3.f * yy));
                        for (int ch = 0; ch < 3; ch++) {
                            dRGBdx[ch] += SH_C3[0] * S(9, ch) * 3.f * 2.f * xy + SH_C3[1] * S(10, ch) * yz + SH_C3[2] * S(11, ch) * -2.f * xy +
                                          SH_C3[3] * S(12, ch) * -3.f * 2.f * xz + SH_C3[4] * S(13, ch) * (-3.f * xx + 4.f * zz - yy) +
                                          SH_C3[5] * S(14, ch) * 2.f * xz + SH_C3[6] * S(15, ch) * 3.f * (xx - yy);
                            dRGBdy[ch] += SH_C3[0] * S(9, ch) * 3.f * (xx - yy) + SH_C3[1] * S(10, ch) * xz + SH_C3[2] * S(11, ch) * (-3.f * yy + 4.f * zz - xx) +
                                          SH_C3[3] * S(12, ch) * -3.f * 2.f * yz + SH_C3[4] * S(13, ch) * -2.f * xy +
                                          SH_C3[5] * S(14, ch) * -2.f * yz + SH_C3[6] * S(15, ch) * -3.f * 2.f * xy;
                            dRGBdz[ch] += SH_C3[1] * S(10, ch) * xy + SH_C3[2] * S(11, ch) * 4.f * 2.f * yz + SH_C3[3] * S(12, ch) * 3.f * (2.f * zz - xx - yy) +
                                          SH_C3[4] * S(13, ch) * 4.f * 2.f * xz + SH_C3[5] * S(14, ch) * (xx - yy);
                        }
                    }
                }
            }
            float ddir[3] = {0, 0, 0};
            for (int ch = 0; ch < 3; ch++) { ddir[0] += dRGBdx[ch] * dRGB[ch]; ddir[1] += dRGBdy[ch] * dRGB[ch]; ddir[2] += dRGBdz[ch] * dRGB[ch]; }
            // dnormvdv (aux.h:107-118)
            const float sum2 = ox * ox + oy * oy + oz * oz;
            const float invsum32 = 1.0f / std::sqrt(sum2 * sum2 * sum2);
            dmx += ((+sum2 - ox * ox) * ddir[0] - oy * ox * ddir[1] - oz * ox * ddir[2]) * invsum32;
            dmy += (-ox * oy * ddir[0] + (sum2 - oy * oy) * ddir[1] - oz * oy * ddir[2]) * invsum32;
            dmz += (-ox * oz * ddir[0] - oy * oz * ddir[1] + (sum2 - oz * oz) * ddir[2]) * invsum32;
        }
        dL_dmean3D[3 * i] = dmx; dL_dmean3D[3 * i + 1] = dmy; dL_dmean3D[3 * i + 2] = dmz;

        // ---- computeCov3D backward (bwd.cu:278-341) ----
        if (scales) {
            const float* q = rotations + 4 * i;
            const float r = q[0], x = q[1], y = q[2], z = q[3];
            // glm R[c][r] as in the forward
            const float R[3][3] = {
                {1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
            const float s[3] = {scale_modifier * scales[3 * i], scale_modifier * scales[3 * i + 1], scale_modifier * scales[3 * i + 2]};
            float Mm[3][3];  // M = S*R: M[c][r] = s_r * R[c][r]
            for (int c2 = 0; c2 < 3; c2++) for (int r2 = 0; r2 < 3; r2++) Mm[c2][r2] = s[r2] * R[c2][r2];
            const float* g = dL_dcov3D + 6 * i;
            const float dS[3][3] = {{g[0], 0.5f * g[1], 0.5f * g[2]}, {0.5f * g[1], g[3], 0.5f * g[4]}, {0.5f * g[2], 0.5f * g[4], g[5]}};
            // dL_dM = 2 * M * dL_dSigma: (A*B)[c][r] = sum_k A[k][r] * B[c][k]
            float dM[3][3];
            for (int c2 = 0; c2 < 3; c2++) for (int r2 = 0; r2 < 3; r2++) {
                float acc = 0.f;
                for (int k = 0; k < 3; k++) acc += (2.0f * Mm[k][r2]) * dS[c2][k];
                dM[c2][r2] = acc;
            }
            // Rt = transpose(R): Rt[c][r] = R[r][c]; dL_dMt[c][r] = dM[r][c]
            float dMt[3][3];
            for (int c2 = 0; c2 < 3; c2++) for (int r2 = 0; r2 < 3; r2++) dMt[c2][r2] = dM[r2][c2];
            for (int k = 0; k < 3; k++) {
                // dot(Rt[k], dL_dMt[k]) = sum_r R[r][k] * dMt[k][r]
                float acc = 0.f;
                for (int r2 = 0; r2 < 3; r2++) acc += R[r2][k] * dMt[k][r2];
                dL_dscale[3 * i + k] = acc;
            }
            for (int k = 0; k < 3; k++) for (int r2 = 0; r2 < 3; r2++) dMt[k][r2] *= s[k];
            float dq[4];
            dq[0] = 2 * z * (dMt[0][1] - dMt[1][0]) + 2 * y * (dMt[2][0] - dMt[0][2]) + 2 * x * (dMt[1][2] - dMt[2][1]);
            dq[1] = 2 * y * (dMt[1][0] + dMt[0][1]) + 2 * z * (dMt[2][0] + dMt[0][2]) + 2 * r * (dMt[1][2] - dMt[2][1]) - 4 * x * (dMt[2][2] + dMt[1][1]);
            dq[2] = 2 * x * (dMt[1][0] + dMt[0][1]) + 2 * r * (dMt[2][0] - dMt[0][2]) + 2 * z * (dMt[1][2] + dMt[2][1]) - 4 * y * (dMt[2][2] + dMt[0][0]);
            dq[3] = 2 * r * (dMt[0][1] - dMt[1][0]) + 2 * x * (dMt[2][0] + dMt[0][2]) + 2 * y * (dMt[1][2] + dMt[2][1]) - 4 * z * (dMt[1][1] + dMt[0][0]);
            for (int k = 0; k < 4; k++) dL_drot[4 * i + k] = dq[k];  // no normalisation Jacobian (bwd.cu:340)
        }
    }
}

// ---------------------------------------------------------------------------------------
// visible_filter (impl.cu:348-401, fwd.cu:404-473): radii only.
// ---------------------------------------------------------------------------------------
void gso_visible_filter(int P, int width, int height, const float* means3D, const float* scales,
                        float scale_modifier, const float* rotations, const float* viewmatrix,
                        const float* projmatrix, float tan_fovx, float tan_fovy, int* radii)
{
    const float focal_y = height / (2.0f * tan_fovy), focal_x = width / (2.0f * tan_fovx);
    const int gx = (width + BLOCK_X - 1) / BLOCK_X, gy = (height + BLOCK_Y - 1) / BLOCK_Y;
#pragma omp parallel for schedule(static)
    for (int idx = 0; idx < P; idx++) {
        radii[idx] = 0;
        const float* p = means3D + 3 * (size_t)idx;
        const float pvz = xform_row(viewmatrix, 2, p[0], p[1], p[2]);
        if (pvz <= 0.2f) continue;
        const float hx = xform_row(projmatrix, 0, p[0], p[1], p[2]);
        const float hy = xform_row(projmatrix, 1, p[0], p[1], p[2]);
        const float hw = xform_row(projmatrix, 3, p[0], p[1], p[2]);
        const float p_w = 1.0f / (hw + 0.0000001f);
        float cov3D[6], cov[3];
        compute_cov3d(scales + 3 * (size_t)idx, scale_modifier, rotations + 4 * (size_t)idx, cov3D);
        compute_cov2d(p, focal_x, focal_y, tan_fovx, tan_fovy, cov3D, viewmatrix, cov, nullptr);
        const float det = fmaf(cov[0], cov[2], -(cov[1] * cov[1]));
        if (det == 0.0f) continue;
        const float mid = (cov[0] + cov[2]) * 0.5f;
        const float sq = std::sqrt(std::max(fmaf(mid, mid, -det), 0.1f));
        const float my_radius = std::ceil(std::sqrt(std::max(mid + sq, mid - sq)) * 3.f);
        const float px = ndc2pix(hx * p_w, width), py = ndc2pix(hy * p_w, height);
        uint32_t minx, miny, maxx, maxy;
        get_rect(px, py, (int)my_radius, gx, gy, minx, miny, maxx, maxy);
        if ((maxx - minx) * (maxy - miny) == 0) continue;
        radii[idx] = (int)my_radius;
    }
}

// markVisible / checkFrustum (impl.cu:55-67, 142-154)
void gso_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                      uint8_t* present)
{
    (void)projmatrix;
    for (int idx = 0; idx < P; idx++) {
        const float* p = means3D + 3 * (size_t)idx;
        present[idx] = xform_row(viewmatrix, 2, p[0], p[1], p[2]) > 0.2f;
    }
}

// ---------------------------------------------------------------------------------------
// SimpleKNN::knn (src/simple_knn.cu:185-221): mean squared distance to the 3 nearest
// neighbours, found through a Morton ordering and 1024-point boxes.
// ---------------------------------------------------------------------------------------
static uint32_t prep_morton(uint32_t x)  // simple_knn.cu:45-52
{
    x = (x | (x << 16)) & 0x030000FF;
    x = (x | (x << 8)) & 0x0300F00F;
    x = (x | (x << 4)) & 0x030C30C3;
    x = (x | (x << 2)) & 0x09249249;
    return x;
}
static inline void update_kbest3(const float* ref, const float* pt, float* knn)  // simple_knn.cu:131-145
{
    const float dx = pt[0] - ref[0], dy = pt[1] - ref[1], dz = pt[2] - ref[2];
    float dist = nv3(dx, dx, dy, dy, dz, dz);
    for (int j = 0; j < 3; j++)
        if (knn[j] > dist) { float t = knn[j]; knn[j] = dist; dist = t; }
}
void gso_knn(int P, const float* points, float* mean_dists)
{
    const int BOX = 1024;
    // min / max reductions are seeded with (0,0,0) (simple_knn.cu:191-199): kept for parity
    float mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
    for (int i = 0; i < P; i++)
        for (int k = 0; k < 3; k++) {
            mn[k] = std::min(mn[k], points[3 * (size_t)i + k]);
            mx[k] = std::max(mx[k], points[3 * (size_t)i + k]);
        }
    std::vector<uint32_t> codes(P), idx(P);
    for (int i = 0; i < P; i++) {  // simple_knn.cu:54-61
        uint32_t c[3];
        for (int k = 0; k < 3; k++) {
            const float f = ((points[3 * (size_t)i + k] - mn[k]) / (mx[k] - mn[k])) * 1023.0f;
            c[k] = prep_morton((uint32_t)f);
        }
        codes[i] = c[0] | (c[1] << 1) | (c[2] << 2);
        idx[i] = (uint32_t)i;
    }
    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return codes[a] < codes[b]; });
    const int nb = (P + BOX - 1) / BOX;
    std::vector<float> bmin(3 * (size_t)nb, FLT_MAX), bmax(3 * (size_t)nb, -FLT_MAX);
    for (int i = 0; i < P; i++) {  // boxMinMax simple_knn.cu:78-117
        const int b = i / BOX;
        for (int k = 0; k < 3; k++) {
            const float vv = points[3 * (size_t)idx[i] + k];
            bmin[3 * b + k] = std::min(bmin[3 * b + k], vv);
            bmax[3 * b + k] = std::max(bmax[3 * b + k], vv);
        }
    }
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < P; i++) {  // boxMeanDist simple_knn.cu:147-183
        const float* pt = points + 3 * (size_t)idx[i];
        float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
        for (int j = std::max(0, i - 3); j <= std::min(P - 1, i + 3); j++) {
            if (j == i) continue;
            update_kbest3(pt, points + 3 * (size_t)idx[j], best);
        }
        const float reject = best[2];
        best[0] = best[1] = best[2] = FLT_MAX;
        for (int b = 0; b < nb; b++) {
            float d[3] = {0, 0, 0};  // distBoxPoint simple_knn.cu:119-129
            for (int k = 0; k < 3; k++)
                if (pt[k] < bmin[3 * b + k] || pt[k] > bmax[3 * b + k])
                    d[k] = std::min(std::fabs(pt[k] - bmin[3 * b + k]), std::fabs(pt[k] - bmax[3 * b + k]));
            const float dist = nv3(d[0], d[0], d[1], d[1], d[2], d[2]);
            if (dist > reject || dist > best[2]) continue;
            for (int j = b * BOX; j < std::min(P, (b + 1) * BOX); j++) {
                if (j == i) continue;
                update_kbest3(pt, points + 3 * (size_t)idx[j], best);
            }
        }
        mean_dists[idx[i]] = (best[0] + best[1] + best[2]) / 3.0f;
    }
}

int gso_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
