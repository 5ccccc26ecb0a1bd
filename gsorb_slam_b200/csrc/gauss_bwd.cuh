// gauss_bwd.cuh -- the per-Gaussian backward of ONE row in registers (K8 + K9: BACKWARD::preprocess, backward.cu:560-621 =
// computeCov2DCUDA :144-274 + preprocessCUDA<3> :346-396 + computeCov3D backward :278-341 + SH backward :20-139), shared by
// gauss_backward_kernel (gauss_bwd.cu: writes the reference-layout gradient arrays) and map_update_kernel (map_update.cu: carries
// on through the activation chain rule and the Adam step without writing them).
#pragma once
#include "common.cuh"

namespace gsb {

__device__ __forceinline__ void sh_backward(int deg, int M, const float* __restrict__ sh, float* __restrict__ dsh,
                                            float mx, float my, float mz, const float* __restrict__ campos,
                                            const float* dcol, uint32_t clamped, float& dmx, float& dmy, float& dmz)
{
    const float ox = mx - campos[0], oy = my - campos[1], oz = mz - campos[2];
    const float len = sqrtf(ox * ox + oy * oy + oz * oz);
    const float x = ox / len, y = oy / len, z = oz / len;
    float dRGB[3];
#pragma unroll
    for (int ch = 0; ch < 3; ch++) dRGB[ch] = (clamped >> ch) & 1 ? 0.f : dcol[ch];
    float ddir[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        auto S = [&](int k) { return sh[k * 3 + ch]; };
        auto setd = [&](int k, float w) { if (dsh) dsh[k * 3 + ch] = w * dRGB[ch]; };
        float gx = 0.f, gy = 0.f, gz = 0.f;
        setd(0, SH_C0);
        if (deg > 0) {
            setd(1, -SH_C1 * y); setd(2, SH_C1 * z); setd(3, -SH_C1 * x);
            gx = -SH_C1 * S(3); gy = -SH_C1 * S(1); gz = SH_C1 * S(2);
            if (deg > 1) {
                const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                setd(4, SH_C2[0] * xy); setd(5, SH_C2[1] * yz); setd(6, SH_C2[2] * (2.f * zz - xx - yy));
                setd(7, SH_C2[3] * xz); setd(8, SH_C2[4] * (xx - yy));
                gx += SH_C2[0] * y * S(4) + SH_C2[2] * 2.f * -x * S(6) + SH_C2[3] * z * S(7) + SH_C2[4] * 2.f * x * S(8);
                gy += SH_C2[0] * x * S(4) + SH_C2[1] * z * S(5) + SH_C2[2] * 2.f * -y * S(6) + SH_C2[4] * 2.f * -y * S(8);
                gz += SH_C2[1] * y * S(5) + SH_C2[2] * 2.f * 2.f * z * S(6) + SH_C2[3] * x * S(7);
                if (deg > 2) {
                    setd(9, SH_C3[0] * y * (3.f * xx - yy)); setd(10, SH_C3[1] * xy * z);
                    setd(11, SH_C3[2] * y * (4.f * zz - xx - yy));
                    setd(12, SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy));
                    setd(13, SH_C3[4] * x * (4.f * zz - xx - yy)); setd(14, SH_C3[5] * z * (xx - yy));
                    setd(15, SH_C3[6] * x * (xx - 3.f * yy));
                    gx += SH_C3[0] * S(9) * 3.f * 2.f * xy + SH_C3[1] * S(10) * yz + SH_C3[2] * S(11) * -2.f * xy +
                          SH_C3[3] * S(12) * -3.f * 2.f * xz + SH_C3[4] * S(13) * (-3.f * xx + 4.f * zz - yy) +
                          SH_C3[5] * S(14) * 2.f * xz + SH_C3[6] * S(15) * 3.f * (xx - yy);
                    gy += SH_C3[0] * S(9) * 3.f * (xx - yy) + SH_C3[1] * S(10) * xz +
                          SH_C3[2] * S(11) * (-3.f * yy + 4.f * zz - xx) + SH_C3[3] * S(12) * -3.f * 2.f * yz +
                          SH_C3[4] * S(13) * -2.f * xy + SH_C3[5] * S(14) * -2.f * yz + SH_C3[6] * S(15) * -3.f * 2.f * xy;
                    gz += SH_C3[1] * S(10) * xy + SH_C3[2] * S(11) * 4.f * 2.f * yz +
                          SH_C3[3] * S(12) * 3.f * (2.f * zz - xx - yy) + SH_C3[4] * S(13) * 4.f * 2.f * xz +
                          SH_C3[5] * S(14) * (xx - yy);
                }
            }
        }
        for (int k = (deg + 1) * (deg + 1); k < M; k++) setd(k, 0.f);  // inactive coefficients
        ddir[0] += gx * dRGB[ch];
        ddir[1] += gy * dRGB[ch];
        ddir[2] += gz * dRGB[ch];
    }
    // dnormvdv (auxiliary.h:107-118)
    const float sum2 = ox * ox + oy * oy + oz * oz;
    const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
    dmx += ((+sum2 - ox * ox) * ddir[0] - oy * ox * ddir[1] - oz * ox * ddir[2]) * invsum32;
    dmy += (-ox * oy * ddir[0] + (sum2 - oy * oy) * ddir[1] - oz * oy * ddir[2]) * invsum32;
    dmz += (-ox * oz * ddir[0] - oy * oz * ddir[1] + (sum2 - oz * oz) * ddir[2]) * invsum32;
}

// packed sums from the blend backward (GradAcc, common.cuh) -> a[0..1] dL/dmean2D, a[2..4] dL/dconic (xx, xy, yy), a[5] dL/dopacity,
// a[6..8] dL/dcolour; rb = {conic.x, conic.y, conic.z, opacity} of the Gaussian's record
__device__ __forceinline__ void gauss_moments_to_2d(const FwdParams& p, bool rendered, const float4& a0, const float4& a1, const float4& a2,
                                                    const float4& a3, const float4& rb, float* a)
{
    // packed sums from the blend backward -> reference-layout 2D gradients (backward.cu:536-554):
    //   dL/dmean2D = -0.5 W o (A X + B Y), -0.5 H o (C Y + B X);  dL/dconic = -0.5 o (XX, XY, YY);  dL/dopacity = U
    if (rendered) {
        // GradAcc: a = {S u dx, S u dx^2, S w d_b, -}, b = {S u dxdy, S w d_r, -, -}, c = {S u dy, S u dy^2, S w d_z, -}, d = {S u, S w d_g, -, -}
        const float X = a0.x, XX = a0.y, XY = a1.x, Y = a2.x, YY = a2.y, U = a3.x;
        a[0] = -0.5f * p.W * rb.w * (rb.x * X + rb.y * Y);
        a[1] = -0.5f * p.H * rb.w * (rb.z * Y + rb.y * X);
        a[2] = -0.5f * rb.w * XX;
        a[3] = -0.5f * rb.w * XY;
        a[4] = -0.5f * rb.w * YY;
        a[5] = U;
        a[6] = a1.y; a[7] = a3.y;
        a[8] = a0.z;
    } else {
#pragma unroll
        for (int k = 0; k < 9; k++) a[k] = 0.f;
    }
}

// Everything behind the 2D gradients for a RENDERED Gaussian: dL/dcov3D, dL/dmean3D, dL/dscale, dL/drot (and dL/dsh when SH).
// (mx, my, mz) is the mean the rasterizer saw, qv / s0..s2 the rotation and scale it saw.
template <bool SH>
__device__ __forceinline__ void gauss_backward_chain(const FwdParams& p, size_t i, const float* a, float mx, float my, float mz,
                                                     const float4& qv, float s0, float s1, float s2, float* dL_dsh, uint32_t clamped,
                                                     float* dcov, float& dmx, float& dmy, float& dmz, float* dscale, float* drot)
{
    float cov3D[6];
    const float qr = qv.x, qx = qv.y, qy = qv.z, qz = qv.w;
    if (p.cov3D_precomp) {
#pragma unroll
        for (int k = 0; k < 6; k++) cov3D[k] = p.cov3D_precomp[6 * i + k];
    } else {
        compute_cov3d(s0, s1, s2, p.scale_modifier, qr, qx, qy, qz, cov3D);
    }
    // ---- computeCov2DCUDA (backward.cu:144-274) ----
    const float* v = p.viewmatrix;
    float cov[3];
    Cov2DInter in;
    compute_cov2d(mx, my, mz, p.focal_x, p.focal_y, p.tan_fovx, p.tan_fovy, cov3D, v, cov, &in);
    const float limx = 1.3f * p.tan_fovx, limy = 1.3f * p.tan_fovy;
    const float x_grad_mul = (in.txtz < -limx || in.txtz > limx) ? 0.f : 1.f;
    const float y_grad_mul = (in.tytz < -limy || in.tytz > limy) ? 0.f : 1.f;
    const float ca = cov[0], cb = cov[1], cc = cov[2];
    const float dcx = a[2], dcy = a[3], dcz = a[4];
    const float denom = ca * cc - cb * cb;
    float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
    const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
    const float T00 = in.T00, T01 = in.T01, T02 = in.T02, T10 = in.T10, T11 = in.T11, T12 = in.T12;
    if (denom2inv != 0.f) {
        dL_da = denom2inv * (-cc * cc * dcx + 2 * cb * cc * dcy + (denom - ca * cc) * dcz);
        dL_dc = denom2inv * (-ca * ca * dcz + 2 * ca * cb * dcy + (denom - ca * cc) * dcx);
        dL_db = denom2inv * 2 * (cb * cc * dcx - (denom + 2 * cb * cb) * dcy + ca * cb * dcz);
        dcov[0] = T00 * T00 * dL_da + T00 * T10 * dL_db + T10 * T10 * dL_dc;
        dcov[3] = T01 * T01 * dL_da + T01 * T11 * dL_db + T11 * T11 * dL_dc;
        dcov[5] = T02 * T02 * dL_da + T02 * T12 * dL_db + T12 * T12 * dL_dc;
        dcov[1] = 2 * T00 * T01 * dL_da + (T00 * T11 + T01 * T10) * dL_db + 2 * T10 * T11 * dL_dc;
        dcov[2] = 2 * T00 * T02 * dL_da + (T00 * T12 + T02 * T10) * dL_db + 2 * T10 * T12 * dL_dc;
        dcov[4] = 2 * T02 * T01 * dL_da + (T01 * T12 + T02 * T11) * dL_db + 2 * T11 * T12 * dL_dc;
    }
    // rows of Vrk (symmetric)
    const float V0[3] = {cov3D[0], cov3D[1], cov3D[2]}, V1[3] = {cov3D[1], cov3D[3], cov3D[4]},
                V2[3] = {cov3D[2], cov3D[4], cov3D[5]};
    const float t0v0 = T00 * V0[0] + T01 * V0[1] + T02 * V0[2], t0v1 = T00 * V1[0] + T01 * V1[1] + T02 * V1[2],
                t0v2 = T00 * V2[0] + T01 * V2[1] + T02 * V2[2];
    const float t1v0 = T10 * V0[0] + T11 * V0[1] + T12 * V0[2], t1v1 = T10 * V1[0] + T11 * V1[1] + T12 * V1[2],
                t1v2 = T10 * V2[0] + T11 * V2[1] + T12 * V2[2];
    const float dL_dT00 = 2 * t0v0 * dL_da + t1v0 * dL_db, dL_dT01 = 2 * t0v1 * dL_da + t1v1 * dL_db,
                dL_dT02 = 2 * t0v2 * dL_da + t1v2 * dL_db;
    const float dL_dT10 = 2 * t1v0 * dL_dc + t0v0 * dL_db, dL_dT11 = 2 * t1v1 * dL_dc + t0v1 * dL_db,
                dL_dT12 = 2 * t1v2 * dL_dc + t0v2 * dL_db;
    const float dL_dJ00 = v[0] * dL_dT00 + v[4] * dL_dT01 + v[8] * dL_dT02;
    const float dL_dJ02 = v[2] * dL_dT00 + v[6] * dL_dT01 + v[10] * dL_dT02;
    const float dL_dJ11 = v[1] * dL_dT10 + v[5] * dL_dT11 + v[9] * dL_dT12;
    const float dL_dJ12 = v[2] * dL_dT10 + v[6] * dL_dT11 + v[10] * dL_dT12;
    const float tz = 1.f / in.tz, tz2 = tz * tz, tz3 = tz2 * tz;
    const float dL_dtx = x_grad_mul * -p.focal_x * tz2 * dL_dJ02;
    const float dL_dty = y_grad_mul * -p.focal_y * tz2 * dL_dJ12;
    const float dL_dtz = -p.focal_x * tz2 * dL_dJ00 - p.focal_y * tz2 * dL_dJ11 +
                         (2 * p.focal_x * in.tx) * tz3 * dL_dJ02 + (2 * p.focal_y * in.ty) * tz3 * dL_dJ12;
    // transformVec4x3Transpose (auxiliary.h:89-97); assignment, backward.cu:273
    dmx = v[0] * dL_dtx + v[1] * dL_dty + v[2] * dL_dtz;
    dmy = v[4] * dL_dtx + v[5] * dL_dty + v[6] * dL_dtz;
    dmz = v[8] * dL_dtx + v[9] * dL_dty + v[10] * dL_dtz;
    // ---- projection part (backward.cu:346-396) ----
    const float* pr = p.projmatrix;
    const float m_homw = pr[3] * mx + pr[7] * my + pr[11] * mz + pr[15];
    const float m_w = 1.0f / (m_homw + 0.0000001f);
    const float mul1 = (pr[0] * mx + pr[4] * my + pr[8] * mz + pr[12]) * m_w * m_w;
    const float mul2 = (pr[1] * mx + pr[5] * my + pr[9] * mz + pr[13]) * m_w * m_w;
    dmx += (pr[0] * m_w - pr[3] * mul1) * a[0] + (pr[1] * m_w - pr[3] * mul2) * a[1];
    dmy += (pr[4] * m_w - pr[7] * mul1) * a[0] + (pr[5] * m_w - pr[7] * mul2) * a[1];
    dmz += (pr[8] * m_w - pr[11] * mul1) * a[0] + (pr[9] * m_w - pr[11] * mul2) * a[1];
    // ---- SH (backward.cu:20-139) ----
    if (SH && p.shs) {
        sh_backward(p.D, p.M, p.shs + i * p.M * 3, dL_dsh ? dL_dsh + i * p.M * 3 : nullptr, mx, my, mz, p.cam_pos,
                    &a[6], clamped, dmx, dmy, dmz);
    }
    // ---- computeCov3D backward (backward.cu:278-341) ----
    if (!p.cov3D_precomp) {
        const float r = qr, x = qx, y = qy, z = qz;
        // Rm[c][r] as glm stores it (columns of R)
        const float Rm[3][3] = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                                {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                                {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
        const float s[3] = {p.scale_modifier * s0, p.scale_modifier * s1, p.scale_modifier * s2};
        float Mm[3][3];
#pragma unroll
        for (int c2 = 0; c2 < 3; c2++)
#pragma unroll
            for (int r2 = 0; r2 < 3; r2++) Mm[c2][r2] = s[r2] * Rm[c2][r2];
        const float dS[3][3] = {{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]},
                                {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]},
                                {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}};
        float dMt[3][3];  // transpose of dL_dM = 2 * M * dL_dSigma
#pragma unroll
        for (int c2 = 0; c2 < 3; c2++)
#pragma unroll
            for (int r2 = 0; r2 < 3; r2++) {
                float acc2 = 0.f;
#pragma unroll
                for (int k = 0; k < 3; k++) acc2 += (2.0f * Mm[k][r2]) * dS[c2][k];
                dMt[r2][c2] = acc2;
            }
#pragma unroll
        for (int k = 0; k < 3; k++) {
            float acc2 = 0.f;
#pragma unroll
            for (int r2 = 0; r2 < 3; r2++) acc2 += Rm[r2][k] * dMt[k][r2];
            dscale[k] = acc2;
        }
#pragma unroll
        for (int k = 0; k < 3; k++)
#pragma unroll
            for (int r2 = 0; r2 < 3; r2++) dMt[k][r2] *= s[k];
        drot[0] = 2 * z * (dMt[0][1] - dMt[1][0]) + 2 * y * (dMt[2][0] - dMt[0][2]) + 2 * x * (dMt[1][2] - dMt[2][1]);
        drot[1] = 2 * y * (dMt[1][0] + dMt[0][1]) + 2 * z * (dMt[2][0] + dMt[0][2]) + 2 * r * (dMt[1][2] - dMt[2][1]) -
                  4 * x * (dMt[2][2] + dMt[1][1]);
        drot[2] = 2 * x * (dMt[1][0] + dMt[0][1]) + 2 * r * (dMt[2][0] - dMt[0][2]) + 2 * z * (dMt[1][2] + dMt[2][1]) -
                  4 * y * (dMt[2][2] + dMt[0][0]);
        drot[3] = 2 * r * (dMt[0][1] - dMt[1][0]) + 2 * x * (dMt[2][0] + dMt[0][2]) + 2 * y * (dMt[1][2] + dMt[2][1]) -
                  4 * z * (dMt[1][1] + dMt[0][0]);
    }
}

}  // namespace gsb
