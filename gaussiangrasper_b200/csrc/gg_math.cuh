// Per-Gaussian math of the projection / SH stages, shared by the CUDA kernels and by the
// host-side unit-test shim (tests/hostmath.cu).  Everything here is fp32 with an explicit
// operation order; project.cu is compiled with -fmad=false so that radii / num_tiles_hit /
// depth bits are reproducible bit for bit against the CPU oracle.
//
// Reference call sites: nerfstudio/models/gaussian_splatting.py:699-713 (ProjectGaussians),
// :730 (SphericalHarmonics).  Algorithm: gsplat 0.1.0 as restated in SURVEY.md Appendix A.
#pragma once
#include <math.h>
#include <stdint.h>

#ifndef GG_TILE
#define GG_TILE 16
#endif

#ifdef __CUDACC__
#define GG_HD __host__ __device__ __forceinline__
#else
#define GG_HD inline
#endif

namespace gg {

struct Camera {
    float vm[12];  // world->camera, row-major 3x4
    float fm[16];  // projmat @ viewmat, row-major 4x4
    float fx, fy, cx, cy;
};

GG_HD int f2i_rz_sat(float x) {
#ifdef __CUDA_ARCH__
    return __float2int_rz(x);  // saturating, NaN -> 0
#else
    if (x != x) return 0;
    if (x >= 2147483648.0f) return INT32_MAX;
    if (x <= -2147483648.0f) return INT32_MIN;
    return (int)x;
#endif
}

GG_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

struct TileBox {
    int x0, y0, x1, y1;
    GG_HD int area() const { return (x1 - x0) * (y1 - y0); }
};

// A6: tiles touched by the disc (centre, radius); truncation toward zero happens before clamping.
GG_HD TileBox tile_box(float ux, float uy, float radius, int tiles_x, int tiles_y) {
    const float inv = (float)GG_TILE;
    const float tcx = ux / inv, tcy = uy / inv, tr = radius / inv;
    TileBox b;
    b.x0 = clampi(f2i_rz_sat(tcx - tr), 0, tiles_x);
    b.x1 = clampi(f2i_rz_sat(tcx + tr + 1.0f), 0, tiles_x);
    b.y0 = clampi(f2i_rz_sat(tcy - tr), 0, tiles_y);
    b.y1 = clampi(f2i_rz_sat(tcy + tr + 1.0f), 0, tiles_y);
    return b;
}

struct Rot3 {
    float m[9];
};

// A2: unit quaternion (w,x,y,z) -> rotation, normalising with 1/sqrt (not rsqrt).
GG_HD Rot3 quat_to_rot(float qw, float qx, float qy, float qz, float* inv_norm = nullptr) {
    const float inv = 1.0f / sqrtf(qw * qw + qx * qx + qy * qy + qz * qz);
    if (inv_norm) *inv_norm = inv;
    const float w = qw * inv, x = qx * inv, y = qy * inv, z = qz * inv;
    Rot3 R;
    R.m[0] = 1.0f - 2.0f * (y * y + z * z);
    R.m[1] = 2.0f * (x * y - w * z);
    R.m[2] = 2.0f * (x * z + w * y);
    R.m[3] = 2.0f * (x * y + w * z);
    R.m[4] = 1.0f - 2.0f * (x * x + z * z);
    R.m[5] = 2.0f * (y * z - w * x);
    R.m[6] = 2.0f * (x * z - w * y);
    R.m[7] = 2.0f * (y * z + w * x);
    R.m[8] = 1.0f - 2.0f * (x * x + y * y);
    return R;
}

struct ProjOut {
    float cov3d[6];
    float ux, uy, depth;
    float conic[3];
    int radius, tiles;
    bool cov3d_valid;
};

// A1..A6 for one Gaussian.  Outputs are all-zero when the Gaussian is culled (cov3d is kept
// once the near-plane test passed, as the upstream kernel writes it before the later culls).
GG_HD ProjOut project_one(const float p[3], const float s[3], float glob, const float q[4], const Camera& cam,
                          int img_h, int img_w, int tiles_x, int tiles_y, float clip) {
    ProjOut o;
#pragma unroll
    for (int k = 0; k < 6; ++k) o.cov3d[k] = 0.0f;
    o.ux = o.uy = o.depth = 0.0f;
    o.conic[0] = o.conic[1] = o.conic[2] = 0.0f;
    o.radius = 0;
    o.tiles = 0;
    o.cov3d_valid = false;
    const float* vm = cam.vm;
    const float* fm = cam.fm;
    float tx = vm[0] * p[0] + vm[1] * p[1] + vm[2] * p[2] + vm[3];
    float ty = vm[4] * p[0] + vm[5] * p[1] + vm[6] * p[2] + vm[7];
    const float tz = vm[8] * p[0] + vm[9] * p[1] + vm[10] * p[2] + vm[11];
    if (tz <= clip) return o;

    const Rot3 R = quat_to_rot(q[0], q[1], q[2], q[3]);
    const float sx = glob * s[0], sy = glob * s[1], sz = glob * s[2];
    const float M00 = R.m[0] * sx, M01 = R.m[1] * sy, M02 = R.m[2] * sz;
    const float M10 = R.m[3] * sx, M11 = R.m[4] * sy, M12 = R.m[5] * sz;
    const float M20 = R.m[6] * sx, M21 = R.m[7] * sy, M22 = R.m[8] * sz;
    const float V00 = M00 * M00 + M01 * M01 + M02 * M02;
    const float V01 = M00 * M10 + M01 * M11 + M02 * M12;
    const float V02 = M00 * M20 + M01 * M21 + M02 * M22;
    const float V11 = M10 * M10 + M11 * M11 + M12 * M12;
    const float V12 = M10 * M20 + M11 * M21 + M12 * M22;
    const float V22 = M20 * M20 + M21 * M21 + M22 * M22;
    o.cov3d[0] = V00; o.cov3d[1] = V01; o.cov3d[2] = V02;
    o.cov3d[3] = V11; o.cov3d[4] = V12; o.cov3d[5] = V22;
    o.cov3d_valid = true;

    const float limx = 1.3f * (0.5f * (float)img_w / cam.fx);
    const float limy = 1.3f * (0.5f * (float)img_h / cam.fy);
    tx = tz * fminf(limx, fmaxf(-limx, tx / tz));
    ty = tz * fminf(limy, fmaxf(-limy, ty / tz));
    const float rz = 1.0f / tz, rz2 = rz * rz;
    const float J00 = cam.fx * rz, J02 = -cam.fx * tx * rz2;
    const float J11 = cam.fy * rz, J12 = -cam.fy * ty * rz2;
    const float T00 = J00 * vm[0] + J02 * vm[8], T01 = J00 * vm[1] + J02 * vm[9], T02 = J00 * vm[2] + J02 * vm[10];
    const float T10 = J11 * vm[4] + J12 * vm[8], T11 = J11 * vm[5] + J12 * vm[9], T12 = J11 * vm[6] + J12 * vm[10];
    const float U00 = T00 * V00 + T01 * V01 + T02 * V02;
    const float U01 = T00 * V01 + T01 * V11 + T02 * V12;
    const float U02 = T00 * V02 + T01 * V12 + T02 * V22;
    const float U10 = T10 * V00 + T11 * V01 + T12 * V02;
    const float U11 = T10 * V01 + T11 * V11 + T12 * V12;
    const float U12 = T10 * V02 + T11 * V12 + T12 * V22;
    const float a = (U00 * T00 + U01 * T01 + U02 * T02) + 0.3f;
    const float b = U00 * T10 + U01 * T11 + U02 * T12;
    const float c = (U10 * T10 + U11 * T11 + U12 * T12) + 0.3f;

    const float det = a * c - b * b;
    if (det == 0.0f) return o;
    const float inv_det = 1.0f / det;
    const float mid = 0.5f * (a + c);
    const float sq = sqrtf(fmaxf(0.1f, mid * mid - det));
    const float v1 = mid + sq, v2 = mid - sq;
    const float radius = ceilf(3.0f * sqrtf(fmaxf(v1, v2)));

    const float hx = fm[0] * p[0] + fm[1] * p[1] + fm[2] * p[2] + fm[3];
    const float hy = fm[4] * p[0] + fm[5] * p[1] + fm[6] * p[2] + fm[7];
    const float hw = fm[12] * p[0] + fm[13] * p[1] + fm[14] * p[2] + fm[15];
    const float rw = 1.0f / (hw + 1e-6f);
    const float ux = 0.5f * (float)img_w * (hx * rw) + cam.cx - 0.5f;
    const float uy = 0.5f * (float)img_h * (hy * rw) + cam.cy - 0.5f;

    const TileBox tb = tile_box(ux, uy, radius, tiles_x, tiles_y);
    const int area = tb.area();
    if (area <= 0) return o;
    o.tiles = area;
    o.depth = tz;
    o.radius = f2i_rz_sat(radius);
    o.ux = ux;
    o.uy = uy;
    o.conic[0] = c * inv_det;
    o.conic[1] = -b * inv_det;
    o.conic[2] = a * inv_det;
    return o;
}

struct ProjGrad {
    float v_mean[3], v_scale[3], v_quat[4];
};

// A11 (projection part): exact vector-Jacobian product of project_one for a Gaussian that was
// not culled.  v_R_extra (optional, 9 floats) is added to the gradient of the rotation matrix.  v_conic is the gradient w.r.t. the stored (A, B, C) triple, B being the single
// off-diagonal entry.
GG_HD ProjGrad project_bwd_one(const float p[3], const float s[3], float glob, const float q[4], const Camera& cam,
                               int img_h, int img_w, const float conic[3], const float v_xy[2], float v_depth,
                               const float v_conic[3], const float* v_R_extra = nullptr) {
    ProjGrad g;
    const float* vm = cam.vm;
    const float* fm = cam.fm;
    // ---- recompute forward intermediates ----
    const float tx0 = vm[0] * p[0] + vm[1] * p[1] + vm[2] * p[2] + vm[3];
    const float ty0 = vm[4] * p[0] + vm[5] * p[1] + vm[6] * p[2] + vm[7];
    const float tz = vm[8] * p[0] + vm[9] * p[1] + vm[10] * p[2] + vm[11];
    float inv_qn;
    const Rot3 R = quat_to_rot(q[0], q[1], q[2], q[3], &inv_qn);
    const float sc[3] = {glob * s[0], glob * s[1], glob * s[2]};
    float M[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) M[3 * r + c] = R.m[3 * r + c] * sc[c];
    float V[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) V[3 * r + c] = M[3 * r] * M[3 * c] + M[3 * r + 1] * M[3 * c + 1] + M[3 * r + 2] * M[3 * c + 2];
    const float limx = 1.3f * (0.5f * (float)img_w / cam.fx);
    const float limy = 1.3f * (0.5f * (float)img_h / cam.fy);
    const float rxz = tx0 / tz, ryz = ty0 / tz;
    const bool clx = (rxz < -limx) || (rxz > limx);
    const bool cly = (ryz < -limy) || (ryz > limy);
    const float cxr = fminf(limx, fmaxf(-limx, rxz));
    const float cyr = fminf(limy, fmaxf(-limy, ryz));
    const float tx = tz * cxr, ty = tz * cyr;
    const float rz = 1.0f / tz, rz2 = rz * rz, rz3 = rz2 * rz;
    const float J00 = cam.fx * rz, J02 = -cam.fx * tx * rz2;
    const float J11 = cam.fy * rz, J12 = -cam.fy * ty * rz2;
    float T[6];
    T[0] = J00 * vm[0] + J02 * vm[8]; T[1] = J00 * vm[1] + J02 * vm[9]; T[2] = J00 * vm[2] + J02 * vm[10];
    T[3] = J11 * vm[4] + J12 * vm[8]; T[4] = J11 * vm[5] + J12 * vm[9]; T[5] = J11 * vm[6] + J12 * vm[10];

    // ---- conic -> cov2d : Gs = -X Gm X with X the conic matrix ----
    const float A = conic[0], B = conic[1], C = conic[2];
    const float hB = 0.5f * v_conic[1];
    const float XG00 = A * v_conic[0] + B * hB, XG01 = A * hB + B * v_conic[2];
    const float XG10 = B * v_conic[0] + C * hB, XG11 = B * hB + C * v_conic[2];
    const float Gs00 = -(XG00 * A + XG01 * B);
    const float Gs01 = -(XG00 * B + XG01 * C);
    const float Gs11 = -(XG10 * B + XG11 * C);
    // ---- cov2d = T V T^T : Gv = T^T Gs T (3x3 sym), v_T = 2 Gs T V ----
    float GT[6];  // Gs * T (2x3)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        GT[c] = Gs00 * T[c] + Gs01 * T[3 + c];
        GT[3 + c] = Gs01 * T[c] + Gs11 * T[3 + c];
    }
    float Gv[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) Gv[3 * r + c] = T[r] * GT[c] + T[3 + r] * GT[3 + c];
    float vT[6];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            vT[3 * i + c] = 2.0f * (GT[3 * i] * V[c] + GT[3 * i + 1] * V[3 + c] + GT[3 * i + 2] * V[6 + c]);
    // ---- V = M M^T : v_M = 2 Gv M ; M = R diag(sc) ----
    float vM[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            vM[3 * r + c] = 2.0f * (Gv[3 * r] * M[c] + Gv[3 * r + 1] * M[3 + c] + Gv[3 * r + 2] * M[6 + c]);
    float vR[9];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        g.v_scale[c] = glob * (R.m[c] * vM[c] + R.m[3 + c] * vM[3 + c] + R.m[6 + c] * vM[6 + c]);
#pragma unroll
        for (int r = 0; r < 3; ++r) vR[3 * r + c] = vM[3 * r + c] * sc[c];
    }
    if (v_R_extra) {  // gradient that reaches R directly (the normal channel reads a column of R)
#pragma unroll
        for (int k = 0; k < 9; ++k) vR[k] += v_R_extra[k];
    }
    // ---- R -> unit quaternion -> raw quaternion ----
    const float w = q[0] * inv_qn, x = q[1] * inv_qn, y = q[2] * inv_qn, z = q[3] * inv_qn;
    const float vw = 2.0f * (-z * vR[1] + y * vR[2] + z * vR[3] - x * vR[5] - y * vR[6] + x * vR[7]);
    const float vx = 2.0f * (y * vR[1] + z * vR[2] + y * vR[3] - 2.0f * x * vR[4] - w * vR[5] + z * vR[6] + w * vR[7] -
                             2.0f * x * vR[8]);
    const float vy = 2.0f * (-2.0f * y * vR[0] + x * vR[1] + w * vR[2] + x * vR[3] + z * vR[5] - w * vR[6] + z * vR[7] -
                             2.0f * y * vR[8]);
    const float vz = 2.0f * (-2.0f * z * vR[0] - w * vR[1] + x * vR[2] + w * vR[3] - 2.0f * z * vR[4] + y * vR[5] +
                             x * vR[6] + y * vR[7]);
    const float dotq = w * vw + x * vx + y * vy + z * vz;
    g.v_quat[0] = (vw - w * dotq) * inv_qn;
    g.v_quat[1] = (vx - x * dotq) * inv_qn;
    g.v_quat[2] = (vy - y * dotq) * inv_qn;
    g.v_quat[3] = (vz - z * dotq) * inv_qn;
    // ---- T = J W : v_J = v_T W^T ----
    const float vJ00 = vT[0] * vm[0] + vT[1] * vm[1] + vT[2] * vm[2];
    const float vJ02 = vT[0] * vm[8] + vT[1] * vm[9] + vT[2] * vm[10];
    const float vJ11 = vT[3] * vm[4] + vT[4] * vm[5] + vT[5] * vm[6];
    const float vJ12 = vT[3] * vm[8] + vT[4] * vm[9] + vT[5] * vm[10];
    const float v_txc = -cam.fx * rz2 * vJ02;
    const float v_tyc = -cam.fy * rz2 * vJ12;
    float v_tz = -cam.fx * rz2 * vJ00 - cam.fy * rz2 * vJ11 + 2.0f * cam.fx * tx * rz3 * vJ02 +
                 2.0f * cam.fy * ty * rz3 * vJ12;
    float v_tx = 0.0f, v_ty = 0.0f;
    if (clx) v_tz += cxr * v_txc; else v_tx = v_txc;
    if (cly) v_tz += cyr * v_tyc; else v_ty = v_tyc;
    v_tz += v_depth;
    // ---- pixel centre through the full matrix ----
    const float hx = fm[0] * p[0] + fm[1] * p[1] + fm[2] * p[2] + fm[3];
    const float hy = fm[4] * p[0] + fm[5] * p[1] + fm[6] * p[2] + fm[7];
    const float hw = fm[12] * p[0] + fm[13] * p[1] + fm[14] * p[2] + fm[15];
    const float rw = 1.0f / (hw + 1e-6f);
    const float kx = 0.5f * (float)img_w * v_xy[0], ky = 0.5f * (float)img_h * v_xy[1];
    const float v_hx = kx * rw, v_hy = ky * rw;
    const float v_hw = -(rw * rw) * (kx * hx + ky * hy);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        g.v_mean[c] = (vm[c] * v_tx + vm[4 + c] * v_ty + vm[8 + c] * v_tz) +
                      (fm[c] * v_hx + fm[4 + c] * v_hy + fm[12 + c] * v_hw);
    }
    return g;
}

// ---------------------------------------------------------------------------------------------
// A10: real spherical-harmonics basis in the 3DGS ordering.  Y must hold (deg+1)^2 floats.
// ---------------------------------------------------------------------------------------------
GG_HD int sh_num_bases(int degree) {
    return degree == 0 ? 1 : degree == 1 ? 4 : degree == 2 ? 9 : degree == 3 ? 16 : 25;
}

GG_HD void sh_basis(int deg, float dx, float dy, float dz, float* Y) {
    Y[0] = 0.28209479177387814f;
    if (deg < 1) return;
    const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
    const float x = dx / nrm, y = dy / nrm, z = dz / nrm;
    const float C1 = 0.4886025119029199f;
    Y[1] = C1 * (-y);
    Y[2] = C1 * z;
    Y[3] = C1 * (-x);
    if (deg < 2) return;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    Y[4] = 1.0925484305920792f * xy;
    Y[5] = -1.0925484305920792f * yz;
    Y[6] = 0.31539156525252005f * (2.0f * zz - xx - yy);
    Y[7] = -1.0925484305920792f * xz;
    Y[8] = 0.5462742152960396f * (xx - yy);
    if (deg < 3) return;
    Y[9] = -0.5900435899266435f * y * (3.0f * xx - yy);
    Y[10] = 2.890611442640554f * xy * z;
    Y[11] = -0.4570457994644658f * y * (4.0f * zz - xx - yy);
    Y[12] = 0.3731763325901154f * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
    Y[13] = -0.4570457994644658f * x * (4.0f * zz - xx - yy);
    Y[14] = 1.445305721320277f * z * (xx - yy);
    Y[15] = -0.5900435899266435f * x * (xx - 3.0f * yy);
    if (deg < 4) return;
    Y[16] = 2.5033429417967046f * xy * (xx - yy);
    Y[17] = -1.7701307697799304f * yz * (3.0f * xx - yy);
    Y[18] = 0.9461746957575601f * xy * (7.0f * zz - 1.0f);
    Y[19] = -0.6690465435572892f * yz * (7.0f * zz - 3.0f);
    Y[20] = 0.10578554691520431f * (zz * (35.0f * zz - 30.0f) + 3.0f);
    Y[21] = -0.6690465435572892f * xz * (7.0f * zz - 3.0f);
    Y[22] = 0.47308734787878004f * (xx - yy) * (7.0f * zz - 1.0f);
    Y[23] = -1.7701307697799304f * xz * (xx - 3.0f * yy);
    Y[24] = 0.6258357354491761f * (xx * (xx - 3.0f * yy) - yy * (3.0f * xx - yy));
}

}  // namespace gg
