/*
 * gg_oracle.c -- CPU restatement of the gsplat 0.1.0 rasterization path that
 * GaussianGrasper drives from nerfstudio/models/gaussian_splatting.py:699-784.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (gaussiangrasper_b200/,
 * gsplat/) may import, link or execute this file.  Only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() use it, and
 * only as the checker.
 *
 * PARITY UNPINNED: the arithmetic of this path lives in the third-party wheel
 * gsplat==0.1.0 (reference requirements.txt:78), which is neither vendored under
 * /root/reference nor installable here, and the reference's tests hold no golden
 * vector for it.  This file restates gsplat 0.1.0's
 * published algorithm and is anchored on the
 * reference's call sites:
 *   ProjectGaussians.apply      gaussian_splatting.py:699-713
 *   SphericalHarmonics.apply    gaussian_splatting.py:730
 *   RasterizeGaussians.apply    gaussian_splatting.py:735,759,773
 *   NDRasterizeGaussians.apply  gaussian_splatting.py:747
 * Two pieces of it ARE held by the reference and are checked against it
 * (tests/test_reference_golden_cpu.py): the 25 spherical-harmonics basis functions
 * (nerfstudio/utils/math.py:29-92, equal up to gsplat's (-1)^|m| sign) and the
 * (w, x, y, z) quaternion -> rotation convention behind the 3D covariance
 * (nerfstudio/cameras/camera_utils.py:142-161).
 *
 * Floating point discipline: every fp32 expression below is written with an
 * explicit left-to-right operation order and must be compiled with
 * -ffp-contract=off (no FMA contraction) so that the integer outputs (radii,
 * num_tiles_hit, sort keys, tile ranges) are reproducible bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define GG_TILE 16

static inline int f2i_sat(float x) {
    /* float -> int32 truncation toward zero, saturating, NaN -> 0 */
    if (x != x) return 0;
    if (x >= 2147483648.0f) return INT32_MAX;
    if (x <= -2147483648.0f) return INT32_MIN;
    return (int)x;
}
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* A6: tile bounding box of a disc (centre, radius) in 16x16 tiles. */
static void tile_bbox(float cx, float cy, float radius, int tiles_x, int tiles_y,
                      int *minx, int *miny, int *maxx, int *maxy) {
    float tcx = cx / (float)GG_TILE, tcy = cy / (float)GG_TILE;
    float tr = radius / (float)GG_TILE;
    *minx = imin(imax(0, f2i_sat(tcx - tr)), tiles_x);
    *maxx = imin(imax(0, f2i_sat(tcx + tr + 1.0f)), tiles_x);
    *miny = imin(imax(0, f2i_sat(tcy - tr)), tiles_y);
    *maxy = imin(imax(0, f2i_sat(tcy + tr + 1.0f)), tiles_y);
}

/* ------------------------------------------------------------------------- */
/* A1..A6: EWA projection (ProjectGaussians.forward, gaussian_splatting.py:699) */
/* ------------------------------------------------------------------------- */
void gg_oracle_project_fwd(int n, const float *means, const float *scales, float glob_scale,
                           const float *quats, const float *vm /*>=12, row-major 3x4*/,
                           const float *fm /*16*/, float fx, float fy, float cx, float cy,
                           int img_h, int img_w, int tiles_x, int tiles_y, float clip,
                           float *cov3d, float *xys, float *depths, int32_t *radii,
                           float *conics, int32_t *num_tiles_hit) {
    memset(cov3d, 0, sizeof(float) * 6 * (size_t)n);
    memset(xys, 0, sizeof(float) * 2 * (size_t)n);
    memset(depths, 0, sizeof(float) * (size_t)n);
    memset(radii, 0, sizeof(int32_t) * (size_t)n);
    memset(conics, 0, sizeof(float) * 3 * (size_t)n);
    memset(num_tiles_hit, 0, sizeof(int32_t) * (size_t)n);
    const float tan_fovx = 0.5f * (float)img_w / fx;
    const float tan_fovy = 0.5f * (float)img_h / fy;
    const float limx = 1.3f * tan_fovx, limy = 1.3f * tan_fovy;
    for (int i = 0; i < n; ++i) {
        const float px = means[3 * i], py = means[3 * i + 1], pz = means[3 * i + 2];
        /* A1 near-plane clip */
        float tx = vm[0] * px + vm[1] * py + vm[2] * pz + vm[3];
        float ty = vm[4] * px + vm[5] * py + vm[6] * pz + vm[7];
        const float tz = vm[8] * px + vm[9] * py + vm[10] * pz + vm[11];
        if (tz <= clip) continue;
        /* A2 cov3d = (R S)(R S)^T */
        const float qw = quats[4 * i], qx = quats[4 * i + 1], qy = quats[4 * i + 2], qz = quats[4 * i + 3];
        const float inv = 1.0f / sqrtf(qw * qw + qx * qx + qy * qy + qz * qz);
        const float w = qw * inv, x = qx * inv, y = qy * inv, z = qz * inv;
        const float R00 = 1.0f - 2.0f * (y * y + z * z), R01 = 2.0f * (x * y - w * z), R02 = 2.0f * (x * z + w * y);
        const float R10 = 2.0f * (x * y + w * z), R11 = 1.0f - 2.0f * (x * x + z * z), R12 = 2.0f * (y * z - w * x);
        const float R20 = 2.0f * (x * z - w * y), R21 = 2.0f * (y * z + w * x), R22 = 1.0f - 2.0f * (x * x + y * y);
        const float sx = glob_scale * scales[3 * i], sy = glob_scale * scales[3 * i + 1], sz = glob_scale * scales[3 * i + 2];
        const float M00 = R00 * sx, M01 = R01 * sy, M02 = R02 * sz;
        const float M10 = R10 * sx, M11 = R11 * sy, M12 = R12 * sz;
        const float M20 = R20 * sx, M21 = R21 * sy, M22 = R22 * sz;
        const float V00 = M00 * M00 + M01 * M01 + M02 * M02;
        const float V01 = M00 * M10 + M01 * M11 + M02 * M12;
        const float V02 = M00 * M20 + M01 * M21 + M02 * M22;
        const float V11 = M10 * M10 + M11 * M11 + M12 * M12;
        const float V12 = M10 * M20 + M11 * M21 + M12 * M22;
        const float V22 = M20 * M20 + M21 * M21 + M22 * M22;
        float *cv = cov3d + 6 * (size_t)i;
        cv[0] = V00; cv[1] = V01; cv[2] = V02; cv[3] = V11; cv[4] = V12; cv[5] = V22;
        /* A3 EWA: cov2d = (J W) V (J W)^T + 0.3 I */
        tx = tz * fminf(limx, fmaxf(-limx, tx / tz));
        ty = tz * fminf(limy, fmaxf(-limy, ty / tz));
        const float rz = 1.0f / tz, rz2 = rz * rz;
        const float J00 = fx * rz, J02 = -fx * tx * rz2;
        const float J11 = fy * rz, J12 = -fy * ty * rz2;
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
        /* A4 conic and radius */
        const float det = a * c - b * b;
        if (det == 0.0f) continue;
        const float inv_det = 1.0f / det;
        const float mid = 0.5f * (a + c);
        const float sq = sqrtf(fmaxf(0.1f, mid * mid - det));
        const float v1 = mid + sq, v2 = mid - sq;
        const float radius = ceilf(3.0f * sqrtf(fmaxf(v1, v2)));
        /* A5 pixel centre through the full projection matrix */
        const float hx = fm[0] * px + fm[1] * py + fm[2] * pz + fm[3];
        const float hy = fm[4] * px + fm[5] * py + fm[6] * pz + fm[7];
        const float hw = fm[12] * px + fm[13] * py + fm[14] * pz + fm[15];
        const float rw = 1.0f / (hw + 1e-6f);
        const float ux = 0.5f * (float)img_w * (hx * rw) + cx - 0.5f;
        const float uy = 0.5f * (float)img_h * (hy * rw) + cy - 0.5f;
        /* A6 tile bbox */
        int minx, miny, maxx, maxy;
        tile_bbox(ux, uy, radius, tiles_x, tiles_y, &minx, &miny, &maxx, &maxy);
        const int area = (maxx - minx) * (maxy - miny);
        if (area <= 0) continue;
        num_tiles_hit[i] = area;
        depths[i] = tz;
        radii[i] = f2i_sat(radius);
        xys[2 * i] = ux; xys[2 * i + 1] = uy;
        conics[3 * i] = c * inv_det; conics[3 * i + 1] = -b * inv_det; conics[3 * i + 2] = a * inv_det;
    }
}

/* ------------------------------------------------------------------------- */
/* A10: spherical harmonics -> colour (SphericalHarmonics.forward, :730)       */
/* ------------------------------------------------------------------------- */
static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                               -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};
static const float SH_C4[9] = {2.5033429417967046f, -1.7701307697799304f, 0.9461746957575601f,
                               -0.6690465435572892f, 0.10578554691520431f, -0.6690465435572892f,
                               0.47308734787878004f, -1.7701307697799304f, 0.6258357354491761f};

int gg_oracle_num_sh_bases(int degree) {
    if (degree == 0) return 1;
    if (degree == 1) return 4;
    if (degree == 2) return 9;
    if (degree == 3) return 16;
    return 25;
}

/* basis values Y_b(d) for b < (deg+1)^2, direction normalised here */
static void sh_basis(int deg, float dx, float dy, float dz, float *Y) {
    Y[0] = SH_C0;
    if (deg < 1) return;
    const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
    const float x = dx / nrm, y = dy / nrm, z = dz / nrm;
    Y[1] = SH_C1 * (-y); Y[2] = SH_C1 * z; Y[3] = SH_C1 * (-x);
    if (deg < 2) return;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    Y[4] = SH_C2[0] * xy; Y[5] = SH_C2[1] * yz; Y[6] = SH_C2[2] * (2.0f * zz - xx - yy);
    Y[7] = SH_C2[3] * xz; Y[8] = SH_C2[4] * (xx - yy);
    if (deg < 3) return;
    Y[9] = SH_C3[0] * y * (3.0f * xx - yy);
    Y[10] = SH_C3[1] * xy * z;
    Y[11] = SH_C3[2] * y * (4.0f * zz - xx - yy);
    Y[12] = SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
    Y[13] = SH_C3[4] * x * (4.0f * zz - xx - yy);
    Y[14] = SH_C3[5] * z * (xx - yy);
    Y[15] = SH_C3[6] * x * (xx - 3.0f * yy);
    if (deg < 4) return;
    Y[16] = SH_C4[0] * xy * (xx - yy);
    Y[17] = SH_C4[1] * yz * (3.0f * xx - yy);
    Y[18] = SH_C4[2] * xy * (7.0f * zz - 1.0f);
    Y[19] = SH_C4[3] * yz * (7.0f * zz - 3.0f);
    Y[20] = SH_C4[4] * (zz * (35.0f * zz - 30.0f) + 3.0f);
    Y[21] = SH_C4[5] * xz * (7.0f * zz - 3.0f);
    Y[22] = SH_C4[6] * (xx - yy) * (7.0f * zz - 1.0f);
    Y[23] = SH_C4[7] * xz * (xx - 3.0f * yy);
    Y[24] = SH_C4[8] * (xx * (xx - 3.0f * yy) - yy * (3.0f * xx - yy));
}

void gg_oracle_sh_fwd(int n, int degree, int degrees_to_use, const float *dirs, const float *coeffs,
                      float *colors) {
    const int nb = gg_oracle_num_sh_bases(degree);
    const int nuse = gg_oracle_num_sh_bases(degrees_to_use);
    for (int i = 0; i < n; ++i) {
        float Y[25];
        sh_basis(degrees_to_use, dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2], Y);
        const float *cf = coeffs + (size_t)i * nb * 3;
        for (int c = 0; c < 3; ++c) {
            float acc = 0.0f;
            for (int b = 0; b < nuse; ++b) acc = acc + Y[b] * cf[3 * b + c];
            colors[3 * i + c] = acc;
        }
    }
}

void gg_oracle_sh_bwd(int n, int degree, int degrees_to_use, const float *dirs, const float *v_colors,
                      float *v_coeffs) {
    const int nb = gg_oracle_num_sh_bases(degree);
    const int nuse = gg_oracle_num_sh_bases(degrees_to_use);
    memset(v_coeffs, 0, sizeof(float) * (size_t)n * nb * 3);
    for (int i = 0; i < n; ++i) {
        float Y[25];
        sh_basis(degrees_to_use, dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2], Y);
        float *vc = v_coeffs + (size_t)i * nb * 3;
        for (int b = 0; b < nuse; ++b)
            for (int c = 0; c < 3; ++c) vc[3 * b + c] = Y[b] * v_colors[3 * i + c];
    }
}

/* ------------------------------------------------------------------------- */
/* A7/A8: binning (compute_cumulative_intersects + bin_and_sort_gaussians)     */
/* ------------------------------------------------------------------------- */
void gg_oracle_cumsum(int n, const int32_t *num_tiles_hit, int32_t *cum) {
    int32_t acc = 0;
    for (int i = 0; i < n; ++i) { acc += num_tiles_hit[i]; cum[i] = acc; }
}

/* A6 alone: tiles touched by (centre, integer radius) -- what the projection reports as num_tiles_hit; lets the
 * binning tests make up projection outputs directly (0 for radius <= 0 or an empty box). */
void gg_oracle_tile_counts(int n, const float *xys, const int32_t *radii, int tiles_x, int tiles_y,
                           int32_t *num_tiles_hit) {
    for (int i = 0; i < n; ++i) {
        num_tiles_hit[i] = 0;
        if (radii[i] <= 0) continue;
        int minx, miny, maxx, maxy;
        tile_bbox(xys[2 * i], xys[2 * i + 1], (float)radii[i], tiles_x, tiles_y, &minx, &miny, &maxx, &maxy);
        const int area = (maxx - minx) * (maxy - miny);
        num_tiles_hit[i] = area > 0 ? area : 0;
    }
}

/* keys/ids must hold cum[n-1] entries */
void gg_oracle_map_to_intersects(int n, const float *xys, const float *depths, const int32_t *radii,
                                 const int32_t *cum, int tiles_x, int tiles_y, int64_t *keys,
                                 int32_t *ids) {
    for (int i = 0; i < n; ++i) {
        if (radii[i] <= 0) continue;
        int minx, miny, maxx, maxy;
        tile_bbox(xys[2 * i], xys[2 * i + 1], (float)radii[i], tiles_x, tiles_y, &minx, &miny, &maxx, &maxy);
        int64_t cur = (i == 0) ? 0 : cum[i - 1];
        int32_t dbits;
        memcpy(&dbits, &depths[i], 4);
        for (int ty = miny; ty < maxy; ++ty)
            for (int tx = minx; tx < maxx; ++tx) {
                const int64_t tile = (int64_t)ty * tiles_x + tx;
                keys[cur] = (tile << 32) | (int64_t)(uint32_t)dbits;
                ids[cur] = i;
                ++cur;
            }
    }
}

typedef struct { int64_t key; int32_t id; } kv_t;
static int kv_cmp(const void *a, const void *b) {
    const kv_t *x = (const kv_t *)a, *y = (const kv_t *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return (x->id > y->id) - (x->id < y->id); /* tie rule: ascending gaussian id */
}

void gg_oracle_sort(int64_t m, const int64_t *keys, const int32_t *ids, int64_t *keys_sorted,
                    int32_t *ids_sorted) {
    kv_t *kv = (kv_t *)malloc(sizeof(kv_t) * (size_t)(m > 0 ? m : 1));
    for (int64_t i = 0; i < m; ++i) { kv[i].key = keys[i]; kv[i].id = ids[i]; }
    qsort(kv, (size_t)m, sizeof(kv_t), kv_cmp);
    for (int64_t i = 0; i < m; ++i) { keys_sorted[i] = kv[i].key; ids_sorted[i] = kv[i].id; }
    free(kv);
}

/* tile_ranges: [num_tiles, 2] int32, (start, end), (0,0) for empty tiles */
void gg_oracle_tile_ranges(int64_t m, const int64_t *keys_sorted, int num_tiles, int32_t *tile_ranges) {
    memset(tile_ranges, 0, sizeof(int32_t) * 2 * (size_t)num_tiles);
    for (int64_t i = 0; i < m; ++i) {
        const int32_t t = (int32_t)(keys_sorted[i] >> 32);
        if (i == 0 || (int32_t)(keys_sorted[i - 1] >> 32) != t) tile_ranges[2 * t] = (int32_t)i;
        if (i == m - 1 || (int32_t)(keys_sorted[i + 1] >> 32) != t) tile_ranges[2 * t + 1] = (int32_t)(i + 1);
    }
}

/* ------------------------------------------------------------------------- */
/* A9: forward alpha compositing (rasterize_forward / nd_rasterize_forward)    */
/*                                                                             */
/* final_idx[p] = one past the index (in the sorted list) of the last entry    */
/* that contributed to pixel p, or the tile's range start if none did.         */
/* fragile[p] (optional) is set when any visited pair lies within `eps`        */
/* (relative) of a branch threshold, so that a 1-ulp difference in exp() could */
/* flip it; parity tests compare those pixels at a looser bound.               */
/* Returns the number of pixel-Gaussian pairs visited (K of SURVEY 8d).        */
/* ------------------------------------------------------------------------- */
/* T is a product of up to thousands of fp32 factors: two correct fp32 evaluations differ by sqrt(n) ulps in it,   */
/* not by one, so the window around the T = 1e-4 stop is GG_T_WINDOW times wider than the one around alpha.        */
#define GG_T_WINDOW 16.0f
int64_t gg_oracle_blend_fwd(int channels, int img_h, int img_w, int tiles_x, int tiles_y,
                            const int32_t *ids_sorted, const int32_t *tile_ranges, const float *xys,
                            const float *conics, const float *opac, const float *colors,
                            const float *bg, float *out, float *final_T, int32_t *final_idx,
                            uint8_t *fragile, float eps) {
    int64_t pairs = 0;
    (void)tiles_y;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : pairs)
    for (int py = 0; py < img_h; ++py) {
        for (int pxi = 0; pxi < img_w; ++pxi) {
            const int tile = (py / GG_TILE) * tiles_x + (pxi / GG_TILE);
            const int start = tile_ranges[2 * tile], end = tile_ranges[2 * tile + 1];
            const size_t pix = (size_t)py * img_w + pxi;
            float *o = out + pix * channels;
            for (int c = 0; c < channels; ++c) o[c] = 0.0f;
            float T = 1.0f;
            int last = start;
            uint8_t frag = 0;
            const float fx = (float)pxi, fy = (float)py;
            for (int k = start; k < end; ++k) {
                ++pairs;
                const int g = ids_sorted[k];
                const float dx = xys[2 * g] - fx, dy = xys[2 * g + 1] - fy;
                const float A = conics[3 * g], B = conics[3 * g + 1], C = conics[3 * g + 2];
                const float sigma = 0.5f * (A * dx * dx + C * dy * dy) + B * dx * dy;
                if (fabsf(sigma) <= 1e-6f) frag = 1;
                if (sigma < 0.0f) continue;
                const float araw = opac[g] * expf(-sigma);
                const float alpha = fminf(0.999f, araw);
                if (fabsf(alpha - (1.0f / 255.0f)) <= eps * (1.0f / 255.0f)) frag = 1;
                if (alpha < 1.0f / 255.0f) continue;
                const float next_T = T * (1.0f - alpha);
                if (fabsf(next_T - 1e-4f) <= GG_T_WINDOW * eps * 1e-4f) frag = 1;
                if (next_T <= 1e-4f) break;
                const float vis = alpha * T;
                const float *col = colors + (size_t)g * channels;
                for (int c = 0; c < channels; ++c) o[c] = o[c] + vis * col[c];
                T = next_T;
                last = k + 1;
            }
            for (int c = 0; c < channels; ++c) o[c] = o[c] + T * bg[c];
            final_T[pix] = T;
            final_idx[pix] = last;
            if (fragile) fragile[pix] = frag;
        }
    }
    return pairs;
}

/* ------------------------------------------------------------------------- */
/* A11: backward of the blend, analytic, accumulated in double.                */
/* Convention: exact derivative of A9; alpha clamped at      */
/* 0.999 passes no gradient to sigma/opacity.                                  */
/* v_xy[N,2], v_conic[N,3], v_colors[N,C], v_opac[N] are doubles.              */
/* ------------------------------------------------------------------------- */
/* Extended form.  Besides the gradients it can return, per gradient component, two error scales:              */
/*   abs_*  : the sum of the magnitudes the component is accumulated from -- per pixel, the contribution with    */
/*            every product of the alpha gradient taken in absolute value BEFORE it is summed over the channels  */
/*            (col T - S/(1-a) cancels inside a pair as well as between pixels), times 1 + (|A dx^2|/2 +          */
/*            |C dy^2|/2 + |B dx dy|): alpha = o exp(-sigma) inherits the ABSOLUTE fp32 error of sigma, which     */
/*            grows with the three products sigma is summed from.  fp32 evaluation and accumulation errors are    */
/*            proportional to this scale, not to |sum|.                                                           */
/*   taint_*: what flipped branches change.  A pixel with a test within its window of a threshold (alpha = 1/255  */
/*            and sigma = 0: `eps` relative / 1e-6 absolute; T = 1e-4: GG_T_WINDOW * eps, the forward's `fragile` */
/*            flag) is evaluated a second time with every such test inverted; taint accumulates, per pair and     */
/*            component, |contribution(flipped) - contribution(as decided)| -- the blended-or-not pair itself,    */
/*            the rescaling of what lies behind it and the shift of the sums in front of it.                      */
/* The parity tests bound |got - ref| <= rtol |ref| + k eps_fp32 abs + 2 taint, element by element.               */
/* abs / taint pointers may be NULL (all or none of each group).  Rows of pixels run in parallel; the            */
/* accumulations are atomic (fp64: the summation order does not matter at the tested tolerances).                */
#define GG_ACC(dst, val) do { const double gg_v_ = (val); _Pragma("omp atomic") (dst) += gg_v_; } while (0)

typedef struct {
    int channels, n;
    const int32_t *ids_sorted;
    const float *xys, *conics, *opac, *colors, *bg;
    float eps;
} gg_bwd_ctx;

/* One pixel: fp32 replay of A9 (with the near-threshold tests inverted when `flip`), then the fp64 backward over  */
/* the pairs the replay blended.  Per pair k - start: six geometry terms, `channels` colour terms, and (base pass   */
/* only) their magnitudes.  Returns 1 when some test sat inside its window.                                        */
static int gg_pixel_bwd(const gg_bwd_ctx *c, int start, int end, float fx, float fy, const float *vo, int flip,
                        unsigned char *inc, double *S, double *t_geo, double *t_col, double *a_geo, double *a_col) {
    const int ch = c->channels;
    float T = 1.0f;
    int last = start, frag = 0;
    for (int k = start; k < end; ++k) {
        inc[k - start] = 0;
        if (last < 0) continue; /* stopped: nothing behind is blended */
        const int g = c->ids_sorted[k];
        const float dx = c->xys[2 * g] - fx, dy = c->xys[2 * g + 1] - fy;
        const float A = c->conics[3 * g], B = c->conics[3 * g + 1], C = c->conics[3 * g + 2];
        const float sigma = 0.5f * (A * dx * dx + C * dy * dy) + B * dx * dy;
        int neg = sigma < 0.0f;
        if (fabsf(sigma) <= 1e-6f) { frag = 1; if (flip) neg = !neg; }
        if (neg) continue;
        const float alpha = fminf(0.999f, c->opac[g] * expf(-sigma));
        int low = alpha < 1.0f / 255.0f;
        if (fabsf(alpha - (1.0f / 255.0f)) <= c->eps * (1.0f / 255.0f)) { frag = 1; if (flip) low = !low; }
        if (low) continue;
        const float next_T = T * (1.0f - alpha);
        int stop = next_T <= 1e-4f;
        if (fabsf(next_T - 1e-4f) <= GG_T_WINDOW * c->eps * 1e-4f) { frag = 1; if (flip) stop = !stop; }
        if (stop) { last = -(last + 1); continue; }
        T = next_T;
        last = k + 1;
        inc[k - start] = 1;
    }
    if (last < 0) last = -last - 1;
    /* transmittance behind the last blended pair, in double */
    double Tfin = 1.0;
    for (int k = start; k < last; ++k) {
        if (!inc[k - start]) continue;
        const int g = c->ids_sorted[k];
        const double dx = (double)c->xys[2 * g] - fx, dy = (double)c->xys[2 * g + 1] - fy;
        const double sg = 0.5 * ((double)c->conics[3 * g] * dx * dx + (double)c->conics[3 * g + 2] * dy * dy) +
                          (double)c->conics[3 * g + 1] * dx * dy;
        double al = (double)c->opac[g] * exp(-sg);
        if (al > 0.999) al = 0.999;
        Tfin *= (1.0 - al);
    }
    double bgdot = 0.0, bgabs = 0.0;
    for (int q = 0; q < ch; ++q) {
        S[q] = 0.0;
        bgdot += (double)c->bg[q] * vo[q];
        bgabs += fabs((double)c->bg[q] * vo[q]);
    }
    double Tcur = Tfin; /* transmittance after entry k */
    for (int k = end - 1; k >= start; --k) {
        double *tg = t_geo + 6 * (size_t)(k - start), *tc = t_col + (size_t)ch * (k - start);
        for (int q = 0; q < 6; ++q) tg[q] = 0.0;
        for (int q = 0; q < ch; ++q) tc[q] = 0.0;
        if (a_geo) {
            for (int q = 0; q < 6; ++q) a_geo[6 * (size_t)(k - start) + q] = 0.0;
            for (int q = 0; q < ch; ++q) a_col[(size_t)ch * (k - start) + q] = 0.0;
        }
        if (!inc[k - start]) continue;
        const int g = c->ids_sorted[k];
        const double A = c->conics[3 * g], B = c->conics[3 * g + 1], C = c->conics[3 * g + 2];
        const double dx = (double)c->xys[2 * g] - fx, dy = (double)c->xys[2 * g + 1] - fy;
        const double sg = 0.5 * (A * dx * dx + C * dy * dy) + B * dx * dy;
        const double amp = 1.0 + 0.5 * fabs(A) * dx * dx + 0.5 * fabs(C) * dy * dy + fabs(B * dx * dy);
        const double vis = exp(-sg);
        double al = (double)c->opac[g] * vis;
        const int clamped = al > 0.999;
        if (clamped) al = 0.999;
        const double ra = 1.0 / (1.0 - al);
        const double Tb = Tcur * ra; /* transmittance before entry k */
        const double fac = al * Tb;
        const float *col = c->colors + (size_t)g * ch;
        double v_alpha = 0.0, v_alpha_abs = 0.0;
        for (int q = 0; q < ch; ++q) {
            tc[q] = fac * vo[q];
            if (a_col) a_col[(size_t)ch * (k - start) + q] = fabs(tc[q]) * amp;
            v_alpha += ((double)col[q] * Tb - S[q] * ra) * vo[q];
            v_alpha_abs += (fabs((double)col[q] * Tb) + fabs(S[q] * ra)) * fabs((double)vo[q]);
            S[q] += (double)col[q] * fac;
        }
        v_alpha += -Tfin * ra * bgdot;
        v_alpha_abs += Tfin * ra * bgabs;
        Tcur = Tb;
        if (clamped) continue; /* App. B-6: a clamped alpha passes no gradient to sigma / opacity */
        /* d alpha / d sigma = -o e^-s; the six factors multiply the alpha gradient */
        const double f6[6] = {-al * (A * dx + B * dy), -al * (B * dx + C * dy), -al * 0.5 * dx * dx, -al * dx * dy,
                              -al * 0.5 * dy * dy, vis};
        for (int q = 0; q < 6; ++q) {
            tg[q] = f6[q] * v_alpha;
            if (a_geo) a_geo[6 * (size_t)(k - start) + q] = fabs(f6[q]) * v_alpha_abs * amp;
        }
    }
    return frag;
}

void gg_oracle_blend_bwd_ex(int n, int channels, int img_h, int img_w, int tiles_x,
                            const int32_t *ids_sorted, const int32_t *tile_ranges, const float *xys,
                            const float *conics, const float *opac, const float *colors,
                            const float *bg, const float *v_out, double *v_xy, double *v_conic,
                            double *v_colors, double *v_opac, float eps, double *abs_geo /*[n,6]*/,
                            double *abs_colors /*[n,C]*/, double *taint_geo /*[n,6]*/,
                            double *taint_colors /*[n,C]*/) {
    memset(v_xy, 0, sizeof(double) * 2 * (size_t)n);
    memset(v_conic, 0, sizeof(double) * 3 * (size_t)n);
    memset(v_colors, 0, sizeof(double) * (size_t)channels * (size_t)n);
    memset(v_opac, 0, sizeof(double) * (size_t)n);
    if (abs_geo) memset(abs_geo, 0, sizeof(double) * 6 * (size_t)n);
    if (abs_colors) memset(abs_colors, 0, sizeof(double) * (size_t)channels * (size_t)n);
    if (taint_geo) memset(taint_geo, 0, sizeof(double) * 6 * (size_t)n);
    if (taint_colors) memset(taint_colors, 0, sizeof(double) * (size_t)channels * (size_t)n);
    const int n_tiles = tiles_x * ((img_h + GG_TILE - 1) / GG_TILE);
    int longest = 1;
    for (int t = 0; t < n_tiles; ++t)
        if (tile_ranges[2 * t + 1] - tile_ranges[2 * t] > longest) longest = tile_ranges[2 * t + 1] - tile_ranges[2 * t];
    const gg_bwd_ctx ctx = {channels, n, ids_sorted, xys, conics, opac, colors, bg, eps};
    const int want_abs = abs_geo != NULL, want_taint = taint_geo != NULL;
#pragma omp parallel
    {
    const size_t L = (size_t)longest;
    double *S = (double *)malloc(sizeof(double) * (size_t)channels);
    unsigned char *inc = (unsigned char *)malloc(L);
    double *t_geo = (double *)malloc(sizeof(double) * 6 * L), *t_col = (double *)malloc(sizeof(double) * channels * L);
    double *a_geo = want_abs ? (double *)malloc(sizeof(double) * 6 * L) : NULL;
    double *a_col = want_abs ? (double *)malloc(sizeof(double) * channels * L) : NULL;
    double *f_geo = want_taint ? (double *)malloc(sizeof(double) * 6 * L) : NULL;
    double *f_col = want_taint ? (double *)malloc(sizeof(double) * channels * L) : NULL;
#pragma omp for schedule(dynamic, 4)
    for (int py = 0; py < img_h; ++py) {
        for (int pxi = 0; pxi < img_w; ++pxi) {
            const int tile = (py / GG_TILE) * tiles_x + (pxi / GG_TILE);
            const int start = tile_ranges[2 * tile], end = tile_ranges[2 * tile + 1];
            if (end <= start) continue;
            const float *vo = v_out + ((size_t)py * img_w + pxi) * channels;
            const int frag = gg_pixel_bwd(&ctx, start, end, (float)pxi, (float)py, vo, 0, inc, S, t_geo, t_col, a_geo, a_col);
            for (int k = start; k < end; ++k) {
                if (!inc[k - start]) continue;
                const size_t g = (size_t)ids_sorted[k];
                const double *tg = t_geo + 6 * (size_t)(k - start), *tc = t_col + (size_t)channels * (k - start);
                GG_ACC(v_xy[2 * g], tg[0]); GG_ACC(v_xy[2 * g + 1], tg[1]);
                GG_ACC(v_conic[3 * g], tg[2]); GG_ACC(v_conic[3 * g + 1], tg[3]); GG_ACC(v_conic[3 * g + 2], tg[4]);
                GG_ACC(v_opac[g], tg[5]);
                for (int q = 0; q < channels; ++q) GG_ACC(v_colors[g * channels + q], tc[q]);
                if (want_abs) {
                    for (int q = 0; q < 6; ++q) GG_ACC(abs_geo[6 * g + q], a_geo[6 * (size_t)(k - start) + q]);
                    for (int q = 0; q < channels; ++q) GG_ACC(abs_colors[g * channels + q], a_col[(size_t)channels * (k - start) + q]);
                }
            }
            if (want_taint && frag) {
                gg_pixel_bwd(&ctx, start, end, (float)pxi, (float)py, vo, 1, inc, S, f_geo, f_col, NULL, NULL);
                for (int k = start; k < end; ++k) {
                    const size_t g = (size_t)ids_sorted[k], o6 = 6 * (size_t)(k - start), oc = (size_t)channels * (k - start);
                    for (int q = 0; q < 6; ++q) {
                        const double d = fabs(f_geo[o6 + q] - t_geo[o6 + q]);
                        if (d > 0.0) GG_ACC(taint_geo[6 * g + q], d);
                    }
                    for (int q = 0; q < channels; ++q) {
                        const double d = fabs(f_col[oc + q] - t_col[oc + q]);
                        if (d > 0.0) GG_ACC(taint_colors[g * channels + q], d);
                    }
                }
            }
        }
    }
    free(S); free(inc); free(t_geo); free(t_col); free(a_geo); free(a_col); free(f_geo); free(f_col);
    }
}
#undef GG_ACC

void gg_oracle_blend_bwd(int n, int channels, int img_h, int img_w, int tiles_x,
                         const int32_t *ids_sorted, const int32_t *tile_ranges, const float *xys,
                         const float *conics, const float *opac, const float *colors,
                         const float *bg, const float *v_out, double *v_xy, double *v_conic,
                         double *v_colors, double *v_opac) {
    gg_oracle_blend_bwd_ex(n, channels, img_h, img_w, tiles_x, ids_sorted, tile_ranges, xys, conics, opac, colors, bg,
                           v_out, v_xy, v_conic, v_colors, v_opac, 0.0f, NULL, NULL, NULL, NULL);
}

/* ------------------------------------------------------------------------- */
/* Workload statistics for DESIGN.md / bench roofline accounting (not parity). */
/* stats[0] = pairs visited, stats[1] = pairs that pass the alpha test         */
/* (contributors), stats[2] = sum over (tile, entry) of warps (16x2 pixel      */
/* strips) with at least one contributor, stats[3] = entries summed over tiles */
/* ------------------------------------------------------------------------- */
void gg_oracle_blend_stats(int img_h, int img_w, int tiles_x, int tiles_y, const int32_t *ids_sorted,
                           const int32_t *tile_ranges, const float *xys, const float *conics,
                           const float *opac, int64_t *stats) {
    int64_t visited = 0, hits = 0, warp_hits = 0, entries = 0;
    for (int ty = 0; ty < tiles_y; ++ty)
        for (int tx = 0; tx < tiles_x; ++tx) {
            const int tile = ty * tiles_x + tx;
            const int start = tile_ranges[2 * tile], end = tile_ranges[2 * tile + 1];
            entries += end - start;
            float T[256];
            unsigned char done[256];
            for (int p = 0; p < 256; ++p) {
                T[p] = 1.0f;
                const int px = tx * 16 + (p & 15), py = ty * 16 + (p >> 4);
                done[p] = !(px < img_w && py < img_h);
            }
            for (int k = start; k < end; ++k) {
                const int g = ids_sorted[k];
                unsigned char warp_hit[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                for (int p = 0; p < 256; ++p) {
                    if (done[p]) continue;
                    ++visited;
                    const float dx = xys[2 * g] - (float)(tx * 16 + (p & 15));
                    const float dy = xys[2 * g + 1] - (float)(ty * 16 + (p >> 4));
                    const float sigma = 0.5f * (conics[3 * g] * dx * dx + conics[3 * g + 2] * dy * dy) +
                                        conics[3 * g + 1] * dx * dy;
                    if (sigma < 0.0f) continue;
                    const float alpha = fminf(0.999f, opac[g] * expf(-sigma));
                    if (alpha < 1.0f / 255.0f) continue;
                    const float nT = T[p] * (1.0f - alpha);
                    if (nT <= 1e-4f) { done[p] = 1; continue; }
                    T[p] = nT;
                    ++hits;
                    warp_hit[p >> 5] = 1;
                }
                for (int w = 0; w < 8; ++w) warp_hits += warp_hit[w];
            }
        }
    stats[0] = visited; stats[1] = hits; stats[2] = warp_hits; stats[3] = entries;
}
