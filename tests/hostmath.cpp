// Test-only host build of the per-Gaussian device math (gaussiangrasper_b200/csrc/gg_math.cuh),
// so that the projection forward/backward formulas the CUDA kernels execute can be checked on a
// box without a GPU.  Not part of the product; never loaded outside tests/.
#include "gg_math.cuh"

using namespace gg;

static Camera make_cam(const float* vm, const float* fm, float fx, float fy, float cx, float cy) {
    Camera c;
    for (int i = 0; i < 12; ++i) c.vm[i] = vm[i];
    for (int i = 0; i < 16; ++i) c.fm[i] = fm[i];
    c.fx = fx; c.fy = fy; c.cx = cx; c.cy = cy;
    return c;
}

extern "C" void hm_project_fwd(int n, const float* means, const float* scales, float glob, const float* quats,
                               const float* vm, const float* fm, float fx, float fy, float cx, float cy, int H, int W,
                               int tiles_x, int tiles_y, float clip, float* cov3d, float* xys, float* depths,
                               int* radii, float* conics, int* nth) {
    const Camera cam = make_cam(vm, fm, fx, fy, cx, cy);
    for (int i = 0; i < n; ++i) {
        const ProjOut o = project_one(means + 3 * i, scales + 3 * i, glob, quats + 4 * i, cam, H, W, tiles_x, tiles_y, clip);
        for (int k = 0; k < 6; ++k) cov3d[6 * i + k] = o.cov3d[k];
        xys[2 * i] = o.ux; xys[2 * i + 1] = o.uy;
        depths[i] = o.depth; radii[i] = o.radius; nth[i] = o.tiles;
        for (int k = 0; k < 3; ++k) conics[3 * i + k] = o.conic[k];
    }
}

extern "C" void hm_project_bwd(int n, const float* means, const float* scales, float glob, const float* quats,
                               const float* vm, const float* fm, float fx, float fy, float cx, float cy, int H, int W,
                               const int* radii, const float* conics, const float* v_xys, const float* v_depths,
                               const float* v_conics, float* v_means, float* v_scales, float* v_quats) {
    const Camera cam = make_cam(vm, fm, fx, fy, cx, cy);
    for (int i = 0; i < n; ++i) {
        for (int k = 0; k < 3; ++k) v_means[3 * i + k] = v_scales[3 * i + k] = 0.f;
        for (int k = 0; k < 4; ++k) v_quats[4 * i + k] = 0.f;
        if (radii[i] <= 0) continue;
        const ProjGrad g = project_bwd_one(means + 3 * i, scales + 3 * i, glob, quats + 4 * i, cam, H, W, conics + 3 * i,
                                           v_xys + 2 * i, v_depths[i], v_conics + 3 * i);
        for (int k = 0; k < 3; ++k) { v_means[3 * i + k] = g.v_mean[k]; v_scales[3 * i + k] = g.v_scale[k]; }
        for (int k = 0; k < 4; ++k) v_quats[4 * i + k] = g.v_quat[k];
    }
}

extern "C" void hm_sh_basis(int n, int deg, const float* dirs, float* Y /*[n,25]*/) {
    for (int i = 0; i < n; ++i) sh_basis(deg, dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2], Y + 25 * i);
}
