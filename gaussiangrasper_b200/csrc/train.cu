// "Next rows" of the hot-path scope table: the two per-Gaussian streaming steps that sit
// directly behind the rasterizer's backward in a training iteration.
//   gg_adam_step     : one fused Adam update over the flat (all-reduced) gradient buffer, replacing the
//                      reference's nine per-group torch.optim.Adam instances
//                      (nerfstudio/configs/method_configs.py:618-664, engine/optimizers.py:138-171)
//   gg_densify_stats : the densification statistics of GaussianSplattingModel.after_train
//                      (nerfstudio/models/gaussian_splatting.py:373-393)
// Both are HBM-bound elementwise kernels: 28 B and 24 B per element / Gaussian.
#include "gg_common.cuh"
#include "gg_b200.h"

namespace gg {

constexpr int kMaxSegments = 8;

struct AdamArgs {
    int n_segments;
    float* param[kMaxSegments];
    long long offset[kMaxSegments + 1];  // offsets into the flat buffers; offset[n_segments] = total
    float step_size[kMaxSegments];       // lr / (1 - beta1^t)
    float beta1, beta2, eps, inv_sqrt_bc2;
};

// torch.optim.Adam (no amsgrad, no weight decay): m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
// p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
__global__ void __launch_bounds__(256)
adam_kernel(const AdamArgs a, const float* __restrict__ grad, float* __restrict__ exp_avg,
            float* __restrict__ exp_avg_sq) {
    const long long total = a.offset[a.n_segments];
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        int seg = 0;
#pragma unroll
        for (int s = 1; s < kMaxSegments; ++s)
            if (s < a.n_segments && i >= a.offset[s]) seg = s;
        const float g = grad[i];
        const float m = a.beta1 * exp_avg[i] + (1.0f - a.beta1) * g;
        const float v = a.beta2 * exp_avg_sq[i] + (1.0f - a.beta2) * g * g;
        exp_avg[i] = m;
        exp_avg_sq[i] = v;
        float* p = a.param[seg] + (i - a.offset[seg]);
        *p = *p - a.step_size[seg] * (m / (sqrtf(v) * a.inv_sqrt_bc2 + a.eps));
    }
}

__global__ void __launch_bounds__(256)
densify_stats_kernel(long long n, int n_views, const float* __restrict__ v_geo, const int32_t* __restrict__ radii,
                     float inv_max_dim, int first_call, float* __restrict__ xys_grad_norm,
                     float* __restrict__ vis_counts, float* __restrict__ max_2dsize) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    float gn = first_call ? 0.0f : xys_grad_norm[g];
    float vc = first_call ? 1.0f : vis_counts[g];
    float ms = max_2dsize[g];
    for (int v = 0; v < n_views; ++v) {
        const long long i = (long long)v * n + g;
        const int r = radii[i];
        const float2 d = *reinterpret_cast<const float2*>(v_geo + 8 * i);
        const float nrm = sqrtf(d.x * d.x + d.y * d.y);
        if (first_call && v == 0) {
            gn = nrm;  // the reference initialises with the norms of every Gaussian, visible or not
        } else if (r > 0) {
            gn += nrm;
            vc += 1.0f;
        }
        if (r > 0) ms = fmaxf(ms, (float)r * inv_max_dim);
    }
    xys_grad_norm[g] = gn;
    vis_counts[g] = vc;
    max_2dsize[g] = ms;
}

}  // namespace gg

using namespace gg;

extern "C" int gg_adam_step(int n_segments, float* const* params, const long long* offsets, const long long* counts,
                            const float* lrs, const float* grad_flat, float* exp_avg_flat, float* exp_avg_sq_flat,
                            float beta1, float beta2, float eps, int step, void* stream) {
    GG_REQUIRE(n_segments >= 1 && n_segments <= kMaxSegments, "gg_adam_step: 1..8 segments");
    GG_REQUIRE(params && offsets && counts && lrs && grad_flat && exp_avg_flat && exp_avg_sq_flat,
               "gg_adam_step: null pointer");
    GG_REQUIRE(step >= 1 && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f, "gg_adam_step: bad hyper-parameters");
    AdamArgs a;
    a.n_segments = n_segments;
    long long end = 0;
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    for (int s = 0; s < n_segments; ++s) {
        GG_REQUIRE(params[s] && counts[s] >= 0 && offsets[s] >= end, "gg_adam_step: segments must be ordered, disjoint");
        a.param[s] = params[s];
        a.offset[s] = offsets[s];
        a.step_size[s] = (float)((double)lrs[s] / bc1);
        end = offsets[s] + counts[s];
        if (s + 1 < n_segments) GG_REQUIRE(offsets[s + 1] == end, "gg_adam_step: segments must tile the flat buffer");
    }
    a.offset[n_segments] = end;
    for (int s = n_segments + 1; s <= kMaxSegments; ++s) a.offset[s] = end;
    a.beta1 = beta1; a.beta2 = beta2; a.eps = eps;
    a.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    if (end == 0) return GG_OK;
    int blocks = div_up(end, 256 * 8);
    if (blocks > 148 * 16) blocks = 148 * 16;
    adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a, grad_flat, exp_avg_flat, exp_avg_sq_flat);
    count_launch();
    return check_launch("adam_kernel");
}

extern "C" int gg_densify_stats(long long n, int n_views, const float* v_geo, const int32_t* radii, int img_h,
                                int img_w, int first_call, float* xys_grad_norm, float* vis_counts,
                                float* max_2dsize, void* stream) {
    GG_REQUIRE(n >= 1 && n_views >= 1 && img_h > 0 && img_w > 0, "gg_densify_stats: bad sizes");
    GG_REQUIRE(v_geo && radii && xys_grad_norm && vis_counts && max_2dsize, "gg_densify_stats: null pointer");
    GG_REQUIRE(((uintptr_t)v_geo & 7) == 0, "gg_densify_stats: v_geo misaligned");
    const float inv = 1.0f / (float)(img_h > img_w ? img_h : img_w);
    densify_stats_kernel<<<div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(n, n_views, v_geo, radii, inv, first_call,
                                                                           xys_grad_norm, vis_counts, max_2dsize);
    count_launch();
    return check_launch("densify_stats_kernel");
}
