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
    long long count[kMaxSegments];       // elements of each segment (segments may be padded apart)
    float step_size[kMaxSegments];       // lr / (1 - beta1^t), t = that segment's own update count
    float inv_sqrt_bc2[kMaxSegments];    // 1 / sqrt(1 - beta2^t)
    int mode[kMaxSegments];              // GG_ADAM_*
    float beta1, beta2, eps;
};

// torch.optim.Adam (no amsgrad, no weight decay): m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
// p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// Gradient accumulation (trainer.py:466-481: a group's gradients are zeroed at step % k == 0 and its optimizer
// steps at step % k == k-1 on the SUM of the k gradients): the modes keep that sum in `accum`.
__global__ void __launch_bounds__(256)
adam_kernel(const AdamArgs a, const float* __restrict__ grad, float* __restrict__ accum, float* __restrict__ exp_avg,
            float* __restrict__ exp_avg_sq) {
    const long long total = a.offset[a.n_segments];
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        int seg = 0;
#pragma unroll
        for (int s = 1; s < kMaxSegments; ++s)
            if (s < a.n_segments && i >= a.offset[s]) seg = s;
        if (i - a.offset[seg] >= a.count[seg]) continue;  // alignment padding between two segments
        const int mode = a.mode[seg];
        if (mode == GG_ADAM_SKIP) continue;
        float g = grad[i];
        if (mode == GG_ADAM_ACC_FIRST) { accum[i] = g; continue; }
        if (mode == GG_ADAM_ACC) { accum[i] += g; continue; }
        if (mode == GG_ADAM_ACC_STEP) g += accum[i];
        const float m = a.beta1 * exp_avg[i] + (1.0f - a.beta1) * g;
        const float v = a.beta2 * exp_avg_sq[i] + (1.0f - a.beta2) * g * g;
        exp_avg[i] = m;
        exp_avg_sq[i] = v;
        float* p = a.param[seg] + (i - a.offset[seg]);
        *p = *p - a.step_size[seg] * (m / (sqrtf(v) * a.inv_sqrt_bc2[seg] + a.eps));
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
                            const float* lrs, const int* steps, const int* modes, const float* grad_flat,
                            float* accum_flat, float* exp_avg_flat, float* exp_avg_sq_flat, float beta1, float beta2,
                            float eps, void* stream) {
    GG_REQUIRE(n_segments >= 1 && n_segments <= kMaxSegments, "gg_adam_step: 1..8 segments");
    GG_REQUIRE(params && offsets && counts && lrs && steps && grad_flat && exp_avg_flat && exp_avg_sq_flat,
               "gg_adam_step: null pointer");
    GG_REQUIRE(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f, "gg_adam_step: bad hyper-parameters");
    AdamArgs a;
    a.n_segments = n_segments;
    long long end = 0;
    for (int s = 0; s < n_segments; ++s) {
        GG_REQUIRE(params[s] && counts[s] >= 0 && offsets[s] >= end, "gg_adam_step: segments must be ordered, disjoint");
        const int mode = modes ? modes[s] : GG_ADAM_STEP;
        GG_REQUIRE(mode >= GG_ADAM_STEP && mode <= GG_ADAM_SKIP, "gg_adam_step: unknown segment mode");
        GG_REQUIRE((mode != GG_ADAM_ACC_FIRST && mode != GG_ADAM_ACC && mode != GG_ADAM_ACC_STEP) || accum_flat,
                   "gg_adam_step: an accumulating segment needs accum_flat");
        const bool stepping = mode == GG_ADAM_STEP || mode == GG_ADAM_ACC_STEP;
        GG_REQUIRE(!stepping || steps[s] >= 1, "gg_adam_step: the update count of a stepping segment starts at 1");
        a.param[s] = params[s];
        a.offset[s] = offsets[s];
        a.count[s] = counts[s];
        a.mode[s] = mode;
        const int t = steps[s] >= 1 ? steps[s] : 1;
        const double bc1 = 1.0 - pow((double)beta1, (double)t);
        const double bc2 = 1.0 - pow((double)beta2, (double)t);
        a.step_size[s] = (float)((double)lrs[s] / bc1);
        a.inv_sqrt_bc2[s] = (float)(1.0 / sqrt(bc2));
        end = offsets[s] + counts[s];
    }
    for (int s = n_segments; s < kMaxSegments; ++s) { a.count[s] = 0; a.mode[s] = GG_ADAM_SKIP; a.step_size[s] = 0.f; a.inv_sqrt_bc2[s] = 1.f; a.param[s] = nullptr; }
    a.offset[n_segments] = end;
    for (int s = n_segments + 1; s <= kMaxSegments; ++s) a.offset[s] = end;
    a.beta1 = beta1; a.beta2 = beta2; a.eps = eps;
    if (end == 0) return GG_OK;
    int blocks = div_up(end, 256 * 8);
    if (blocks > 148 * 16) blocks = 148 * 16;
    adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a, grad_flat, accum_flat, exp_avg_flat, exp_avg_sq_flat);
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

// ---------------------------------------------------------------------------------------------
// Refinement on the device: densify (split / duplicate) + cull, with the Adam-state surgery, as
// decide -> scan -> map -> gather (GaussianSplattingModel.refinement_after,
// split_gaussians, dup_gaussians, cull_gaussians, dup_in_optim, remove_from_optim:
// nerfstudio/models/gaussian_splatting.py:396-546, 333-371).  Output order is the reference's:
// [kept originals | kept split children, sample-major | kept duplicates].
// ---------------------------------------------------------------------------------------------
namespace gg {

constexpr float kSplitShrink = 1.6f;  // size_fac, gaussian_splatting.py:513
enum : uint8_t { kFlagSplit = 1, kFlagDup = 2, kFlagKeep = 4, kFlagKeepSplit = 8, kFlagKeepDup = 16 };

__device__ __forceinline__ float shrunk_log_scale(float ls) { return logf(expf(ls) / kSplitShrink); }

__global__ void __launch_bounds__(256)
refine_decide_kernel(int n, const float* __restrict__ xys_grad_norm, const float* __restrict__ vis_counts,
                     const float* __restrict__ max_2dsize, const float* __restrict__ log_scales,
                     const float* __restrict__ opacity_logit, const gg_refine_config c, uint8_t* __restrict__ flags,
                     int32_t* __restrict__ counts, long long cs /* stride between the four count arrays */) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float ls[3] = {log_scales[3 * i], log_scales[3 * i + 1], log_scales[3 * i + 2]};
    const float m2d = max_2dsize ? max_2dsize[i] : 0.0f;
    bool split = false, dup = false;
    if (c.do_densify) {
        // :404-418  avg_grad_norm = (xys_grad_norm / vis_counts) * 0.5 * max(W, H)
        const float avg = (xys_grad_norm[i] / vis_counts[i]) * 0.5f * c.max_dim;
        const bool high = avg > c.densify_grad_thresh;
        const float smax = fmaxf(fmaxf(expf(ls[0]), expf(ls[1])), expf(ls[2]));
        split = smax > c.densify_size_thresh;
        if (c.split_by_screen) split = split || (m2d > c.split_screen_size);
        split = split && high;
        if (split) {  // :513-514 the parent is shrunk in place before the duplicate test reads the scales (:419)
#pragma unroll
            for (int k = 0; k < 3; ++k) ls[k] = shrunk_log_scale(ls[k]);
        }
        const float smax2 = fmaxf(fmaxf(expf(ls[0]), expf(ls[1])), expf(ls[2]));
        dup = (smax2 <= c.densify_size_thresh) && high;
    }
    bool keep = true, keep_child = true;
    if (c.do_cull) {  // :466-483 over the concatenated set; children carry the parent's opacity / current scales
        const float op = 1.0f / (1.0f + expf(-opacity_logit[i]));
        bool cull = op < c.cull_alpha_thresh;
        bool cull_child = cull;
        if (c.cull_by_scale) {
            const float smax = fmaxf(fmaxf(expf(ls[0]), expf(ls[1])), expf(ls[2]));
            const bool big = smax > c.cull_scale_thresh;
            cull = cull || big;
            cull_child = cull_child || big;
            if (c.cull_by_screen) cull = cull || (m2d > c.cull_screen_size);  // children start with max_2Dsize 0
        }
        keep = !cull;
        keep_child = !cull_child;
    }
    const bool ks = split && keep_child, kd = dup && keep_child;
    flags[i] = (uint8_t)((split ? kFlagSplit : 0) | (dup ? kFlagDup : 0) | (keep ? kFlagKeep : 0) |
                         (ks ? kFlagKeepSplit : 0) | (kd ? kFlagKeepDup : 0));
    counts[i] = keep ? 1 : 0;
    counts[cs + i] = ks ? 1 : 0;
    counts[2 * cs + i] = kd ? 1 : 0;
    counts[3 * cs + i] = split ? 1 : 0;
}

// source row / role of every output row: role 0 original, 1 split child, 2 duplicate; aux = row of the
// normal sample the child's position is drawn with
__global__ void __launch_bounds__(256)
refine_map_kernel(int n, int samps, const uint8_t* __restrict__ flags, const int32_t* __restrict__ scans, long long cs,
                  int n_keep, int n_split_kept, int n_split_all, int32_t* __restrict__ src_row,
                  uint8_t* __restrict__ role, int32_t* __restrict__ aux) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t f = flags[i];
    if (f & kFlagKeep) {
        const int j = scans[i] - 1;
        src_row[j] = i; role[j] = 0; aux[j] = 0;
    }
    if (f & kFlagKeepSplit) {
        const int r = scans[cs + i] - 1, ra = scans[3 * cs + i] - 1;
        for (int s = 0; s < samps; ++s) {
            const int j = n_keep + s * n_split_kept + r;
            src_row[j] = i; role[j] = 1; aux[j] = s * n_split_all + ra;
        }
    }
    if (f & kFlagKeepDup) {
        const int j = n_keep + samps * n_split_kept + scans[2 * cs + i] - 1;
        src_row[j] = i; role[j] = 2; aux[j] = 0;
    }
}

// Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11): a counter-based generator,
// so the split samples are a pure function of (seed, refinement step, parent row, sample index).  Every rank of a
// view-sharded run derives identical children without a broadcast (the reference draws torch.randn per process,
// gaussian_splatting.py:491, which is what breaks its multi-GPU training).
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                      uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// component e (0..2) of the standard-normal sample of (parent row, sample index): Box-Muller on the four words
__device__ __forceinline__ float philox_normal(unsigned long long seed, uint32_t step, uint32_t parent, uint32_t sample, int e) {
    uint32_t w[4];
    philox4x32_10(parent, sample, step, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), w);
    const float inv = 2.3283064365386963e-10f;  // 2^-32
    const float u1 = ((float)(e < 2 ? w[0] : w[2]) + 0.5f) * inv, u2 = ((float)(e < 2 ? w[1] : w[3]) + 0.5f) * inv;
    const float r = sqrtf(-2.0f * logf(fminf(fmaxf(u1, 1e-12f), 1.0f)));
    const float th = 6.283185307179586f * u2;
    return e == 1 ? r * sinf(th) : r * cosf(th);
}

__global__ void __launch_bounds__(256)
philox_normals_kernel(long long count, const int32_t* __restrict__ parents, int n_samples, unsigned long long seed,
                      uint32_t step, float* __restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n_samples * count * 3) return;
    const int e = (int)(idx % 3);
    const long long row = idx / 3;
    const long long r = row % count;
    const uint32_t s = (uint32_t)(row / count);
    const uint32_t parent = parents ? (uint32_t)parents[r] : (uint32_t)r;
    out[idx] = philox_normal(seed, step, parent, s, e);
}

struct RefineArrays {
    int n_arrays;
    const float* src[GG_REFINE_MAX_ARRAYS];
    float* dst[GG_REFINE_MAX_ARRAYS];
    int row[GG_REFINE_MAX_ARRAYS];
    int kind[GG_REFINE_MAX_ARRAYS];
};

__global__ void __launch_bounds__(256)
refine_gather_kernel(long long n_out, const RefineArrays t, const int32_t* __restrict__ src_row,
                     const uint8_t* __restrict__ role, const int32_t* __restrict__ aux, const uint8_t* __restrict__ flags,
                     const float* __restrict__ means, const float* __restrict__ log_scales,
                     const float* __restrict__ quats, const float* __restrict__ samples, int n_split_all,
                     unsigned long long seed, uint32_t step) {
    const int a = blockIdx.y;
    const int row = t.row[a], kind = t.kind[a];
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_out * row) return;
    const long long j = idx / row;
    const int e = (int)(idx - j * row);
    const int i = src_row[j];
    const int r = role[j];
    float v = t.src[a][(long long)i * row + e];
    if (kind == GG_REFINE_MOMENT) {
        if (r != 0) v = 0.0f;  // dup_in_optim appends zeros (:355-366); survivors keep their moments (:345-346)
    } else if (kind == GG_REFINE_LOG_SCALES) {
        if (flags[i] & kFlagSplit) v = shrunk_log_scale(v);  // parent, its children and a duplicate of it (:512-514)
    } else if (kind == GG_REFINE_MEANS && r == 1) {
        // :499-507  mean + R(q/|q|) (exp(scale_before_shrink) * z)
        const float4 q = *reinterpret_cast<const float4*>(quats + 4 * (long long)i);
        const float inv = 1.0f / sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
        const float w = q.x * inv, x = q.y * inv, y = q.z * inv, z = q.w * inv;
        float z0, z1, z2;
        if (samples) {  // caller-provided normals, row s * n_split_all + (rank of the parent among the split parents)
            const float* smp = samples + 3 * (long long)aux[j];
            z0 = smp[0]; z1 = smp[1]; z2 = smp[2];
        } else {        // counter-based: keyed on the parent's row and the sample index, not on any rank-local state
            const uint32_t smp_i = (uint32_t)(aux[j] / n_split_all);
            z0 = philox_normal(seed, step, (uint32_t)i, smp_i, 0);
            z1 = philox_normal(seed, step, (uint32_t)i, smp_i, 1);
            z2 = philox_normal(seed, step, (uint32_t)i, smp_i, 2);
        }
        const float s0 = expf(log_scales[3 * (long long)i]) * z0;
        const float s1 = expf(log_scales[3 * (long long)i + 1]) * z1;
        const float s2 = expf(log_scales[3 * (long long)i + 2]) * z2;
        float r0, r1, r2;  // row e of the rotation matrix (same convention as quat_to_rotmat)
        if (e == 0) { r0 = 1.f - 2.f * (y * y + z * z); r1 = 2.f * (x * y - w * z); r2 = 2.f * (x * z + w * y); }
        else if (e == 1) { r0 = 2.f * (x * y + w * z); r1 = 1.f - 2.f * (x * x + z * z); r2 = 2.f * (y * z - w * x); }
        else { r0 = 2.f * (x * z - w * y); r1 = 2.f * (y * z + w * x); r2 = 1.f - 2.f * (x * x + y * y); }
        v = (r0 * s0 + r1 * s1 + r2 * s2) + v;
    }
    t.dst[a][j * row + e] = v;
}

}  // namespace gg

// the four count / scan arrays are `refine_stride(n)` entries apart so that each starts 16-byte aligned
static size_t refine_stride(int n) { return ((size_t)(n > 0 ? n : 1) + 3) & ~(size_t)3; }

extern "C" size_t gg_refine_workspace_bytes(int n) {
    const size_t t = refine_stride(n);
    return 256 + ((t + 255) & ~(size_t)255) + 2 * 4 * t * sizeof(int32_t) + 4 * gg_cumsum_workspace_bytes(n) + 1024;
}

extern "C" int gg_refine_plan(int n, const float* xys_grad_norm, const float* vis_counts, const float* max_2dsize,
                              const float* log_scales, const float* opacity_logit, const gg_refine_config* cfg,
                              void* workspace, size_t workspace_bytes, int32_t* totals_host, void* stream) {
    GG_REQUIRE(n >= 1, "gg_refine_plan: need n >= 1");
    GG_REQUIRE(cfg && log_scales && opacity_logit && workspace && totals_host, "gg_refine_plan: null pointer");
    GG_REQUIRE(!cfg->do_densify || (xys_grad_norm && vis_counts), "gg_refine_plan: densify needs the gradient statistics");
    GG_REQUIRE((!cfg->split_by_screen && !cfg->cull_by_screen) || max_2dsize, "gg_refine_plan: screen-size rules need max_2dsize");
    GG_REQUIRE(workspace_bytes >= gg_refine_workspace_bytes(n), "gg_refine_plan: workspace too small");
    GG_REQUIRE(((uintptr_t)workspace & 255) == 0, "gg_refine_plan: workspace must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
    int32_t* totals_dev = reinterpret_cast<int32_t*>(w);
    uint8_t* flags = w + 256;
    const size_t t = refine_stride(n);
    int32_t* counts = reinterpret_cast<int32_t*>(w + 256 + ((t + 255) & ~(size_t)255));
    int32_t* scans = counts + 4 * t;
    unsigned char* scan_ws = reinterpret_cast<unsigned char*>(scans + 4 * t);
    scan_ws = reinterpret_cast<unsigned char*>(((uintptr_t)scan_ws + 255) & ~(uintptr_t)255);
    refine_decide_kernel<<<div_up(n, 256), 256, 0, st>>>(n, xys_grad_norm, vis_counts, max_2dsize, log_scales,
                                                        opacity_logit, *cfg, flags, counts, (long long)t);
    count_launch();
    int rc = check_launch("refine_decide_kernel");
    if (rc) return rc;
    const size_t sws = gg_cumsum_workspace_bytes(n);
    for (int k = 0; k < 4; ++k) {
        if ((rc = gg_cumsum(n, counts + k * t, scans + k * t, totals_dev + k, scan_ws, sws, stream))) return rc;
    }
    GG_CUDA(cudaMemcpyAsync(totals_host, totals_dev, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    return GG_OK;
}

extern "C" int gg_refine_apply(int n, int n_split_samples, const int32_t* totals, const void* plan_workspace,
                               int n_arrays, const float* const* src, float* const* dst, const int* row_floats,
                               const int* kinds, const float* means, const float* log_scales, const float* quats,
                               const float* samples, unsigned long long seed, unsigned int step, void* scratch,
                               size_t scratch_bytes, void* stream) {
    GG_REQUIRE(n >= 1 && n_split_samples >= 1 && totals && plan_workspace, "gg_refine_apply: bad arguments");
    GG_REQUIRE(n_arrays >= 1 && n_arrays <= GG_REFINE_MAX_ARRAYS && src && dst && row_floats && kinds,
               "gg_refine_apply: 1..GG_REFINE_MAX_ARRAYS arrays");
    GG_REQUIRE(means && log_scales && quats, "gg_refine_apply: null parameter pointer");
    const long long n_keep = totals[0], n_sk = totals[1], n_dk = totals[2], n_sa = totals[3];
    GG_REQUIRE(n_keep >= 0 && n_sk >= 0 && n_dk >= 0 && n_sa >= n_sk && n_keep <= n, "gg_refine_apply: inconsistent totals");
    const long long n_out = n_keep + (long long)n_split_samples * n_sk + n_dk;
    GG_REQUIRE(n_out < (1ll << 31), "gg_refine_apply: too many Gaussians");
    // samples == NULL: the children's offsets come from Philox4x32-10 keyed on (seed, step, parent row, sample)
    if (n_out == 0) return GG_OK;
    GG_REQUIRE(scratch && scratch_bytes >= (size_t)n_out * 9 + 512, "gg_refine_apply: scratch too small (9 B per output row + 512)");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned char* w = reinterpret_cast<const unsigned char*>(plan_workspace);
    const uint8_t* flags = w + 256;
    const size_t t = refine_stride(n);
    const int32_t* scans = reinterpret_cast<const int32_t*>(w + 256 + ((t + 255) & ~(size_t)255)) + 4 * t;
    unsigned char* s = reinterpret_cast<unsigned char*>(scratch);
    int32_t* src_row = reinterpret_cast<int32_t*>(s);
    int32_t* aux = src_row + n_out;
    uint8_t* role = reinterpret_cast<uint8_t*>(aux + n_out);
    refine_map_kernel<<<div_up(n, 256), 256, 0, st>>>(n, n_split_samples, flags, scans, (long long)t, (int)n_keep, (int)n_sk, (int)n_sa,
                                                     src_row, role, aux);
    count_launch();
    int rc = check_launch("refine_map_kernel");
    if (rc) return rc;
    RefineArrays tab;
    tab.n_arrays = n_arrays;
    int max_row = 1;
    for (int a = 0; a < n_arrays; ++a) {
        GG_REQUIRE(src[a] && dst[a] && row_floats[a] >= 1, "gg_refine_apply: bad array entry");
        GG_REQUIRE(kinds[a] >= GG_REFINE_COPY && kinds[a] <= GG_REFINE_LOG_SCALES, "gg_refine_apply: unknown array kind");
        GG_REQUIRE((kinds[a] != GG_REFINE_MEANS && kinds[a] != GG_REFINE_LOG_SCALES) || row_floats[a] == 3,
                   "gg_refine_apply: means / log_scales rows are 3 floats");
        tab.src[a] = src[a]; tab.dst[a] = dst[a]; tab.row[a] = row_floats[a]; tab.kind[a] = kinds[a];
        if (row_floats[a] > max_row) max_row = row_floats[a];
    }
    dim3 grid((unsigned)div_up(n_out * max_row, 256), (unsigned)n_arrays);
    refine_gather_kernel<<<grid, 256, 0, st>>>(n_out, tab, src_row, role, aux, flags, means, log_scales, quats, samples,
                                               (int)(n_sa > 0 ? n_sa : 1), seed, step);
    count_launch();
    return check_launch("refine_gather_kernel");
}

extern "C" int gg_philox_normals(long long count, const int32_t* parents, int n_samples, unsigned long long seed,
                                 unsigned int step, float* out, void* stream) {
    GG_REQUIRE(count >= 0 && n_samples >= 1 && count < (1ll << 31), "gg_philox_normals: bad sizes");
    if (count == 0) return GG_OK;
    GG_REQUIRE(out, "gg_philox_normals: null output");
    const long long total = (long long)n_samples * count * 3;
    philox_normals_kernel<<<(unsigned)div_up(total, 256), 256, 0, (cudaStream_t)stream>>>(count, parents, n_samples, seed,
                                                                                         step, out);
    count_launch();
    return check_launch("philox_normals_kernel");
}

extern "C" void gg_philox4x32_10_host(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]) {
    gg::philox4x32_10(counter[0], counter[1], counter[2], counter[3], key[0], key[1], out);
}

// ---------------------------------------------------------------------------------------------
// Per-pixel L1 / L2 loss with its gradient in one pass (the L1 term of
// GaussianSplattingModel.get_loss_dict, gaussian_splatting.py:861-866, and the masked variant :853-858 where
// masked pixels are zeroed in both images but still count in the mean).
//   loss = weight * mean(|p - t|)  or  weight * mean((p - t)^2)
//   grad = d loss / d p, written for every element (0 where masked)
// The mean runs over all n elements, or -- with valid_pixels, a device int32 holding the number of unmasked
// pixels -- over the valid ones only, which is what :882 computes.
// One read of each image, one write of the gradient; the loss is reduced per block and summed by the last
// block to finish, in block order (deterministic).
// ---------------------------------------------------------------------------------------------
namespace gg {

__device__ __forceinline__ float pixel_loss_term(float d, int l2, float scale, float& acc) {
    if (l2) { acc += d * d; return d * (2.0f * scale); }
    acc += fabsf(d);
    return d > 0.0f ? scale : (d < 0.0f ? -scale : 0.0f);
}

__global__ void __launch_bounds__(256)
pixel_loss_kernel(long long n, int channels, const float* __restrict__ pred, const float* __restrict__ target,
                  const uint8_t* __restrict__ mask, const int32_t* __restrict__ valid_pixels, int l2, float weight,
                  float* __restrict__ grad, float* __restrict__ partial, unsigned int* __restrict__ counter,
                  float* __restrict__ loss, int vec4) {
    __shared__ float warp_sum[8];
    __shared__ bool last;
    __shared__ float s_scale;
    // mean over every element, or over the elements of the valid pixels only (gaussian_splatting.py:882).
    // ONE thread per block does the double-precision division: fp64 throughput on this part is ~1/64 of fp32, and
    // half a million threads each dividing in double cost more than the whole memory traffic of the kernel.
    if (threadIdx.x == 0) {
        const double denom = valid_pixels ? (double)max(*valid_pixels, 1) * (double)channels : (double)n;
        s_scale = (float)((double)weight / denom);
    }
    __syncthreads();
    const float scale = s_scale;
    float acc = 0.0f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    if (vec4) {
        // 16-byte accesses (the host checked alignment, n % 4 == 0 and, with a mask, channels % 4 == 0 so that a
        // vector never straddles two pixels): three 16-byte streams per thread, two vectors in flight
        const long long n4 = n >> 2;
        const float4* p4 = reinterpret_cast<const float4*>(pred);
        const float4* t4 = reinterpret_cast<const float4*>(target);
        float4* g4 = reinterpret_cast<float4*>(grad);
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
            const bool on = mask == nullptr || mask[(4 * i) / channels] != 0;
            const float4 a = p4[i], b = t4[i];
            float4 g;
            g.x = pixel_loss_term(on ? a.x - b.x : 0.0f, l2, scale, acc);
            g.y = pixel_loss_term(on ? a.y - b.y : 0.0f, l2, scale, acc);
            g.z = pixel_loss_term(on ? a.z - b.z : 0.0f, l2, scale, acc);
            g.w = pixel_loss_term(on ? a.w - b.w : 0.0f, l2, scale, acc);
            g4[i] = g;
        }
    } else {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            const bool on = mask == nullptr || mask[i / channels] != 0;
            grad[i] = pixel_loss_term(on ? pred[i] - target[i] : 0.0f, l2, scale, acc);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.0f;
        for (int w = 0; w < 8; ++w) s += warp_sum[w];
        partial[blockIdx.x] = s;
        __threadfence();
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
        // the last block to finish adds the partials in a fixed order (deterministic): strided per thread, then a
        // shared-memory tree -- one thread walking 2048 partials alone was most of this kernel's time
        __shared__ double tree[256];
        __threadfence();
        double s = 0.0;
        for (unsigned b = threadIdx.x; b < gridDim.x; b += 256) s += (double)__ldcg(partial + b);
        tree[threadIdx.x] = s;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if ((int)threadIdx.x < o) tree[threadIdx.x] += tree[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            *loss = (float)(tree[0] * (double)scale);
            *counter = 0u;  // ready for the next call
        }
    }
}

}  // namespace gg

extern "C" size_t gg_pixel_loss_workspace_bytes(void) { return sizeof(float) * 2048 + 16; }

extern "C" int gg_pixel_loss(long long n, int channels, const float* pred, const float* target, const uint8_t* mask,
                             const int32_t* valid_pixels, int kind, float weight, float* grad, float* loss,
                             void* workspace, size_t workspace_bytes, void* stream) {
    GG_REQUIRE(n >= 1 && channels >= 1 && n % channels == 0, "gg_pixel_loss: n must be a positive multiple of channels");
    GG_REQUIRE(kind == 1 || kind == 2, "gg_pixel_loss: kind is 1 (L1) or 2 (L2)");
    GG_REQUIRE(pred && target && grad && loss && workspace, "gg_pixel_loss: null pointer");
    GG_REQUIRE(workspace_bytes >= gg_pixel_loss_workspace_bytes(), "gg_pixel_loss: workspace too small");
    const int vec4 = (n % 4 == 0) && (mask == nullptr || channels % 4 == 0) &&
                     ((((uintptr_t)pred) | ((uintptr_t)target) | ((uintptr_t)grad)) & 15) == 0;
    int blocks = div_up(vec4 ? n / 4 : n, 256 * 4);
    if (blocks > 2048) blocks = 2048;
    if (blocks < 1) blocks = 1;
    float* partial = reinterpret_cast<float*>(workspace);
    unsigned int* counter = reinterpret_cast<unsigned int*>(partial + 2048);  // zero-initialised by the caller once
    pixel_loss_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(n, channels, pred, target, mask, valid_pixels,
                                                               kind == 2 ? 1 : 0, weight, grad, partial, counter, loss, vec4);
    count_launch();
    return check_launch("pixel_loss_kernel");
}

// ---------------------------------------------------------------------------------------------
// SSIM loss with its gradient: weight * (1 - SSIM(pred, target)) as the reference uses it
// (pytorch_msssim.SSIM(data_range=1, size_average=True, channel=3): 11x11 Gaussian window, sigma 1.5, no
// padding, K = (0.01, 0.03); gaussian_splatting.py:284, :885).  Images are channel-last [n_img, H, W, stride];
// the first `channels` channels are compared.
//   ssim_stats_kernel : per output position, the five windowed moments -> SSIM value (summed per block) and the
//                       partial derivatives w.r.t. E[x], E[x^2], E[xy], written as three maps
//   ssim_grad_kernel  : d loss / d x[q] = -scale * sum_p w(q - p) (dmu[p] + 2 x[q] ds11[p] + y[q] ds12[p])
// Both stage a 26x26 patch per 16x16 tile in shared memory and apply the window directly in 2-D.
// ---------------------------------------------------------------------------------------------
namespace gg {

constexpr int kSsimWin = 11, kSsimTile = 16, kSsimPatch = kSsimTile + kSsimWin - 1;

struct SsimArgs {
    int n_img, H, W, channels, oh, ow;
    int ps, ts, gs;  // floats per pixel of pred, target, grad
    float g[kSsimWin];
    float c1, c2;
};

__global__ void __launch_bounds__(kSsimTile * kSsimTile)
ssim_stats_kernel(const SsimArgs a, const float* __restrict__ pred, const float* __restrict__ target,
                  float* __restrict__ maps, float* __restrict__ partial) {
    __shared__ float xs[kSsimPatch][kSsimPatch + 1], ys[kSsimPatch][kSsimPatch + 1];
    __shared__ float warp_sum[8];
    const int img = blockIdx.z / a.channels, c = blockIdx.z - img * a.channels;
    const int ox0 = blockIdx.x * kSsimTile, oy0 = blockIdx.y * kSsimTile;
    const int tid = threadIdx.y * kSsimTile + threadIdx.x;
    for (int k = tid; k < kSsimPatch * kSsimPatch; k += kSsimTile * kSsimTile) {
        const int py = k / kSsimPatch, px = k - py * kSsimPatch;
        const int y = oy0 + py, x = ox0 + px;
        float vx = 0.0f, vy = 0.0f;
        if (y < a.H && x < a.W) {
            const long long pix = ((long long)img * a.H + y) * a.W + x;
            vx = pred[pix * a.ps + c];
            vy = target[pix * a.ts + c];
        }
        xs[py][px] = vx;
        ys[py][px] = vy;
    }
    __syncthreads();
    const int ox = ox0 + threadIdx.x, oy = oy0 + threadIdx.y;
    float val = 0.0f;
    if (ox < a.ow && oy < a.oh) {
        float mu1 = 0.f, mu2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
        for (int dy = 0; dy < kSsimWin; ++dy) {
            float r1 = 0.f, r2 = 0.f, r11 = 0.f, r22 = 0.f, r12 = 0.f;  // the window is applied row by row
#pragma unroll
            for (int dx = 0; dx < kSsimWin; ++dx) {
                const float x = xs[threadIdx.y + dy][threadIdx.x + dx], y = ys[threadIdx.y + dy][threadIdx.x + dx];
                const float w = a.g[dx];
                r1 = fmaf(w, x, r1); r2 = fmaf(w, y, r2);
                r11 = fmaf(w, x * x, r11); r22 = fmaf(w, y * y, r22); r12 = fmaf(w, x * y, r12);
            }
            const float w = a.g[dy];
            mu1 = fmaf(w, r1, mu1); mu2 = fmaf(w, r2, mu2);
            s11 = fmaf(w, r11, s11); s22 = fmaf(w, r22, s22); s12 = fmaf(w, r12, s12);
        }
        const float v1 = s11 - mu1 * mu1, v2 = s22 - mu2 * mu2, cov = s12 - mu1 * mu2;
        const float A1 = 2.0f * mu1 * mu2 + a.c1, A2 = 2.0f * cov + a.c2;
        const float B1 = mu1 * mu1 + mu2 * mu2 + a.c1, B2 = v1 + v2 + a.c2;
        const float inv = 1.0f / (B1 * B2);
        val = A1 * A2 * inv;
        const float dA1 = A2 * inv, dA2 = A1 * inv, dB1 = -val / B1, dB2 = -val / B2;
        // x enters through mu1 (A1, B1, and the covariance / variance terms) and through E[x^2], E[xy]
        const float dmu = 2.0f * mu2 * dA1 + 2.0f * mu1 * dB1 - 2.0f * mu2 * dA2 - 2.0f * mu1 * dB2;
        const long long plane = (long long)a.oh * a.ow;
        const long long o = ((long long)blockIdx.z * 3) * plane + (long long)oy * a.ow + ox;
        maps[o] = dmu;
        maps[o + plane] = dB2;          // d / d E[x^2]
        maps[o + 2 * plane] = 2.0f * dA2;  // d / d E[xy]
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
    if ((tid & 31) == 0) warp_sum[tid >> 5] = val;
    __syncthreads();
    if (tid == 0) {
        float s = 0.0f;
        for (int w = 0; w < 8; ++w) s += warp_sum[w];
        partial[((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(256)
ssim_sum_kernel(long long n_partial, const float* __restrict__ partial, double inv_count, float weight,
                float* __restrict__ loss, int accumulate) {
    __shared__ double sh[256];
    double s = 0.0;
    for (long long i = threadIdx.x; i < n_partial; i += 256) s += (double)partial[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float l = (float)((double)weight * (1.0 - sh[0] * inv_count));
        *loss = accumulate ? *loss + l : l;
    }
}

__global__ void __launch_bounds__(kSsimTile * kSsimTile)
ssim_grad_kernel(const SsimArgs a, const float* __restrict__ pred, const float* __restrict__ target,
                 const float* __restrict__ maps, float scale, float* __restrict__ grad, int accumulate) {
    __shared__ float m0[kSsimPatch][kSsimPatch + 1], m1[kSsimPatch][kSsimPatch + 1], m2[kSsimPatch][kSsimPatch + 1];
    const int img = blockIdx.z / a.channels, c = blockIdx.z - img * a.channels;
    const int qx0 = blockIdx.x * kSsimTile, qy0 = blockIdx.y * kSsimTile;
    const int tid = threadIdx.y * kSsimTile + threadIdx.x;
    const long long plane = (long long)a.oh * a.ow;
    const float* mp = maps + ((long long)blockIdx.z * 3) * plane;
    // output positions p in [q - 10, q] cover input pixel q: patch origin is (qy0 - 10, qx0 - 10)
    for (int k = tid; k < kSsimPatch * kSsimPatch; k += kSsimTile * kSsimTile) {
        const int py = k / kSsimPatch, px = k - py * kSsimPatch;
        const int oy = qy0 - (kSsimWin - 1) + py, ox = qx0 - (kSsimWin - 1) + px;
        float v0 = 0.f, v1 = 0.f, v2 = 0.f;
        if (oy >= 0 && ox >= 0 && oy < a.oh && ox < a.ow) {
            const long long o = (long long)oy * a.ow + ox;
            v0 = mp[o]; v1 = mp[o + plane]; v2 = mp[o + 2 * plane];
        }
        m0[py][px] = v0; m1[py][px] = v1; m2[py][px] = v2;
    }
    __syncthreads();
    const int qx = qx0 + threadIdx.x, qy = qy0 + threadIdx.y;
    if (qx >= a.W || qy >= a.H) return;
    float g0 = 0.f, g1 = 0.f, g2 = 0.f;
#pragma unroll
    for (int dy = 0; dy < kSsimWin; ++dy) {
        float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll
        for (int dx = 0; dx < kSsimWin; ++dx) {
            // position p = q - d sits at patch (ty + 10 - dy, tx + 10 - dx)
            const int py = threadIdx.y + (kSsimWin - 1) - dy, px = threadIdx.x + (kSsimWin - 1) - dx;
            const float w = a.g[dx];
            r0 = fmaf(w, m0[py][px], r0); r1 = fmaf(w, m1[py][px], r1); r2 = fmaf(w, m2[py][px], r2);
        }
        const float w = a.g[dy];
        g0 = fmaf(w, r0, g0); g1 = fmaf(w, r1, g1); g2 = fmaf(w, r2, g2);
    }
    const long long pix = ((long long)img * a.H + qy) * a.W + qx;
    const float x = pred[pix * a.ps + c], y = target[pix * a.ts + c];
    const float v = -scale * (g0 + 2.0f * x * g1 + y * g2);
    float* gp = grad + pix * a.gs + c;
    *gp = accumulate ? *gp + v : v;
}

}  // namespace gg

extern "C" size_t gg_ssim_workspace_bytes(int n_img, int img_h, int img_w, int channels) {
    if (n_img < 1 || channels < 1 || img_h < 11 || img_w < 11) return 0;
    const size_t oh = (size_t)img_h - 10, ow = (size_t)img_w - 10;
    const size_t blocks = (size_t)div_up((long long)ow, 16) * (size_t)div_up((long long)oh, 16) * (size_t)n_img * channels;
    return sizeof(float) * (3 * oh * ow * (size_t)n_img * channels + blocks) + 256;
}

extern "C" int gg_ssim_loss(int n_img, int img_h, int img_w, int channels, const float* pred, int pred_stride,
                            const float* target, int target_stride, float weight, float* grad, int grad_stride,
                            int accumulate, float* loss, void* workspace, size_t workspace_bytes, void* stream) {
    GG_REQUIRE(n_img >= 1 && channels >= 1 && img_h >= 11 && img_w >= 11, "gg_ssim_loss: images must be at least 11x11");
    GG_REQUIRE(pred && target && grad && loss && workspace, "gg_ssim_loss: null pointer");
    GG_REQUIRE(pred_stride >= channels && target_stride >= channels && grad_stride >= channels,
               "gg_ssim_loss: strides must cover the compared channels");
    GG_REQUIRE((long long)n_img * channels <= 65535, "gg_ssim_loss: too many image planes");
    GG_REQUIRE(workspace_bytes >= gg_ssim_workspace_bytes(n_img, img_h, img_w, channels), "gg_ssim_loss: workspace too small");
    SsimArgs a;
    a.n_img = n_img; a.H = img_h; a.W = img_w; a.channels = channels; a.oh = img_h - 10; a.ow = img_w - 10;
    a.ps = pred_stride; a.ts = target_stride; a.gs = grad_stride;
    // pytorch_msssim._fspecial_gauss_1d(11, 1.5) in fp32
    float sum = 0.0f;
    for (int i = 0; i < kSsimWin; ++i) {
        const float d = (float)(i - kSsimWin / 2);
        a.g[i] = expf(-(d * d) / (2.0f * 1.5f * 1.5f));
        sum += a.g[i];
    }
    for (int i = 0; i < kSsimWin; ++i) a.g[i] /= sum;
    a.c1 = 0.01f * 0.01f;
    a.c2 = 0.03f * 0.03f;
    cudaStream_t st = (cudaStream_t)stream;
    float* maps = reinterpret_cast<float*>(workspace);
    const long long plane = (long long)a.oh * a.ow;
    float* partial = maps + 3 * plane * n_img * channels;
    dim3 block(kSsimTile, kSsimTile);
    dim3 ogrid(div_up(a.ow, kSsimTile), div_up(a.oh, kSsimTile), n_img * channels);
    ssim_stats_kernel<<<ogrid, block, 0, st>>>(a, pred, target, maps, partial);
    const long long n_partial = (long long)ogrid.x * ogrid.y * ogrid.z;
    const double count = (double)plane * n_img * channels;
    ssim_sum_kernel<<<1, 256, 0, st>>>(n_partial, partial, 1.0 / count, weight, loss, accumulate);
    dim3 igrid(div_up(img_w, kSsimTile), div_up(img_h, kSsimTile), n_img * channels);
    ssim_grad_kernel<<<igrid, block, 0, st>>>(a, pred, target, maps, (float)((double)weight / count), grad, accumulate);
    count_launch(3);
    return check_launch("gg_ssim_loss");
}


// ---------------------------------------------------------------------------------------------
// k-nearest-neighbour scale initialisation (GaussianSplattingModel.populate_modules,
// gaussian_splatting.py:259-263, k_nearest_sklearn :315-331): scales = log(mean distance to the 3 nearest other
// points), replicated over the three axes.  Exact brute force: one query per thread, candidates staged through
// shared memory in tiles of 1024 points, the three smallest squared distances kept in registers.  N^2 / 2 pair
// evaluations would suffice with symmetry; the simple form is ~50 ms at 500 k points, once per scene.
// ---------------------------------------------------------------------------------------------
namespace gg {

constexpr int kKnnTile = 1024;

__global__ void __launch_bounds__(256)
knn3_kernel(int n, const float* __restrict__ means, float* __restrict__ dist3, float* __restrict__ log_scales) {
    __shared__ float4 pts[kKnnTile];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (i < n) { qx = means[3 * (long long)i]; qy = means[3 * (long long)i + 1]; qz = means[3 * (long long)i + 2]; }
    float b0 = 3.4e38f, b1 = 3.4e38f, b2 = 3.4e38f;   // ascending
    for (int base = 0; base < n; base += kKnnTile) {
        const int cnt = min(kKnnTile, n - base);
        __syncthreads();
        for (int k = threadIdx.x; k < cnt; k += blockDim.x) {
            const long long j = base + k;
            pts[k] = make_float4(means[3 * j], means[3 * j + 1], means[3 * j + 2], 0.f);
        }
        __syncthreads();
        if (i < n) {
#pragma unroll 8
            for (int k = 0; k < cnt; ++k) {
                const float4 p = pts[k];
                const float dx = p.x - qx, dy = p.y - qy, dz = p.z - qz;
                const float d = dx * dx + dy * dy + dz * dz;
                if (d < b2 && base + k != i) {
                    if (d < b1) {
                        b2 = b1;
                        if (d < b0) { b1 = b0; b0 = d; } else { b1 = d; }
                    } else {
                        b2 = d;
                    }
                }
            }
        }
    }
    if (i >= n) return;
    // fewer than four points: the missing neighbours repeat the farthest one found (0 for a single point)
    if (b0 > 3e38f) b0 = 0.f;
    if (b1 > 3e38f) b1 = b0;
    if (b2 > 3e38f) b2 = b1;
    const float d0 = sqrtf(b0), d1 = sqrtf(b1), d2 = sqrtf(b2);
    if (dist3) { dist3[3 * (long long)i] = d0; dist3[3 * (long long)i + 1] = d1; dist3[3 * (long long)i + 2] = d2; }
    if (log_scales) {
        const float l = logf((d0 + d1 + d2) / 3.0f);
        log_scales[3 * (long long)i] = l; log_scales[3 * (long long)i + 1] = l; log_scales[3 * (long long)i + 2] = l;
    }
}

}  // namespace gg

extern "C" int gg_knn3_scales(int n, const float* means, float* dist3, float* log_scales, void* stream) {
    GG_REQUIRE(n >= 1 && means && (dist3 || log_scales), "gg_knn3_scales: bad arguments");
    knn3_kernel<<<div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(n, means, dist3, log_scales);
    count_launch();
    return check_launch("knn3_kernel");
}
