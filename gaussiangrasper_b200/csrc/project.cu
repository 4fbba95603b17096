// Projection (EWA 3D->2D) and spherical-harmonics kernels, forward and backward.
// Compiled with -fmad=false: the integer outputs (radii, num_tiles_hit) and the depth bits that
// feed the sort keys must match the CPU oracle bit for bit.
//
// Replaces gsplat 0.1.0 project_gaussians_forward/backward_kernel and compute_sh_forward/
// backward_kernel as driven by nerfstudio/models/gaussian_splatting.py:699-713 and :730.
//
// All per-Gaussian arrays arrive as the row-major AoS tensors the reference passes ([N,3], [N,4],
// [N,25,3]); every kernel moves them through shared memory so that global traffic is issued as
// fully coalesced 4-byte / 16-byte accesses over the block's contiguous span (HBM-bound stage).
#include "gg_common.cuh"
#include "gg_math.cuh"
#include "gg_b200.h"

namespace gg {

constexpr int kProjThreads = 256;

// Coalesced copy of the block's [256, K] row-major span into shared memory (and back).
template <int K>
__device__ __forceinline__ void span_load(const float* __restrict__ src, long long first_row, long long n_rows,
                                          float* __restrict__ sm) {
    const long long base = first_row * K;
    const long long limit = n_rows * K;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const int e = threadIdx.x + j * kProjThreads;
        const long long idx = base + e;
        sm[e] = idx < limit ? __ldg(src + idx) : 0.0f;
    }
}
template <int K>
__device__ __forceinline__ void span_store(float* __restrict__ dst, long long first_row, long long n_rows,
                                           const float* __restrict__ sm) {
    const long long base = first_row * K;
    const long long limit = n_rows * K;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const int e = threadIdx.x + j * kProjThreads;
        const long long idx = base + e;
        if (idx < limit) dst[idx] = sm[e];
    }
}

__device__ __forceinline__ void load_camera(Camera& cam, const float* __restrict__ viewmats,
                                            const float* __restrict__ fullmats, const float* __restrict__ intrins,
                                            float fx, float fy, float cx, float cy, int view) {
    // executed by the first 32 threads of the block
    const int t = threadIdx.x;
    if (t < 12) cam.vm[t] = viewmats[(size_t)view * 12 + t];
    if (t < 16) cam.fm[t] = fullmats[(size_t)view * 16 + t];
    if (t == 0) {
        if (intrins) {
            cam.fx = intrins[4 * view]; cam.fy = intrins[4 * view + 1];
            cam.cx = intrins[4 * view + 2]; cam.cy = intrins[4 * view + 3];
        } else {
            cam.fx = fx; cam.fy = fy; cam.cx = cx; cam.cy = cy;
        }
    }
}

// grid = (ceil(N/256), V).  Per-view outputs live at [view * N + i].
__global__ void __launch_bounds__(kProjThreads)
project_fwd_kernel(int n, const float* __restrict__ means, const float* __restrict__ scales, float glob,
                   const float* __restrict__ quats, const float* __restrict__ viewmats,
                   const float* __restrict__ fullmats, const float* __restrict__ intrins, float fx, float fy,
                   float cx, float cy, int img_h, int img_w, int tiles_x, int tiles_y, float clip,
                   float* __restrict__ cov3d, float* __restrict__ xys, float* __restrict__ depths,
                   int32_t* __restrict__ radii, float* __restrict__ conics, int32_t* __restrict__ num_tiles_hit) {
    __shared__ Camera cam;
    __shared__ float sm_a[kProjThreads * 3];
    __shared__ float sm_b[kProjThreads * 6];
    const int view = blockIdx.y;
    const long long first = (long long)blockIdx.x * kProjThreads;
    const long long i = first + threadIdx.x;
    if (threadIdx.x < 32) load_camera(cam, viewmats, fullmats, intrins, fx, fy, cx, cy, view);
    span_load<3>(means, first, n, sm_a);
    span_load<3>(scales, first, n, sm_b);
    __syncthreads();
    float p[3], s[3], q[4];
#pragma unroll
    for (int k = 0; k < 3; ++k) { p[k] = sm_a[threadIdx.x * 3 + k]; s[k] = sm_b[threadIdx.x * 3 + k]; }
    ProjOut o;
    if (i < n) {
        const float4 qq = __ldg(reinterpret_cast<const float4*>(quats) + i);
        q[0] = qq.x; q[1] = qq.y; q[2] = qq.z; q[3] = qq.w;
        o = project_one(p, s, glob, q, cam, img_h, img_w, tiles_x, tiles_y, clip);
    }
    __syncthreads();
    const long long vbase = (long long)view * n;
    if (i < n) {
        reinterpret_cast<float2*>(xys)[vbase + i] = make_float2(o.ux, o.uy);
        depths[vbase + i] = o.depth;
        radii[vbase + i] = o.radius;
        num_tiles_hit[vbase + i] = o.tiles;
#pragma unroll
        for (int k = 0; k < 3; ++k) sm_a[threadIdx.x * 3 + k] = o.conic[k];
#pragma unroll
        for (int k = 0; k < 6; ++k) sm_b[threadIdx.x * 6 + k] = o.cov3d[k];
    }
    __syncthreads();
    span_store<3>(conics + vbase * 3, first, n, sm_a);
    span_store<6>(cov3d + vbase * 6, first, n, sm_b);
}

// grid = (ceil(N/256)).  Sums the contribution of all V views (no atomics): per-view gradient
// inputs live at [view * N + i]; outputs are [N, 3|3|4], written (accumulate=0) or added to.
__global__ void __launch_bounds__(kProjThreads)
project_bwd_kernel(int n, int n_views, const float* __restrict__ means, const float* __restrict__ scales, float glob,
                   const float* __restrict__ quats, const float* __restrict__ viewmats,
                   const float* __restrict__ fullmats, const float* __restrict__ intrins, float fx, float fy,
                   float cx, float cy, int img_h, int img_w, const int32_t* __restrict__ radii,
                   const float* __restrict__ conics, const float* __restrict__ v_xys,
                   const float* __restrict__ v_depths, const float* __restrict__ v_conics, int accumulate,
                   float* __restrict__ v_means, float* __restrict__ v_scales, float* __restrict__ v_quats) {
    __shared__ Camera cam;
    __shared__ float sm_a[kProjThreads * 3];
    __shared__ float sm_b[kProjThreads * 3];
    const long long first = (long long)blockIdx.x * kProjThreads;
    const long long i = first + threadIdx.x;
    span_load<3>(means, first, n, sm_a);
    span_load<3>(scales, first, n, sm_b);
    __syncthreads();
    float p[3], s[3], q[4] = {1.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 3; ++k) { p[k] = sm_a[threadIdx.x * 3 + k]; s[k] = sm_b[threadIdx.x * 3 + k]; }
    if (i < n) {
        const float4 qq = __ldg(reinterpret_cast<const float4*>(quats) + i);
        q[0] = qq.x; q[1] = qq.y; q[2] = qq.z; q[3] = qq.w;
    }
    float gm[3] = {0.f, 0.f, 0.f}, gs[3] = {0.f, 0.f, 0.f}, gq[4] = {0.f, 0.f, 0.f, 0.f};
    for (int view = 0; view < n_views; ++view) {
        __syncthreads();
        if (threadIdx.x < 32) load_camera(cam, viewmats, fullmats, intrins, fx, fy, cx, cy, view);
        const long long vbase = (long long)view * n;
        // conics / v_conics of this view through shared memory (coalesced)
        span_load<3>(conics + vbase * 3, first, n, sm_a);
        span_load<3>(v_conics + vbase * 3, first, n, sm_b);
        __syncthreads();
        if (i < n && radii[vbase + i] > 0) {
            float con[3], vcon[3], vxy[2];
#pragma unroll
            for (int k = 0; k < 3; ++k) { con[k] = sm_a[threadIdx.x * 3 + k]; vcon[k] = sm_b[threadIdx.x * 3 + k]; }
            const float2 vx = __ldg(reinterpret_cast<const float2*>(v_xys) + vbase + i);
            vxy[0] = vx.x; vxy[1] = vx.y;
            const float vd = v_depths ? __ldg(v_depths + vbase + i) : 0.0f;
            const ProjGrad g = project_bwd_one(p, s, glob, q, cam, img_h, img_w, con, vxy, vd, vcon);
#pragma unroll
            for (int k = 0; k < 3; ++k) { gm[k] += g.v_mean[k]; gs[k] += g.v_scale[k]; }
#pragma unroll
            for (int k = 0; k < 4; ++k) gq[k] += g.v_quat[k];
        }
    }
    __syncthreads();
    if (accumulate) {
        span_load<3>(v_means, first, n, sm_a);
        span_load<3>(v_scales, first, n, sm_b);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 3; ++k) { sm_a[threadIdx.x * 3 + k] += gm[k]; sm_b[threadIdx.x * 3 + k] += gs[k]; }
    } else {
#pragma unroll
        for (int k = 0; k < 3; ++k) { sm_a[threadIdx.x * 3 + k] = gm[k]; sm_b[threadIdx.x * 3 + k] = gs[k]; }
    }
    __syncthreads();
    span_store<3>(v_means, first, n, sm_a);
    span_store<3>(v_scales, first, n, sm_b);
    if (i < n) {
        float4* dst = reinterpret_cast<float4*>(v_quats) + i;
        float4 r = make_float4(gq[0], gq[1], gq[2], gq[3]);
        if (accumulate) { const float4 o = *dst; r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w; }
        *dst = r;
    }
}

// ---------------------------------------------------------------------------------------------
// Spherical harmonics.  One warp owns 32 consecutive Gaussians; their [32, NB*3] coefficient rows
// are one contiguous span that the warp copies with 16-byte accesses into its shared-memory slab,
// then every lane walks its own row (row stride NB*3 words is odd for NB in {1,9,25} and the
// accesses of a warp then hit 32 distinct banks; NB=4,16 give 12/48 -> 4-way, still off the HBM
// critical path).
// ---------------------------------------------------------------------------------------------
constexpr int kShWarps = 4;

template <bool kBackward>
__global__ void __launch_bounds__(kShWarps * 32)
sh_kernel(int n, int nb, int deg_use, const float* __restrict__ dirs, const float* __restrict__ in,
          float* __restrict__ out) {
    // forward:  in = coeffs [N, nb, 3], out = colors [N, 3]
    // backward: in = v_colors [N, 3],   out = v_coeffs [N, nb, 3]
    extern __shared__ float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = nb * 3;
    float* slab = sm + (size_t)warp * 32 * row;
    const long long first = ((long long)blockIdx.x * kShWarps + warp) * 32;
    if (first >= n) return;
    const long long i = first + lane;
    const int rows_here = (int)min((long long)32, n - first);
    const int nuse = sh_num_bases(deg_use);
    const int span = rows_here * row;
    float* gspan = const_cast<float*>(kBackward ? out : in) + first * row;

    float Y[25];
    float d[3] = {0.f, 0.f, 1.f};
    if (i < n) { d[0] = __ldg(dirs + 3 * i); d[1] = __ldg(dirs + 3 * i + 1); d[2] = __ldg(dirs + 3 * i + 2); }
    sh_basis(deg_use, d[0], d[1], d[2], Y);

    if (!kBackward) {
        // span base is 16B aligned iff (first*row*4) % 16 == 0; first is a multiple of 32 -> always.
        const int nvec = span >> 2;
        const float4* g4 = reinterpret_cast<const float4*>(gspan);
        float4* s4 = reinterpret_cast<float4*>(slab);
        for (int k = lane; k < nvec; k += 32) s4[k] = __ldg(g4 + k);
        for (int k = (nvec << 2) + lane; k < span; k += 32) slab[k] = __ldg(gspan + k);
        __syncwarp();
        if (i < n) {
            const float* cf = slab + lane * row;
            float acc[3] = {0.f, 0.f, 0.f};
            for (int b = 0; b < nuse; ++b) {
#pragma unroll
                for (int c = 0; c < 3; ++c) acc[c] = acc[c] + Y[b] * cf[3 * b + c];
            }
            out[3 * i] = acc[0]; out[3 * i + 1] = acc[1]; out[3 * i + 2] = acc[2];
        }
    } else {
        if (i < n) {
            float v[3] = {__ldg(in + 3 * i), __ldg(in + 3 * i + 1), __ldg(in + 3 * i + 2)};
            float* cf = slab + lane * row;
            for (int b = 0; b < nb; ++b) {
#pragma unroll
                for (int c = 0; c < 3; ++c) cf[3 * b + c] = b < nuse ? Y[b] * v[c] : 0.0f;  // unused bands: +0
            }
        }
        __syncwarp();
        const int nvec = span >> 2;
        float4* g4 = reinterpret_cast<float4*>(gspan);
        const float4* s4 = reinterpret_cast<const float4*>(slab);
        for (int k = lane; k < nvec; k += 32) g4[k] = s4[k];
        for (int k = (nvec << 2) + lane; k < span; k += 32) gspan[k] = slab[k];
    }
}

}  // namespace gg

using namespace gg;

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" int gg_project_fwd_views(int n, int n_views, const float* means, const float* scales, float glob_scale,
                                    const float* quats, const float* viewmats, const float* fullmats,
                                    const float* intrins, float fx, float fy, float cx, float cy, int img_h,
                                    int img_w, int tiles_x, int tiles_y, float clip_thresh, float* cov3d, float* xys,
                                    float* depths, int32_t* radii, float* conics, int32_t* num_tiles_hit,
                                    void* stream) {
    GG_REQUIRE(n >= 1 && n_views >= 1, "gg_project_fwd: need n >= 1 and n_views >= 1");
    GG_REQUIRE(means && scales && quats && viewmats && fullmats, "gg_project_fwd: null input pointer");
    GG_REQUIRE(cov3d && xys && depths && radii && conics && num_tiles_hit, "gg_project_fwd: null output pointer");
    GG_REQUIRE(img_h > 0 && img_w > 0 && tiles_x > 0 && tiles_y > 0, "gg_project_fwd: bad image / tile bounds");
    GG_REQUIRE(((uintptr_t)quats & 15) == 0 && ((uintptr_t)xys & 7) == 0, "gg_project_fwd: quats/xys misaligned");
    dim3 grid(div_up(n, kProjThreads), n_views);
    project_fwd_kernel<<<grid, kProjThreads, 0, (cudaStream_t)stream>>>(
        n, means, scales, glob_scale, quats, viewmats, fullmats, intrins, fx, fy, cx, cy, img_h, img_w, tiles_x,
        tiles_y, clip_thresh, cov3d, xys, depths, radii, conics, num_tiles_hit);
    count_launch();
    return check_launch("project_fwd_kernel");
}

extern "C" int gg_project_fwd(int n, const float* means, const float* scales, float glob_scale, const float* quats,
                              const float* viewmat, const float* fullmat, float fx, float fy, float cx, float cy,
                              int img_h, int img_w, int tiles_x, int tiles_y, float clip_thresh, float* cov3d,
                              float* xys, float* depths, int32_t* radii, float* conics, int32_t* num_tiles_hit,
                              void* stream) {
    return gg_project_fwd_views(n, 1, means, scales, glob_scale, quats, viewmat, fullmat, nullptr, fx, fy, cx, cy,
                                img_h, img_w, tiles_x, tiles_y, clip_thresh, cov3d, xys, depths, radii, conics,
                                num_tiles_hit, stream);
}

extern "C" int gg_project_bwd_views(int n, int n_views, const float* means, const float* scales, float glob_scale,
                                    const float* quats, const float* viewmats, const float* fullmats,
                                    const float* intrins, float fx, float fy, float cx, float cy, int img_h,
                                    int img_w, const int32_t* radii, const float* conics, const float* v_xys,
                                    const float* v_depths, const float* v_conics, int accumulate, float* v_means,
                                    float* v_scales, float* v_quats, void* stream) {
    GG_REQUIRE(n >= 1 && n_views >= 1, "gg_project_bwd: need n >= 1 and n_views >= 1");
    GG_REQUIRE(means && scales && quats && viewmats && fullmats && radii && conics && v_xys && v_conics,
               "gg_project_bwd: null input pointer");
    GG_REQUIRE(v_means && v_scales && v_quats, "gg_project_bwd: null output pointer");
    GG_REQUIRE(((uintptr_t)quats & 15) == 0 && ((uintptr_t)v_quats & 15) == 0 && ((uintptr_t)v_xys & 7) == 0,
               "gg_project_bwd: quats/v_quats/v_xys misaligned");
    project_bwd_kernel<<<div_up(n, kProjThreads), kProjThreads, 0, (cudaStream_t)stream>>>(
        n, n_views, means, scales, glob_scale, quats, viewmats, fullmats, intrins, fx, fy, cx, cy, img_h, img_w,
        radii, conics, v_xys, v_depths, v_conics, accumulate, v_means, v_scales, v_quats);
    count_launch();
    return check_launch("project_bwd_kernel");
}

extern "C" int gg_project_bwd(int n, const float* means, const float* scales, float glob_scale, const float* quats,
                              const float* viewmat, const float* fullmat, float fx, float fy, float cx, float cy,
                              int img_h, int img_w, const int32_t* radii, const float* conics, const float* v_xys,
                              const float* v_depths, const float* v_conics, float* v_means, float* v_scales,
                              float* v_quats, void* stream) {
    return gg_project_bwd_views(n, 1, means, scales, glob_scale, quats, viewmat, fullmat, nullptr, fx, fy, cx, cy,
                                img_h, img_w, radii, conics, v_xys, v_depths, v_conics, 0, v_means, v_scales,
                                v_quats, stream);
}

static int sh_launch(bool backward, int n, int degree, int degrees_to_use, const float* dirs, const float* in,
                     float* out, void* stream) {
    GG_REQUIRE(n >= 1, "gg_sh: need n >= 1");
    GG_REQUIRE(degree >= 0 && degree <= 4 && degrees_to_use >= 0 && degrees_to_use <= degree,
               "gg_sh: need 0 <= degrees_to_use <= degree <= 4");
    GG_REQUIRE(dirs && in && out, "gg_sh: null pointer");
    GG_REQUIRE(((uintptr_t)(backward ? out : in) & 15) == 0, "gg_sh: coefficient array must be 16-byte aligned");
    const int nb = sh_num_bases(degree);
    const size_t smem = (size_t)kShWarps * 32 * nb * 3 * sizeof(float);
    const int grid = div_up(n, kShWarps * 32);
    if (backward)
        sh_kernel<true><<<grid, kShWarps * 32, smem, (cudaStream_t)stream>>>(n, nb, degrees_to_use, dirs, in, out);
    else
        sh_kernel<false><<<grid, kShWarps * 32, smem, (cudaStream_t)stream>>>(n, nb, degrees_to_use, dirs, in, out);
    count_launch();
    return check_launch("sh_kernel");
}

extern "C" int gg_sh_fwd(int n, int degree, int degrees_to_use, const float* dirs, const float* coeffs,
                         float* colors, void* stream) {
    return sh_launch(false, n, degree, degrees_to_use, dirs, coeffs, colors, stream);
}

extern "C" int gg_sh_bwd(int n, int degree, int degrees_to_use, const float* dirs, const float* v_colors,
                         float* v_coeffs, void* stream) {
    return sh_launch(true, n, degree, degrees_to_use, dirs, v_colors, v_coeffs, stream);
}
