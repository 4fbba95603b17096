// The remaining terms of GaussianSplattingModel.get_loss_dict (nerfstudio/models/gaussian_splatting.py:876-925;
// SURVEY 8-f4), each as ONE pass that produces the loss value AND its gradient:
//   gg_geom_loss        depth_loss  = L1(depth[mask], gt_depth[mask])                                   (:880)
//                       normal_loss = 1/2 MSE(normal[:,mask], gt[:,mask]) + 1/2 (1 - mean cos(normal, gt)) (:879)
//                       straight from the blended [.., CP] image (depth = channel 3, normal = 4..6) into the
//                       image's gradient buffer
//   gg_cosine_rows_loss sum_k w_k (1 - cos(a_k, b_k)) over row pairs addressed through optional index lists:
//                       the contrastive feature loss over sampled pixel pairs (:905-912) and the CLIP up-projection
//                       loss over sampled points (:913-914); gradients are scattered back (atomics when indexed)
//   gg_param_regs       sh_reg (:920) and scale_reg (:921-923) over the Gaussians, gradients added to the leaves'
// Reductions are per-block partials summed in block order by the last block to finish (deterministic), like
// gg_pixel_loss.  HBM-bound streaming kernels: one read of every input, one write / update of the gradient.
#include "gg_common.cuh"
#include "gg_b200.h"

namespace gg {

constexpr int kLossBlocks = 1024;
constexpr float kNormEps = 1e-12f;  // F.normalize's eps

// block-level sum of up to kK values per thread -> partial[k * kLossBlocks + block]; the last block adds the
// partials in block order, scales them and writes (or accumulates) out[k]
template <int kK>
__device__ __forceinline__ void finish_sums(float (&acc)[kK], float* partial, unsigned int* counter, float* out,
                                            const float (&scale)[kK], int accumulate) {
    __shared__ float warp_sum[kK][8];
    __shared__ bool last;
#pragma unroll
    for (int k = 0; k < kK; ++k) {
        float v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) warp_sum[k][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < kK; ++k) {
            float s = 0.0f;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += warp_sum[k][w];
            partial[k * kLossBlocks + blockIdx.x] = s;
        }
        __threadfence();
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
        // fixed-order sum of the partials by the whole last block (deterministic)
        __shared__ double tree[256];
        __threadfence();
#pragma unroll
        for (int k = 0; k < kK; ++k) {
            double s = 0.0;
            for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) s += (double)__ldcg(partial + k * kLossBlocks + b);
            tree[threadIdx.x] = s;
            __syncthreads();
            for (int o = 128; o > 0; o >>= 1) {
                if ((int)threadIdx.x < o && threadIdx.x + o < blockDim.x) tree[threadIdx.x] += tree[threadIdx.x + o];
                __syncthreads();
            }
            if (threadIdx.x == 0) {
                const float v = (float)(tree[0] * (double)scale[k]);
                out[k] = accumulate ? out[k] + v : v;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) *counter = 0u;  // ready for the next call
    }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
geom_loss_kernel(long long n_pix, int stride, const float* __restrict__ image, const float* __restrict__ gt_depth,
                 const float* __restrict__ gt_normal, const uint8_t* __restrict__ mask,
                 const int32_t* __restrict__ valid_pixels, float w_depth, float w_normal, float* __restrict__ grad,
                 int grad_stride, int accumulate, float* __restrict__ partial, unsigned int* __restrict__ counter,
                 float* __restrict__ loss) {
    const float cnt = (float)max(*valid_pixels, 1);
    const float inv = 1.0f / cnt;
    float acc[3] = {0.0f, 0.0f, 0.0f};  // sum |d - g|, sum (n - g)^2, sum cos
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n_pix; p += step) {
        float gd = 0.0f, gn[3] = {0.0f, 0.0f, 0.0f};
        if (mask[p]) {
            const float* px = image + p * stride;
            const float d = px[3] - gt_depth[p];
            acc[0] += fabsf(d);
            gd = w_depth * inv * (d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f));
            const float n0 = px[4], n1 = px[5], n2 = px[6];
            const float g0 = gt_normal[3 * p], g1 = gt_normal[3 * p + 1], g2 = gt_normal[3 * p + 2];
            const float e0 = n0 - g0, e1 = n1 - g1, e2 = n2 - g2;
            acc[1] += e0 * e0 + e1 * e1 + e2 * e2;
            // cosine_similarity_loss (:113-118): both sides through F.normalize(dim=0), eps 1e-12
            const float nn = sqrtf(n0 * n0 + n1 * n1 + n2 * n2), ng = sqrtf(g0 * g0 + g1 * g1 + g2 * g2);
            const float in_ = 1.0f / fmaxf(nn, kNormEps), ig = 1.0f / fmaxf(ng, kNormEps);
            const float h0 = g0 * ig, h1 = g1 * ig, h2 = g2 * ig;
            const float u0 = n0 * in_, u1 = n1 * in_, u2 = n2 * in_;
            const float c = u0 * h0 + u1 * h1 + u2 * h2;
            acc[2] += c;
            // d(1/2 MSE)/dn = (n - g) / (3 cnt);  d(-1/2 mean cos)/dn = -1/(2 cnt) (h - c u) / |n|  (0 inside the eps ball)
            const float km = w_normal * inv * (1.0f / 3.0f), kc = nn > kNormEps ? -0.5f * w_normal * inv * in_ : 0.0f;
            gn[0] = km * e0 + kc * (h0 - c * u0);
            gn[1] = km * e1 + kc * (h1 - c * u1);
            gn[2] = km * e2 + kc * (h2 - c * u2);
        }
        float* g = grad + p * grad_stride;
        if (accumulate) { g[3] += gd; g[4] += gn[0]; g[5] += gn[1]; g[6] += gn[2]; }
        else { g[3] = gd; g[4] = gn[0]; g[5] = gn[1]; g[6] = gn[2]; }
    }
    // loss[0] = w_depth * sum|d-g| / cnt;  loss[1] = w_normal * (1/2 sum(e^2) / (3 cnt) + 1/2 (1 - sum cos / cnt))
    float two[2] = {acc[0], acc[1] * (1.0f / 6.0f) - 0.5f * acc[2]};
    const float scale[2] = {w_depth * inv, w_normal * inv};
    finish_sums<2>(two, partial, counter, loss, scale, 0);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        // the constant 1/2 of the cosine term is added by whoever finishes last -- but that thread is not known here;
        // it is added on the host side of the C entry point with a second tiny launch instead (see below)
    }
}

__global__ void add_constant_kernel(float* x, float v) { *x += v; }

// ---------------------------------------------------------------------------------------------
// one warp per row pair; lanes stride over the row
__global__ void __launch_bounds__(256)
cosine_rows_kernel(long long n_rows, int dim, const float* __restrict__ a, long long a_stride,
                   const int64_t* __restrict__ a_idx, const float* __restrict__ b, long long b_stride,
                   const int64_t* __restrict__ b_idx, const float* __restrict__ weight, float* __restrict__ grad_a,
                   long long ga_stride, float* __restrict__ grad_b, long long gb_stride, float* __restrict__ partial,
                   unsigned int* __restrict__ counter, float* __restrict__ loss, int accumulate_loss) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    float acc[1] = {0.0f};
    for (long long k = warp; k < n_rows; k += n_warps) {
        const long long ra = a_idx ? a_idx[k] : k, rb = b_idx ? b_idx[k] : k;
        const float* pa = a + ra * a_stride;
        const float* pb = b + rb * b_stride;
        float saa = 0.0f, sbb = 0.0f, sab = 0.0f;
        for (int c = lane; c < dim; c += 32) {
            const float x = pa[c], y = pb[c];
            saa = fmaf(x, x, saa); sbb = fmaf(y, y, sbb); sab = fmaf(x, y, sab);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            saa += __shfl_xor_sync(0xffffffffu, saa, o);
            sbb += __shfl_xor_sync(0xffffffffu, sbb, o);
            sab += __shfl_xor_sync(0xffffffffu, sab, o);
        }
        const float na = sqrtf(saa), nb = sqrtf(sbb);
        const float ia = 1.0f / fmaxf(na, kNormEps), ib = 1.0f / fmaxf(nb, kNormEps);
        const float cs = sab * ia * ib;
        const float w = weight ? weight[k] : 1.0f;
        if (lane == 0) acc[0] += w * (1.0f - cs);
        // d(-w cos)/da = -w (b^ - cos a^) / |a|   (0 inside the eps ball), same for b
        const float ka = na > kNormEps ? -w * ia : 0.0f, kb = nb > kNormEps ? -w * ib : 0.0f;
        for (int c = lane; c < dim; c += 32) {
            const float x = pa[c] * ia, y = pb[c] * ib;
            if (grad_a) {
                float* g = grad_a + ra * ga_stride + c;
                const float v = ka * (y - cs * x);
                if (a_idx) atomicAdd(g, v); else *g += v;
            }
            if (grad_b) {
                float* g = grad_b + rb * gb_stride + c;
                const float v = kb * (x - cs * y);
                if (b_idx) atomicAdd(g, v); else *g += v;
            }
        }
    }
    const float scale[1] = {1.0f};
    finish_sums<1>(acc, partial, counter, loss, scale, accumulate_loss);
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
param_regs_kernel(long long n, int nb, const float* __restrict__ sh, const float* __restrict__ log_scales,
                  float max_ratio, float w_sh, float w_scale, float* __restrict__ v_sh, float* __restrict__ v_log_scales,
                  float* __restrict__ partial, unsigned int* __restrict__ counter, float* __restrict__ loss) {
    float acc[2] = {0.0f, 0.0f};
    const float inv_n = 1.0f / (float)n;
    const long long step = (long long)gridDim.x * blockDim.x;
    // sh_reg = colors_all[:, 1:, :].norm(dim=1).mean(): one norm per (Gaussian, colour channel) over the bands >= 1
    const long long n_cols = 3 * n;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_cols; t += step) {
        const long long i = t / 3;
        const int c = (int)(t - 3 * i);
        const float* row = sh + i * (long long)nb * 3 + c;
        float ss = 0.0f;
        for (int b = 1; b < nb; ++b) { const float x = row[3 * b]; ss = fmaf(x, x, ss); }
        const float nrm = sqrtf(ss);
        acc[0] += nrm;
        if (v_sh && nrm > 0.0f) {
            const float k = w_sh / (nrm * (float)n_cols);
            float* g = v_sh + i * (long long)nb * 3 + c;
            for (int b = 1; b < nb; ++b) g[3 * b] += k * row[3 * b];
        }
    }
    // scale_reg = 0.1 * mean(max(exp(s).amax / exp(s).amin, R) - R)
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        const float l0 = log_scales[3 * i], l1 = log_scales[3 * i + 1], l2 = log_scales[3 * i + 2];
        const float e0 = expf(l0), e1 = expf(l1), e2 = expf(l2);
        const float hi = fmaxf(fmaxf(e0, e1), e2), lo = fminf(fminf(e0, e1), e2);
        const float r = hi / lo;
        if (r > max_ratio) {
            acc[1] += r - max_ratio;
            if (v_log_scales) {
                // d r / d log s_max = r, d r / d log s_min = -r (amax / amin pass the gradient to the FIRST extremal index)
                const float k = 0.1f * w_scale * inv_n * r;
                const int imax = (e0 >= e1 && e0 >= e2) ? 0 : ((e1 >= e2) ? 1 : 2);
                const int imin = (e0 <= e1 && e0 <= e2) ? 0 : ((e1 <= e2) ? 1 : 2);
                v_log_scales[3 * i + imax] += k;
                v_log_scales[3 * i + imin] -= k;
            }
        }
    }
    const float scale[2] = {w_sh / (float)n_cols, 0.1f * w_scale * inv_n};
    finish_sums<2>(acc, partial, counter, loss, scale, 0);
}

}  // namespace gg

using namespace gg;

extern "C" size_t gg_loss_workspace_bytes(void) { return sizeof(float) * 3 * kLossBlocks + 16; }

static int loss_blocks(long long items, int per_block) {
    int b = div_up(items, per_block);
    return b < 1 ? 1 : (b > kLossBlocks ? kLossBlocks : b);
}

extern "C" int gg_geom_loss(long long n_pixels, int stride, const float* image, const float* gt_depth,
                            const float* gt_normal, const uint8_t* mask, const int32_t* valid_pixels, float w_depth,
                            float w_normal, float* grad, int grad_stride, int accumulate, float* loss, void* workspace,
                            size_t workspace_bytes, void* stream) {
    GG_REQUIRE(n_pixels >= 1 && stride >= 7 && grad_stride >= 7, "gg_geom_loss: the image needs the 7 geometry channels");
    GG_REQUIRE(image && gt_depth && gt_normal && mask && valid_pixels && grad && loss && workspace,
               "gg_geom_loss: null pointer");
    GG_REQUIRE(workspace_bytes >= gg_loss_workspace_bytes(), "gg_geom_loss: workspace too small");
    float* partial = reinterpret_cast<float*>(workspace);
    unsigned int* counter = reinterpret_cast<unsigned int*>(partial + 3 * kLossBlocks);
    cudaStream_t st = (cudaStream_t)stream;
    geom_loss_kernel<<<loss_blocks(n_pixels, 256 * 4), 256, 0, st>>>(n_pixels, stride, image, gt_depth, gt_normal, mask,
                                                                   valid_pixels, w_depth, w_normal, grad, grad_stride,
                                                                   accumulate, partial, counter, loss);
    add_constant_kernel<<<1, 1, 0, st>>>(loss + 1, 0.5f * w_normal);   // the "1 -" of 1/2 (1 - mean cos)
    count_launch(2);
    return check_launch("geom_loss_kernel");
}

extern "C" int gg_cosine_rows_loss(long long n_rows, int dim, const float* a, long long a_stride, const int64_t* a_index,
                                   const float* b, long long b_stride, const int64_t* b_index, const float* weight,
                                   float* grad_a, long long grad_a_stride, float* grad_b, long long grad_b_stride,
                                   float* loss, int accumulate_loss, void* workspace, size_t workspace_bytes,
                                   void* stream) {
    GG_REQUIRE(n_rows >= 0 && dim >= 1, "gg_cosine_rows_loss: bad sizes");
    GG_REQUIRE(loss && workspace && workspace_bytes >= gg_loss_workspace_bytes(), "gg_cosine_rows_loss: null / small workspace");
    if (n_rows == 0) {
        if (!accumulate_loss) GG_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), (cudaStream_t)stream));
        return GG_OK;
    }
    GG_REQUIRE(a && b && a_stride >= (a_index ? 1 : dim) && b_stride >= (b_index ? 1 : dim), "gg_cosine_rows_loss: bad operands");
    float* partial = reinterpret_cast<float*>(workspace);
    unsigned int* counter = reinterpret_cast<unsigned int*>(partial + 3 * kLossBlocks);
    cosine_rows_kernel<<<loss_blocks(n_rows, 8), 256, 0, (cudaStream_t)stream>>>(
        n_rows, dim, a, a_stride, a_index, b, b_stride, b_index, weight, grad_a, grad_a_stride, grad_b, grad_b_stride,
        partial, counter, loss, accumulate_loss);
    count_launch();
    return check_launch("cosine_rows_kernel");
}

extern "C" int gg_param_regs(long long n, int num_bases, const float* sh_coeffs, const float* log_scales,
                             float max_gauss_ratio, float w_sh, float w_scale, float* v_sh_coeffs, float* v_log_scales,
                             float* loss, void* workspace, size_t workspace_bytes, void* stream) {
    GG_REQUIRE(n >= 1 && num_bases >= 1, "gg_param_regs: bad sizes");
    GG_REQUIRE(sh_coeffs && log_scales && loss && workspace, "gg_param_regs: null pointer");
    GG_REQUIRE(workspace_bytes >= gg_loss_workspace_bytes(), "gg_param_regs: workspace too small");
    float* partial = reinterpret_cast<float*>(workspace);
    unsigned int* counter = reinterpret_cast<unsigned int*>(partial + 3 * kLossBlocks);
    param_regs_kernel<<<loss_blocks(3 * n, 256 * 4), 256, 0, (cudaStream_t)stream>>>(
        n, num_bases, sh_coeffs, log_scales, max_gauss_ratio, w_sh, w_scale, v_sh_coeffs, v_log_scales, partial, counter,
        loss);
    count_launch();
    return check_launch("param_regs_kernel");
}
