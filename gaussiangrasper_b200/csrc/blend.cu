// Forward / backward alpha compositing over 16x16 tiles, any channel count, all channels in one
// pass.  Replaces gsplat 0.1.0 rasterize_forward/backward_kernel and nd_rasterize_forward/
// backward_kernel (RasterizeGaussians / NDRasterizeGaussians, called from
// nerfstudio/models/gaussian_splatting.py:735,747,759,773).
//
// Data layout: the per-Gaussian 2D geometry is consumed as packed 32-byte records
//   geo[g] = { x, y, A/2, B, C/2, opacity, tau, 0 }      (one DRAM sector per gather)
// where (A,B,C) is the conic and tau = ln(255*opacity) + margin is the largest sigma for which
// alpha = opacity*exp(-sigma) can still reach 1/255 (pairs beyond it are skipped before the
// exp; pairs inside the margin still take the exact alpha test, so results are unchanged).
// Channel rows are gathered from colors[g * stride + 0..C).  One CTA per tile; entries are staged
// into double-buffered shared memory with cp.async (LDGSTS) while the previous batch is blended.
//
// Backward: per-pixel back-to-front replay; the C+6 partial gradients of one Gaussian are
// reduced across the warp with a transposed butterfly (31 shuffles for 32 values instead of
// 32x5) after which lane l owns component l and issues a single red.global.add -- one atomic
// per warp per component instead of one per pixel.
#include "gg_common.cuh"
#include "gg_b200.h"

namespace gg {

constexpr int kBlendThreads = 256;
constexpr float kAlphaMin = 1.0f / 255.0f;
constexpr float kAlphaMax = 0.999f;
constexpr float kTStop = 1e-4f;
constexpr float kTauMargin = GG_TAU_MARGIN;

struct BlendArgs {
    int channels;        // real channel count handled by this launch (<= CP)
    int color_stride;    // floats between consecutive rows of colors / v_colors
    int out_stride;      // floats between consecutive pixels of out / v_out
    int img_h, img_w, tiles_x, tiles_y;
    long long geo_view_stride;    // rows between views in geo / v_geo (n)
    long long color_view_stride;  // rows between views in colors / v_colors (0: shared by all views)
    const int32_t* ids_sorted;
    const int32_t* tile_ranges;  // [V*T, 2]
    const float* geo;            // [V*N, 8]
    const float* colors;
    const float* bg;             // [channels]
    float* out;                  // [V, H, W, out_stride]
    float* final_T;              // [V, H, W]
    int32_t* final_idx;          // [V, H, W]
    unsigned long long* pair_counter;  // optional: += pixel-Gaussian pairs visited
    // backward only
    const float* v_out;
    float* v_geo;     // [V*N, 8]  (+= v_x, v_y, v_A, v_B, v_C, v_opacity)
    float* v_colors;  // same indexing as colors
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// sigma and alpha exactly as evaluated by both the forward and the backward kernel
__device__ __forceinline__ float eval_sigma(float dx, float dy, float hA, float B, float hC) {
    const float u = __fmaf_rn(hA, dx, __fmul_rn(B, dy));
    return __fmaf_rn(dx, u, __fmul_rn(__fmul_rn(hC, dy), dy));
}

// Thread -> pixel inside the 16x16 tile: every warp owns a compact 8x4 block (2 x 4 blocks per
// tile), which a Gaussian's footprint either misses entirely or covers with many lanes.
__device__ __forceinline__ void tile_pixel(int& tx, int& ty) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    tx = (lane & 7) | ((warp & 1) << 3);
    ty = (lane >> 3) | ((warp >> 1) << 2);
}

// Stage entries [first, first+cnt) of the sorted list into shared memory.
template <int CP, int BATCH, bool kVec>
__device__ __forceinline__ void stage_batch(const BlendArgs& a, long long geo_base, long long color_base, int first,
                                            int cnt, float* geo_sm, float* col_sm) {
    constexpr int kParts = kBlendThreads / BATCH;  // threads cooperating on one entry
    const int e = threadIdx.x % BATCH;
    const int part = threadIdx.x / BATCH;
    if (e < cnt) {
        const int g = __ldg(a.ids_sorted + first + e);
        const float* grow = a.geo + (geo_base + g) * 8;
        const float* crow = a.colors + (color_base + g) * (long long)a.color_stride;
        float* gdst = geo_sm + e * 8;
        float* cdst = col_sm + e * CP;
        if (kVec) {
            constexpr int kChunks = CP / 4 + 2;
            const int cchunks = (a.channels + 3) >> 2;
            for (int q = part; q < kChunks; q += kParts) {
                if (q < 2) cp_async16(gdst + 4 * q, grow + 4 * q);
                else if (q - 2 < cchunks) cp_async16(cdst + 4 * (q - 2), crow + 4 * (q - 2));
            }
        } else {
            if (part == 0) { cp_async16(gdst, grow); cp_async16(gdst + 4, grow + 4); }
            for (int c = part; c < a.channels; c += kParts) cp_async4(cdst + c, crow + c);
        }
    }
}

template <int CP, int BATCH, bool kVec>
__global__ void __launch_bounds__(kBlendThreads)
blend_fwd_kernel(const BlendArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* geo_sm = smem;                      // [2][BATCH][8]
    float* col_sm = smem + 2 * BATCH * 8;      // [2][BATCH][CP]
    const int view = blockIdx.y;
    const int tile = blockIdx.x;
    const int tile_y = tile / a.tiles_x, tile_x = tile - tile_y * a.tiles_x;
    int tx, ty;
    tile_pixel(tx, ty);
    const int px = tile_x * GG_TILE + tx, py = tile_y * GG_TILE + ty;
    const bool inside = px < a.img_w && py < a.img_h;
    const float fpx = (float)px, fpy = (float)py;
    const int2 range = __ldg(reinterpret_cast<const int2*>(a.tile_ranges) + (long long)view * a.tiles_x * a.tiles_y + tile);
    const long long geo_base = (long long)view * a.geo_view_stride;
    const long long color_base = (long long)view * a.color_view_stride;

    float acc[CP];
#pragma unroll
    for (int c = 0; c < CP; ++c) acc[c] = 0.0f;
    float T = 1.0f;
    int last = range.x;
    bool done = !inside;
    long long n_vis = 0;

    const int total = range.y - range.x;
    const int nb = (total + BATCH - 1) / BATCH;
    if (nb > 0) {
        stage_batch<CP, BATCH, kVec>(a, geo_base, color_base, range.x, min(BATCH, total), geo_sm, col_sm);
        cp_async_commit();
    }
    for (int b = 0; b < nb; ++b) {
        const int buf = b & 1;
        if (b + 1 < nb) {
            const int first = range.x + (b + 1) * BATCH;
            stage_batch<CP, BATCH, kVec>(a, geo_base, color_base, first, min(BATCH, range.y - first),
                                         geo_sm + (buf ^ 1) * BATCH * 8, col_sm + (buf ^ 1) * BATCH * CP);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const int first = range.x + b * BATCH;
        const int cnt = min(BATCH, range.y - first);
        if (!done) {
            const float4* g4 = reinterpret_cast<const float4*>(geo_sm + buf * BATCH * 8);
            const float4* c4 = reinterpret_cast<const float4*>(col_sm + buf * BATCH * CP);
            int e_end = cnt;
            for (int e = 0; e < cnt; ++e) {
                const float4 ga = g4[2 * e], gb = g4[2 * e + 1];
                const float dx = ga.x - fpx, dy = ga.y - fpy;
                const float sigma = eval_sigma(dx, dy, ga.z, ga.w, gb.x);
                if (sigma < 0.0f || sigma > gb.z) continue;
                const float alpha = fminf(kAlphaMax, gb.y * __expf(-sigma));
                if (alpha < kAlphaMin) continue;
                const float next_T = T * (1.0f - alpha);
                if (next_T <= kTStop) { done = true; e_end = e + 1; break; }
                const float vis = alpha * T;
#pragma unroll
                for (int q = 0; q < CP / 4; ++q) {
                    const float4 cc = c4[e * (CP / 4) + q];
                    acc[4 * q] = fmaf(vis, cc.x, acc[4 * q]);
                    acc[4 * q + 1] = fmaf(vis, cc.y, acc[4 * q + 1]);
                    acc[4 * q + 2] = fmaf(vis, cc.z, acc[4 * q + 2]);
                    acc[4 * q + 3] = fmaf(vis, cc.w, acc[4 * q + 3]);
                }
                T = next_T;
                last = first + e + 1;
            }
            n_vis += e_end;
        }
        if (__syncthreads_count(done) == kBlendThreads) break;
    }
    cp_async_wait<0>();
    if (inside) {
        const long long pix = ((long long)view * a.img_h + py) * a.img_w + px;
        float* o = a.out + pix * a.out_stride;
#pragma unroll
        for (int c = 0; c < CP; ++c)
            if (c < a.channels) o[c] = fmaf(T, __ldg(a.bg + c), acc[c]);
        a.final_T[pix] = T;
        a.final_idx[pix] = last;
    }
    if (a.pair_counter) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n_vis += __shfl_xor_sync(0xffffffffu, n_vis, o);
        if ((threadIdx.x & 31) == 0 && n_vis) atomicAdd(a.pair_counter, (unsigned long long)n_vis);
    }
}

// ---------------------------------------------------------------------------------------------
// Transposed warp reduction: v holds NV (power of two <= 32) per-lane partials; on return v[0]
// is the warp total of component (lane >> log2(32/NV)).
// ---------------------------------------------------------------------------------------------
template <int N, int OFF>
struct TreeReduce {
    template <int NV>
    static __device__ __forceinline__ void run(float (&v)[NV], int lane) {
        if constexpr (N > 1) {
            constexpr int H = N / 2;
            const bool up = (lane & OFF) != 0;
#pragma unroll
            for (int i = 0; i < H; ++i) {
                const float send = up ? v[i] : v[i + H];
                const float keep = up ? v[i + H] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
            }
            if constexpr (OFF > 1) TreeReduce<H, OFF / 2>::run(v, lane);
        } else {
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], OFF);
            if constexpr (OFF > 1) TreeReduce<1, OFF / 2>::run(v, lane);
        }
    }
};

__host__ __device__ constexpr int next_pow2(int x) { int p = 1; while (p < x) p <<= 1; return p; }

// Reduce NREAL (<=32) values across the warp and add component k to dst(k) with one red per
// component.  `addr(k)` returns the destination of component k (nullptr to skip).
template <int NREAL, typename AddrFn>
__device__ __forceinline__ void reduce_and_add(float (&v)[next_pow2(NREAL)], int lane, AddrFn addr) {
    constexpr int NV = next_pow2(NREAL);
    constexpr int REP = 32 / NV;  // lanes holding the same component
    TreeReduce<NV, 16>::run(v, lane);
    const int comp = lane / REP;
    if ((lane % REP) == 0 && comp < NREAL) {
        float* p = addr(comp);
        if (p && v[0] != 0.0f) atomicAdd(p, v[0]);
    }
}

// channels [START, CP) of the colour gradient, 32 per tree
template <int START, int CP>
__device__ __forceinline__ void reduce_rest(const float (&vo)[CP], float fac, int lane, float* vc, int nch) {
    if constexpr (START < CP) {
        constexpr int CNT = (CP - START) >= 32 ? 32 : (CP - START);
        constexpr int NV = next_pow2(CNT);
        float r2[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) r2[k] = (k < CNT) ? fac * vo[k < CNT ? START + k : START] : 0.0f;
        reduce_and_add<CNT>(r2, lane, [&](int k) -> float* { return START + k < nch ? vc + START + k : nullptr; });
        reduce_rest<START + CNT, CP>(vo, fac, lane, vc, nch);
    }
}

template <int CP, int BATCH, bool kVec>
__global__ void __launch_bounds__(kBlendThreads)
blend_bwd_kernel(const BlendArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* geo_sm = smem;
    float* col_sm = smem + 2 * BATCH * 8;
    int* ids_sm = reinterpret_cast<int*>(smem + 2 * BATCH * (8 + CP));  // [2][BATCH]
    __shared__ int s_max[kBlendThreads / 32];
    const int view = blockIdx.y;
    const int tile = blockIdx.x;
    const int tile_y = tile / a.tiles_x, tile_x = tile - tile_y * a.tiles_x;
    int tx, ty;
    tile_pixel(tx, ty);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int px = tile_x * GG_TILE + tx, py = tile_y * GG_TILE + ty;
    const bool inside = px < a.img_w && py < a.img_h;
    const float fpx = (float)px, fpy = (float)py;
    const int2 range = __ldg(reinterpret_cast<const int2*>(a.tile_ranges) + (long long)view * a.tiles_x * a.tiles_y + tile);
    const long long geo_base = (long long)view * a.geo_view_stride;
    const long long color_base = (long long)view * a.color_view_stride;
    const long long pix = ((long long)view * a.img_h + py) * a.img_w + px;

    // padding lanes of the staged colour rows are read by the gradient sums: keep them finite
    for (int k = threadIdx.x; k < 2 * BATCH * CP; k += kBlendThreads) col_sm[k] = 0.0f;

    float vo[CP];
    float T_final = 1.0f, bgdot = 0.0f;
    int last = range.x;
#pragma unroll
    for (int c = 0; c < CP; ++c) vo[c] = 0.0f;
    if (inside) {
        T_final = a.final_T[pix];
        last = a.final_idx[pix];
        const float* v = a.v_out + pix * a.out_stride;
#pragma unroll
        for (int c = 0; c < CP; ++c)
            if (c < a.channels) { vo[c] = __ldg(v + c); bgdot = fmaf(__ldg(a.bg + c), vo[c], bgdot); }
    }
    float T = T_final;
    // R = T_final * <bg, v_out> + sum over the entries behind the current one of fac * <colour, v_out>:
    // the only state the alpha gradient needs from "behind" (replaces C running colour sums)
    float R = T_final * bgdot;
    // tile-wide last contributing index
    int wmax = last;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if (lane == 0) s_max[warp] = wmax;
    __syncthreads();
    int bmax = range.x;
#pragma unroll
    for (int w = 0; w < kBlendThreads / 32; ++w) bmax = max(bmax, s_max[w]);
    const int total = bmax - range.x;
    const int nb = (total + BATCH - 1) / BATCH;
    if (nb == 0) return;

    auto stage = [&](int b, int buf) {
        const int first = range.x + b * BATCH;
        const int cnt = min(BATCH, bmax - first);
        if (threadIdx.x < cnt) ids_sm[buf * BATCH + threadIdx.x] = __ldg(a.ids_sorted + first + threadIdx.x);
        stage_batch<CP, BATCH, kVec>(a, geo_base, color_base, first, cnt, geo_sm + buf * BATCH * 8,
                                     col_sm + buf * BATCH * CP);
        cp_async_commit();
    };
    stage(nb - 1, (nb - 1) & 1);
    for (int b = nb - 1; b >= 0; --b) {
        const int buf = b & 1;
        if (b > 0) { stage(b - 1, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();
        const int first = range.x + b * BATCH;
        const int cnt = min(BATCH, bmax - first);
        const float4* g4 = reinterpret_cast<const float4*>(geo_sm + buf * BATCH * 8);
        const float4* c4 = reinterpret_cast<const float4*>(col_sm + buf * BATCH * CP);
        const int* idb = ids_sm + buf * BATCH;
        // entries at or beyond the warp's own last contributor are dead for the whole warp
        const int e_hi = min(cnt, wmax - first);
        for (int e = e_hi - 1; e >= 0; --e) {
            const float4 ga = g4[2 * e], gb = g4[2 * e + 1];
            const float dx = ga.x - fpx, dy = ga.y - fpy;
            const float sigma = eval_sigma(dx, dy, ga.z, ga.w, gb.x);
            // branch-free replay of the forward test (same arithmetic as blend_fwd_kernel)
            const float vis = __expf(-sigma);
            const float araw = gb.y * vis;
            const float alpha = fminf(kAlphaMax, araw);
            const bool valid = (first + e < last) && !(sigma < 0.0f || sigma > gb.z) && (alpha >= kAlphaMin);
            if (!__any_sync(0xffffffffu, valid)) continue;
            constexpr int NV0 = (CP + 6 <= 32) ? next_pow2(CP + 6) : 32;  // first tree: 6 geo + channels
            constexpr int C0 = (CP + 6 <= 32) ? CP : 26;                  // channels carried by tree 0
            float r[NV0];
#pragma unroll
            for (int k = 0; k < NV0; ++k) r[k] = 0.0f;
            float fac = 0.0f;
            if (valid) {
                const float ra = 1.0f / (1.0f - alpha);
                T *= ra;  // transmittance in front of this entry
                fac = alpha * T;
                float dot = 0.0f;
#pragma unroll
                for (int q = 0; q < CP / 4; ++q) {
                    const float4 cc = c4[e * (CP / 4) + q];
                    dot = fmaf(cc.x, vo[4 * q], dot);
                    dot = fmaf(cc.y, vo[4 * q + 1], dot);
                    dot = fmaf(cc.z, vo[4 * q + 2], dot);
                    dot = fmaf(cc.w, vo[4 * q + 3], dot);
                }
                const float v_alpha = fmaf(dot, T, -R * ra);
                R = fmaf(fac, dot, R);
                if (araw <= kAlphaMax) {  // a clamped alpha passes no gradient to sigma / opacity
                    const float v_sigma = -alpha * v_alpha;
                    // conic as stored by the caller is (A, B, C); geo holds (A/2, B, C/2)
                    r[0] = v_sigma * fmaf(2.0f * ga.z, dx, ga.w * dy);
                    r[1] = v_sigma * fmaf(ga.w, dx, 2.0f * gb.x * dy);
                    r[2] = 0.5f * v_sigma * dx * dx;
                    r[3] = v_sigma * dx * dy;
                    r[4] = 0.5f * v_sigma * dy * dy;
                    r[5] = vis * v_alpha;
                }
            }
#pragma unroll
            for (int c = 0; c < C0; ++c) r[6 + c] = fac * vo[c];
            const int g = idb[e];
            float* vg = a.v_geo + (geo_base + g) * 8;
            float* vc = a.v_colors + (color_base + g) * (long long)a.color_stride;
            const int nch = a.channels;
            reduce_and_add<C0 + 6>(r, lane, [&](int k) -> float* {
                return k < 6 ? vg + k : (k - 6 < nch ? vc + (k - 6) : nullptr);
            });
            reduce_rest<C0, CP>(vo, fac, lane, vc, nch);
        }
        __syncthreads();
    }
}

// geo[g] = {x, y, A/2, B, C/2, o, tau, 0};  one thread per row
__global__ void __launch_bounds__(256)
pack_geo_kernel(long long rows, long long n, const float* __restrict__ xys, const float* __restrict__ conics,
                const float* __restrict__ opac, int opac_per_view, float* __restrict__ geo) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const float2 c = __ldg(reinterpret_cast<const float2*>(xys) + i);
    const float A = __ldg(conics + 3 * i), B = __ldg(conics + 3 * i + 1), C = __ldg(conics + 3 * i + 2);
    const float o = __ldg(opac + (opac_per_view ? i : i % n));
    const float tau = (o * 255.0f > 1.0f) ? __logf(o * 255.0f) + kTauMargin : -1.0f;
    float4* dst = reinterpret_cast<float4*>(geo) + 2 * i;
    dst[0] = make_float4(c.x, c.y, 0.5f * A, B);
    dst[1] = make_float4(0.5f * C, o, tau, 0.0f);
}

// v_geo[V*N, 8] -> v_xys [V*N,2], v_conics [V*N,3] (per view) and v_opac [N] (summed over views)
__global__ void __launch_bounds__(256)
unpack_vgeo_kernel(long long n, int n_views, const float* __restrict__ v_geo, float* __restrict__ v_xys,
                   float* __restrict__ v_conics, float* __restrict__ v_opac, int accumulate_opac) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    float vo = 0.0f;
    for (int v = 0; v < n_views; ++v) {
        const long long i = (long long)v * n + g;
        const float4 a = __ldg(reinterpret_cast<const float4*>(v_geo) + 2 * i);
        const float4 b = __ldg(reinterpret_cast<const float4*>(v_geo) + 2 * i + 1);
        reinterpret_cast<float2*>(v_xys)[i] = make_float2(a.x, a.y);
        v_conics[3 * i] = a.z; v_conics[3 * i + 1] = a.w; v_conics[3 * i + 2] = b.x;
        vo += b.y;
    }
    if (v_opac) v_opac[g] = accumulate_opac ? v_opac[g] + vo : vo;
}

template <int CP, int BATCH>
static int launch_blend(bool backward, const BlendArgs& a, int n_views, bool vec, cudaStream_t st) {
    dim3 grid(a.tiles_x * a.tiles_y, n_views);
    size_t smem = sizeof(float) * 2 * BATCH * (8 + CP);
    if (backward) smem += sizeof(int) * 2 * BATCH;
    if (backward) {
        if (vec) blend_bwd_kernel<CP, BATCH, true><<<grid, kBlendThreads, smem, st>>>(a);
        else blend_bwd_kernel<CP, BATCH, false><<<grid, kBlendThreads, smem, st>>>(a);
    } else {
        if (vec) blend_fwd_kernel<CP, BATCH, true><<<grid, kBlendThreads, smem, st>>>(a);
        else blend_fwd_kernel<CP, BATCH, false><<<grid, kBlendThreads, smem, st>>>(a);
    }
    count_launch();
    return check_launch(backward ? "blend_bwd_kernel" : "blend_fwd_kernel");
}

static int dispatch_blend(bool backward, const BlendArgs& a, int n_views, cudaStream_t st) {
    const bool vec = (a.color_stride % 4 == 0) && (((uintptr_t)a.colors & 15) == 0);
    const int c = a.channels;
    if (c <= 4) return launch_blend<4, 128>(backward, a, n_views, vec, st);
    if (c <= 8) return launch_blend<8, 128>(backward, a, n_views, vec, st);
    if (c <= 12) return launch_blend<12, 128>(backward, a, n_views, vec, st);
    if (c <= 16) return launch_blend<16, 128>(backward, a, n_views, vec, st);
    if (c <= 24) return launch_blend<24, 128>(backward, a, n_views, vec, st);
    if (c <= 32) return launch_blend<32, 128>(backward, a, n_views, vec, st);
    if (c <= 40) return launch_blend<40, 64>(backward, a, n_views, vec, st);
    if (c <= 48) return launch_blend<48, 64>(backward, a, n_views, vec, st);
    if (c <= 64) return launch_blend<64, 64>(backward, a, n_views, vec, st);
    set_error("gg_blend: more than 64 channels per launch; split the channel range");
    return GG_ERR_CHANNELS;
}

}  // namespace gg

using namespace gg;

extern "C" int gg_blend_max_channels(void) { return 64; }

extern "C" int gg_pack_geo(long long n, int n_views, const float* xys, const float* conics, const float* opac,
                           int opac_per_view, float* geo, void* stream) {
    GG_REQUIRE(n >= 1 && n_views >= 1, "gg_pack_geo: need n >= 1");
    GG_REQUIRE(xys && conics && opac && geo, "gg_pack_geo: null pointer");
    GG_REQUIRE(((uintptr_t)geo & 15) == 0 && ((uintptr_t)xys & 7) == 0, "gg_pack_geo: misaligned");
    const long long rows = n * n_views;
    pack_geo_kernel<<<div_up(rows, 256), 256, 0, (cudaStream_t)stream>>>(rows, n, xys, conics, opac, opac_per_view, geo);
    count_launch();
    return check_launch("pack_geo_kernel");
}

extern "C" int gg_unpack_vgeo(long long n, int n_views, const float* v_geo, float* v_xys, float* v_conics,
                              float* v_opac, int accumulate_opac, void* stream) {
    GG_REQUIRE(n >= 1 && n_views >= 1, "gg_unpack_vgeo: need n >= 1");
    GG_REQUIRE(v_geo && v_xys && v_conics, "gg_unpack_vgeo: null pointer");
    unpack_vgeo_kernel<<<div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(n, n_views, v_geo, v_xys, v_conics, v_opac,
                                                                         accumulate_opac);
    count_launch();
    return check_launch("unpack_vgeo_kernel");
}

extern "C" int gg_blend_fwd(int n_views, long long n, int channels, int color_stride, int colors_per_view,
                            int out_stride, int img_h, int img_w, int tiles_x, int tiles_y,
                            const int32_t* ids_sorted, const int32_t* tile_ranges, const float* geo,
                            const float* colors, const float* bg, float* out, float* final_T, int32_t* final_idx,
                            unsigned long long* pair_counter, void* stream) {
    GG_REQUIRE(n_views >= 1 && n >= 1 && channels >= 1, "gg_blend_fwd: bad sizes");
    GG_REQUIRE(color_stride >= channels && out_stride >= channels, "gg_blend_fwd: stride smaller than channels");
    GG_REQUIRE(img_h > 0 && img_w > 0 && tiles_x == (img_w + GG_TILE - 1) / GG_TILE &&
                   tiles_y == (img_h + GG_TILE - 1) / GG_TILE,
               "gg_blend_fwd: tile bounds must be ceil(size/16)");
    GG_REQUIRE(ids_sorted && tile_ranges && geo && colors && bg && out && final_T && final_idx,
               "gg_blend_fwd: null pointer");
    GG_REQUIRE(((uintptr_t)geo & 15) == 0 && ((uintptr_t)tile_ranges & 7) == 0, "gg_blend_fwd: misaligned");
    BlendArgs a{};
    a.channels = channels; a.color_stride = color_stride; a.out_stride = out_stride;
    a.img_h = img_h; a.img_w = img_w; a.tiles_x = tiles_x; a.tiles_y = tiles_y;
    a.geo_view_stride = n; a.color_view_stride = colors_per_view ? n : 0;
    a.ids_sorted = ids_sorted; a.tile_ranges = tile_ranges; a.geo = geo; a.colors = colors; a.bg = bg;
    a.out = out; a.final_T = final_T; a.final_idx = final_idx; a.pair_counter = pair_counter;
    return dispatch_blend(false, a, n_views, (cudaStream_t)stream);
}

extern "C" int gg_blend_bwd(int n_views, long long n, int channels, int color_stride, int colors_per_view,
                            int out_stride, int img_h, int img_w, int tiles_x, int tiles_y,
                            const int32_t* ids_sorted, const int32_t* tile_ranges, const float* geo,
                            const float* colors, const float* bg, const float* final_T, const int32_t* final_idx,
                            const float* v_out, float* v_geo, float* v_colors, void* stream) {
    GG_REQUIRE(n_views >= 1 && n >= 1 && channels >= 1, "gg_blend_bwd: bad sizes");
    GG_REQUIRE(color_stride >= channels && out_stride >= channels, "gg_blend_bwd: stride smaller than channels");
    GG_REQUIRE(img_h > 0 && img_w > 0 && tiles_x == (img_w + GG_TILE - 1) / GG_TILE &&
                   tiles_y == (img_h + GG_TILE - 1) / GG_TILE,
               "gg_blend_bwd: tile bounds must be ceil(size/16)");
    GG_REQUIRE(ids_sorted && tile_ranges && geo && colors && bg && final_T && final_idx && v_out && v_geo && v_colors,
               "gg_blend_bwd: null pointer");
    GG_REQUIRE(((uintptr_t)geo & 15) == 0 && ((uintptr_t)tile_ranges & 7) == 0, "gg_blend_bwd: misaligned");
    BlendArgs a{};
    a.channels = channels; a.color_stride = color_stride; a.out_stride = out_stride;
    a.img_h = img_h; a.img_w = img_w; a.tiles_x = tiles_x; a.tiles_y = tiles_y;
    a.geo_view_stride = n; a.color_view_stride = colors_per_view ? n : 0;
    a.ids_sorted = ids_sorted; a.tile_ranges = tile_ranges; a.geo = geo; a.colors = colors; a.bg = bg;
    a.final_T = const_cast<float*>(final_T); a.final_idx = const_cast<int32_t*>(final_idx);
    a.v_out = v_out; a.v_geo = v_geo; a.v_colors = v_colors;
    return dispatch_blend(true, a, n_views, (cudaStream_t)stream);
}
