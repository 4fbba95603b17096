// Forward / backward alpha compositing over 16x16 tiles, any channel count, all channels in one
// pass.  Replaces gsplat 0.1.0 rasterize_forward/backward_kernel and nd_rasterize_forward/
// backward_kernel (RasterizeGaussians / NDRasterizeGaussians, called from
// nerfstudio/models/gaussian_splatting.py:735,747,759,773).
//
// Data layout: the per-Gaussian 2D geometry is consumed as packed 32-byte records
//   geo[g] = { x, y, A/2, B, C/2, opacity, tau, rcut2 }  (one DRAM sector per gather)
// where (A,B,C) is the conic and tau = ln(255*opacity) + margin is the largest sigma for which
// alpha = opacity*exp(-sigma) can still reach 1/255 (pairs beyond it are skipped before the
// exp; pairs inside the margin still take the exact alpha test, so results are unchanged).
// Channel rows are gathered from colors[g * stride + 0..C).  One CTA per tile; entries are staged
// into double-buffered shared memory with cp.async (LDGSTS) while the previous batch is blended.
//
// Backward (blend_bwd_warp_kernel, the default): every warp replays, back to front, only the entries the
// forward recorded as contributing to its 8x4 pixel block (hit masks), gathers their rows itself and never
// synchronises with another warp; per (pixel, entry) it stores two scalars in a warp-private shared-memory
// matrix, and every 16 entries lane j turns entry j's column into the C+6 gradient components (transposed
// accumulation) and issues one red.global.add per component -- one atomic per warp and entry instead of one
// per pixel.  blend_bwd_kernel is the CTA-staged fallback for launches without hit masks (C > 64 column blocks).
#include "gg_common.cuh"
#include "gg_geo.cuh"
#include "gg_tma.cuh"
#include "gg_b200.h"

#ifndef GG_BWD_DOT4
#define GG_BWD_DOT4 0
#endif
#ifndef GG_BWD_B_UNROLL
#define GG_BWD_B_UNROLL 16
#endif
// phase B of the warp backward (v_colour = fac . v_out) on mma.sync m16n8k8 TF32 with the 3xTF32 split: measured
// 0.396 -> 0.340 ms at config 1 with unchanged parity (profiles/r02_blend_bwd_variants.txt).  0 = FP32-pipe version.
#ifndef GG_BWD_MMA
#define GG_BWD_MMA 1
#endif
// colour accumulation of the forward on mma.sync TF32 (3xTF32 split), see blend_fwd_kernel
#ifndef GG_FWD_MMA
#define GG_FWD_MMA 0
#endif
// forward staging ring synchronised by mbarriers (data-ready / buffer-free) instead of one CTA barrier per batch:
// a warp only waits for the copies and for the buffer it refills, so it may run one batch ahead of the slowest.
// Measured and killed: 0.2048 -> 0.2281 ms at config 1 with parity unchanged (profiles/r02_blend_bwd_variants.txt).
#ifndef GG_FWD_RING
#define GG_FWD_RING 0
#endif
#ifndef GG_BWD_MIN_BLOCKS
#define GG_BWD_MIN_BLOCKS 5
#endif

namespace gg {

constexpr int kBwdBUnroll = GG_BWD_B_UNROLL;
constexpr int kBlendThreads = 256;
constexpr float kAlphaMin = 1.0f / 255.0f;
constexpr float kAlphaMax = 0.999f;
constexpr float kTStop = 1e-4f;
constexpr int kStages = 3;  // staging ring depth of the blend kernels


struct BlendArgs {
    int channels;        // real channel count handled by this launch (<= CP)
    int color_stride;    // floats between consecutive rows of colors / v_colors
    int out_stride;      // floats between consecutive pixels of out / v_out
    int img_h, img_w, tiles_x, tiles_y;
    long long geo_view_stride;    // rows between views in geo / v_geo (n)
    long long color_view_stride;  // rows between views in colors / v_colors (0: shared by all views)
    const int32_t* ids_sorted;
    const int32_t* tile_ranges;  // [V*T, 2]
    const int32_t* tile_order;   // [V*T] visiting order of the tiles (nullable: identity)
    const float* geo;            // [V*N, 8]
    const float* colors;
    const float* bg;             // [channels]
    float* out;                  // [V, H, W, out_stride]
    float* final_T;              // [V, H, W]
    int32_t* final_idx;          // [V, H, W]
    unsigned long long* pair_counter;  // optional: += pixel-Gaussian pairs visited
    // Per (tile batch, warp) bit masks of the entries to which at least one pixel of the warp's 8x4
    // block contributed, written by the forward and consumed by the backward (which then neither
    // culls nor tests the others).  Record r = range.x / fwd_batch + global tile + batch holds
    // 8 warps x (fwd_batch / 32) words; the index is collision free because
    // floor(a) + ceil(b) <= floor(a + b) + 1.  Nullable: the backward then culls on its own.
    uint32_t* hit_words;
    int fwd_batch;
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
// this thread's earlier cp.async copies arrive on the mbarrier when they have landed
__device__ __forceinline__ void cp_async_arrive_on(uint64_t* bar) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(b) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(b) : "memory");
}
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

// Conservative per-warp culling of a staged batch: lane l tests entries l, l+32, ... against the
// warp's 8x4 pixel rectangle.  The test is the exact minimum of sigma(d) = d^T Q d / 2 over the
// (continuous) rectangle: 0 if the centre is inside, otherwise the smallest of the four edge
// minima (sigma is convex, so the minimum sits on the boundary; on an edge it is a clamped 1-D
// parabola).  An entry is kept iff that minimum is <= tau, i.e. iff some pixel of the warp could
// pass the per-pixel test -- so results are unchanged, only the cost of hopeless entries goes.
__device__ __forceinline__ float edge_min_sigma(float d_fixed, float lo, float hi, float h_fixed, float B, float h_free) {
    // minimise  h_fixed*d_fixed^2 + B*d_fixed*t + h_free*t^2  over t in [lo, hi]
    const float t_opt = (h_free > 0.0f) ? -0.5f * B * d_fixed / h_free : lo;
    const float t = fminf(hi, fmaxf(lo, t_opt));
    return fmaf(t, fmaf(h_free, t, B * d_fixed), h_fixed * d_fixed * d_fixed);
}

template <int BATCH>
__device__ __forceinline__ void cull_batch(const float* __restrict__ geo_buf, int cnt, float rx0, float ry0,
                                           float rx1, float ry1, unsigned (&mask)[BATCH / 32]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < BATCH / 32; ++k) {
        const int e = k * 32 + lane;
        bool pass = false;
        if (e < cnt) {
            const float4 ga = reinterpret_cast<const float4*>(geo_buf)[2 * e];
            const float4 gb = reinterpret_cast<const float4*>(geo_buf)[2 * e + 1];
            const float hA = ga.z, B = ga.w, hC = gb.x, tau = gb.z;
            // rectangle relative to the Gaussian centre (d = pixel - centre; sigma is even in d)
            const float x0 = rx0 - ga.x, x1 = rx1 - ga.x, y0 = ry0 - ga.y, y1 = ry1 - ga.y;
            float smin = 0.0f;
            if (!(x0 <= 0.0f && x1 >= 0.0f && y0 <= 0.0f && y1 >= 0.0f)) {
                smin = fminf(fminf(edge_min_sigma(x0, y0, y1, hA, B, hC), edge_min_sigma(x1, y0, y1, hA, B, hC)),
                             fminf(edge_min_sigma(y0, x0, x1, hC, B, hA), edge_min_sigma(y1, x0, x1, hC, B, hA)));
            }
            // 1e-4 relative slack covers the rounding difference to the per-pixel evaluation
            pass = (tau >= 0.0f) && (smin <= tau + 1e-4f * (1.0f + tau)) ;
            // conics that are not safely positive definite (or NaN) are never culled
            if (!(hA > 0.0f && hC > 0.0f && 4.0f * hA * hC - B * B > 0.0f) || !(smin == smin)) pass = tau >= 0.0f;
        }
        mask[k] = __ballot_sync(0xffffffffu, pass);
    }
}

template <int CP, int BATCH, bool kVec>
__global__ void __launch_bounds__(kBlendThreads, (CP <= 24) ? 4 : ((CP <= 48) ? 2 : 1))
blend_fwd_kernel(const BlendArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* geo_sm = smem;                            // [kStages][BATCH][8]
    float* col_sm = smem + kStages * BATCH * 8;      // [kStages][BATCH][CP]
    // tensor-core accumulation (GG_FWD_MMA): out[32 pixels x CP] += vis[32 x 8 entries] . colour[8 x CP]
    constexpr bool kMma = GG_FWD_MMA && (CP % 8 == 0);
    // per warp [8 entry slots][32 pixels], pixel index XOR-swizzled by 8 * (slot & 3): the A-fragment loads (4 slots
    // x 8 pixels per instruction) then hit 32 different banks without padding
    constexpr int kVisStride = 32;
    float* vis_sm = smem + kStages * BATCH * (8 + CP) + (threadIdx.x >> 5) * 8 * kVisStride;
    const int n_tiles = a.tiles_x * a.tiles_y;
    const int lin = blockIdx.y * n_tiles + blockIdx.x;
    const int gtile = a.tile_order ? __ldg(a.tile_order + lin) : lin;  // longest lists first
    const int view = gtile / n_tiles;
    const int tile = gtile - view * n_tiles;
    const int tile_y = tile / a.tiles_x, tile_x = tile - tile_y * a.tiles_x;
    int tx, ty;
    tile_pixel(tx, ty);
    const int warp = threadIdx.x >> 5;
    const int px = tile_x * GG_TILE + tx, py = tile_y * GG_TILE + ty;
    const bool inside = px < a.img_w && py < a.img_h;
    const float fpx = (float)px, fpy = (float)py;
    const float rx0 = (float)(tile_x * GG_TILE + ((warp & 1) << 3)), ry0 = (float)(tile_y * GG_TILE + ((warp >> 1) << 2));
    const float rx1 = rx0 + 7.0f, ry1 = ry0 + 3.0f;
    const int2 range = __ldg(reinterpret_cast<const int2*>(a.tile_ranges) + (long long)view * a.tiles_x * a.tiles_y + tile);
    const long long geo_base = (long long)view * a.geo_view_stride;
    const long long color_base = (long long)view * a.color_view_stride;

    // FP32 path: acc[c] of this thread's pixel.  MMA path: the same CP registers hold the warp's D fragments
    // d[mt][nt][4]: pixels 16 mt + lane/4 (+8), channels 8 nt + 2 (lane%4) (+1).
    float acc[CP];
#pragma unroll
    for (int c = 0; c < CP; ++c) acc[c] = 0.0f;
    int n_slots = 0;   // MMA path: entries waiting in vis_sm (warp-uniform)
    int slot_e = 0;    // ... lane l keeps the batch-local index of the entry in slot l & 7
    float T = 1.0f;
    int last = range.x;
    int stop = range.y;  // one past the last entry this pixel looked at
    bool done = !inside;
    bool warp_done = __all_sync(0xffffffffu, done);

    const int total = range.y - range.x;
    const int nb = (total + BATCH - 1) / BATCH;
    // kStages-deep ring of staging buffers: batch b+2 is issued while batch b is blended, and one
    // barrier per batch both publishes batch b and retires the buffer batch b+2 will overwrite
#if GG_FWD_RING
    __shared__ uint64_t full_bar[kStages], empty_bar[kStages];
    __shared__ int done_warps;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s_ = 0; s_ < kStages; ++s_) { mbar_init(&full_bar[s_], kBlendThreads); mbar_init(&empty_bar[s_], kBlendThreads / 32); }
        done_warps = 0;
    }
    __syncthreads();
    bool counted = false;
#endif
    auto stage = [&](int b) {
        const int first = range.x + b * BATCH;
        const int buf = b % kStages;
        stage_batch<CP, BATCH, kVec>(a, geo_base, color_base, first, min(BATCH, range.y - first),
                                     geo_sm + buf * BATCH * 8, col_sm + buf * BATCH * CP);
#if GG_FWD_RING
        cp_async_arrive_on(&full_bar[buf]);
#else
        cp_async_commit();
#endif
    };
    if (nb > 0) stage(0);
    if (nb > 1) stage(1);
    for (int b = 0; b < nb; ++b) {
        const int buf = b % kStages;
#if GG_FWD_RING
        mbar_wait(&full_bar[buf], (unsigned)((b / kStages) & 1));            // every thread's copies of batch b landed
        if (*reinterpret_cast<volatile int*>(&done_warps) == kBlendThreads / 32) break;
        if (b + 2 < nb) {
            // the buffer batch b+2 goes into was batch b-1's: every warp must have left it
            if (b >= 1) mbar_wait(&empty_bar[(b - 1) % kStages], (unsigned)(((b - 1) / kStages) & 1));
            stage(b + 2);
        }
#else
        if (b + 1 < nb) cp_async_wait<1>(); else cp_async_wait<0>();
        if (__syncthreads_count(warp_done) == kBlendThreads) break;
        if (b + 2 < nb) stage(b + 2);
#endif
        const int first = range.x + b * BATCH;
        const int cnt = min(BATCH, range.y - first);
        if (!warp_done) {
            const float* gbuf = geo_sm + buf * BATCH * 8;
            const float4* g4 = reinterpret_cast<const float4*>(gbuf);
            const float4* c4 = reinterpret_cast<const float4*>(col_sm + buf * BATCH * CP);
            unsigned mask[BATCH / 32];
            cull_batch<BATCH>(gbuf, cnt, rx0, ry0, rx1, ry1, mask);
            unsigned hits[BATCH / 32];
#pragma unroll
            for (int k = 0; k < BATCH / 32; ++k) hits[k] = 0u;
            const float* cbuf = col_sm + buf * BATCH * CP;
            const int lane_f = threadIdx.x & 31;
            // MMA path: the waiting slots' [32 x n] weights times their colour rows, 3xTF32 split, fp32 accumulate
            auto mma_flush = [&]() {
                if constexpr (kMma) {
                    __syncwarp();
                    const int g4 = lane_f >> 2, t4 = lane_f & 3;
                    const int e_lo = __shfl_sync(0xffffffffu, slot_e, t4), e_hi = __shfl_sync(0xffffffffu, slot_e, t4 + 4);
                    const bool k_lo = t4 < n_slots, k_hi = t4 + 4 < n_slots;     // empty slots hold stale weights
                    uint32_t ah[2][4], al[2][4];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        float af[4];
                        const int sw = 8 * t4;   // == 8 * ((t4 + 4) & 3)
                        af[0] = k_lo ? vis_sm[t4 * kVisStride + ((16 * mt + g4) ^ sw)] : 0.0f;
                        af[1] = k_lo ? vis_sm[t4 * kVisStride + ((16 * mt + g4 + 8) ^ sw)] : 0.0f;
                        af[2] = k_hi ? vis_sm[(t4 + 4) * kVisStride + ((16 * mt + g4) ^ sw)] : 0.0f;
                        af[3] = k_hi ? vis_sm[(t4 + 4) * kVisStride + ((16 * mt + g4 + 8) ^ sw)] : 0.0f;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(ah[mt][q]) : "f"(af[q]));
                            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(al[mt][q]) : "f"(af[q] - __uint_as_float(ah[mt][q])));
                        }
                    }
#pragma unroll
                    for (int nt = 0; nt < CP / 8; ++nt) {
                        const float b0f = k_lo ? cbuf[e_lo * CP + 8 * nt + g4] : 0.0f;
                        const float b1f = k_hi ? cbuf[e_hi * CP + 8 * nt + g4] : 0.0f;
                        uint32_t bh0, bh1, bl0, bl1;
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(bh0) : "f"(b0f));
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(bh1) : "f"(b1f));
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(bl0) : "f"(b0f - __uint_as_float(bh0)));
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(bl1) : "f"(b1f - __uint_as_float(bh1)));
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt) {
                            float* d = acc + (mt * (CP / 8) + nt) * 4;
#define GG_MMA_TF32(A, B0, B1)                                                                                        \
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, " \
                 "{%0, %1, %2, %3};"                                                                                  \
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])                                                     \
                 : "r"(A[0]), "r"(A[1]), "r"(A[2]), "r"(A[3]), "r"(B0), "r"(B1))
                            GG_MMA_TF32(al[mt], bh0, bh1);
                            GG_MMA_TF32(ah[mt], bl0, bl1);
                            GG_MMA_TF32(ah[mt], bh0, bh1);
#undef GG_MMA_TF32
                        }
                    }
                    n_slots = 0;
                    __syncwarp();
                }
            };
            auto blend = [&](int e, float alpha, float& vis_out) -> bool {
                const float next_T = T * (1.0f - alpha);
                if (next_T <= kTStop) { done = true; stop = first + e + 1; return false; }
                const float vis = alpha * T;
                if constexpr (kMma) {
                    vis_out = vis;
                } else {
#pragma unroll
                    for (int q = 0; q < CP / 4; ++q) {
                        const float4 cc = c4[e * (CP / 4) + q];
                        acc[4 * q] = fmaf(vis, cc.x, acc[4 * q]);
                        acc[4 * q + 1] = fmaf(vis, cc.y, acc[4 * q + 1]);
                        acc[4 * q + 2] = fmaf(vis, cc.z, acc[4 * q + 2]);
                        acc[4 * q + 3] = fmaf(vis, cc.w, acc[4 * q + 3]);
                    }
                }
                T = next_T;
                last = first + e + 1;
                return true;
            };
            // MMA path: an entry somebody contributed to takes the next slot
            auto enqueue = [&](int e, float vis) {
                if constexpr (kMma) {
                    vis_sm[n_slots * kVisStride + (lane_f ^ (8 * (n_slots & 3)))] = vis;
                    if ((lane_f & 7) == n_slots) slot_e = e;
                    if (++n_slots == 8) mma_flush();
                }
            };
#pragma unroll
            for (int k = 0; k < BATCH / 32; ++k) {
                unsigned m = mask[k];
                while (m) {
                    // two surviving entries per round: their alpha chains are independent, only the
                    // transmittance update is sequential
                    const int b0 = __ffs(m) - 1;
                    const int e0 = k * 32 + b0;
                    m &= m - 1;
                    const bool two = m != 0;
                    const int b1 = two ? __ffs(m) - 1 : b0;
                    const int e1 = k * 32 + b1;
                    m &= m - 1;
                    const float4 ga0 = g4[2 * e0], gb0 = g4[2 * e0 + 1];
                    const float4 ga1 = g4[2 * e1], gb1 = g4[2 * e1 + 1];
                    const float s0 = eval_sigma(ga0.x - fpx, ga0.y - fpy, ga0.z, ga0.w, gb0.x);
                    const float s1 = eval_sigma(ga1.x - fpx, ga1.y - fpy, ga1.z, ga1.w, gb1.x);
                    const float a0 = fminf(kAlphaMax, gb0.y * __expf(-s0));
                    const float a1 = fminf(kAlphaMax, gb1.y * __expf(-s1));
                    const bool ok0 = !(s0 < 0.0f || s0 > gb0.z) && a0 >= kAlphaMin;
                    const bool ok1 = two && !(s1 < 0.0f || s1 > gb1.z) && a1 >= kAlphaMin;
                    float v0 = 0.0f, v1 = 0.0f;
                    const bool c0 = (ok0 && !done) ? blend(e0, a0, v0) : false;
                    const bool c1 = (ok1 && !done) ? blend(e1, a1, v1) : false;
                    if (__any_sync(0xffffffffu, c0)) { hits[k] |= 1u << b0; enqueue(e0, v0); }
                    if (__any_sync(0xffffffffu, c1)) { hits[k] |= 1u << b1; enqueue(e1, v1); }
                }
                if (__all_sync(0xffffffffu, done)) { warp_done = true; break; }
            }
            if (n_slots) mma_flush();   // the staged colour rows of this batch are recycled two batches from now
            if (a.hit_words) {
                // record index: see BlendArgs::hit_words (a.fwd_batch == BATCH in the forward)
                const long long rec = (long long)(range.x / BATCH) + gtile + b;
                uint32_t* dst = a.hit_words + (rec * (kBlendThreads / 32) + warp) * (BATCH / 32);
                const int lane_id = threadIdx.x & 31;
#pragma unroll
                for (int k = 0; k < BATCH / 32; ++k)
                    if (lane_id == k) dst[k] = hits[k];
            }
        }
#if GG_FWD_RING
        __syncwarp();
        if ((threadIdx.x & 31) == 0) {
            if (warp_done && !counted) atomicAdd(&done_warps, 1);
            mbar_arrive_cta(&empty_bar[buf]);                                   // this warp is out of batch b's buffer
        }
        counted = counted || warp_done;
#endif
    }
#if GG_FWD_RING
    asm volatile("cp.async.wait_all;\n" ::: "memory");
#else
    cp_async_wait<0>();
#endif
    if constexpr (kMma) {
        // D fragments -> image: this lane holds pixels (16 mt + lane/4, +8) of the warp's 8x4 block, whose
        // transmittance sits in the lanes of those pixels
        const int lane_f = threadIdx.x & 31, g4 = lane_f >> 2, t4 = lane_f & 3;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int p = 16 * mt + g4 + 8 * hh;                         // pixel = lane index of its owner
                const float Tp = __shfl_sync(0xffffffffu, T, p);
                const int ppx = tile_x * GG_TILE + ((p & 7) | ((warp & 1) << 3));
                const int ppy = tile_y * GG_TILE + ((p >> 3) | ((warp >> 1) << 2));
                if (ppx < a.img_w && ppy < a.img_h) {
                    float* o = a.out + (((long long)view * a.img_h + ppy) * a.img_w + ppx) * a.out_stride;
#pragma unroll
                    for (int nt = 0; nt < CP / 8; ++nt) {
                        const int c0 = 8 * nt + 2 * t4;
                        const float* d = acc + (mt * (CP / 8) + nt) * 4 + 2 * hh;
                        if (c0 < a.channels) o[c0] = fmaf(Tp, __ldg(a.bg + c0), d[0]);
                        if (c0 + 1 < a.channels) o[c0 + 1] = fmaf(Tp, __ldg(a.bg + c0 + 1), d[1]);
                    }
                }
            }
        if (inside) {
            const long long pix = ((long long)view * a.img_h + py) * a.img_w + px;
            a.final_T[pix] = T;
            a.final_idx[pix] = last;
        }
    } else if (inside) {
        const long long pix = ((long long)view * a.img_h + py) * a.img_w + px;
        float* o = a.out + pix * a.out_stride;
#pragma unroll
        for (int c = 0; c < CP; ++c)
            if (c < a.channels) o[c] = fmaf(T, __ldg(a.bg + c), acc[c]);
        a.final_T[pix] = T;
        a.final_idx[pix] = last;
    }
    if (a.pair_counter) {
        // K of SURVEY 8d: entries between the tile's start and the last one this pixel looked at
        long long n_vis = inside ? (long long)(stop - range.x) : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n_vis += __shfl_xor_sync(0xffffffffu, n_vis, o);
        if ((threadIdx.x & 31) == 0 && n_vis) atomicAdd(a.pair_counter, (unsigned long long)n_vis);
    }
}

// ---------------------------------------------------------------------------------------------
// Backward.  Per warp (8x4 pixels), back to front over the culled entries of the tile:
//   phase A (lane = pixel): replay alpha, T; the alpha gradient needs from "behind" only the scalar
//     R = T_final*<bg,v> + sum_behind fac*<colour,v>; each entry with a contributor stores two
//     scalars per pixel in a warp-private shared-memory matrix: fac = alpha*T_front and
//     w = e^{-sigma} * dL/dalpha (0 when alpha is clamped);
//   phase B (lane = entry, every 32 stored entries): lane e walks the 32 pixels and accumulates
//     v_colour[c] = sum_p fac[p]*v_out[p][c] (v_out rows broadcast from shared memory) and the six
//     moments sum_p w*{1,dx,dy,dx^2,dxdy,dy^2} from which v_xy, v_conic, v_opacity follow; then one
//     red.global.add per component.
// No cross-lane shuffle reduction and no per-pixel atomics: the transposition through shared
// memory turns the per-Gaussian reduction into register accumulation at full lane occupancy.
// ---------------------------------------------------------------------------------------------
constexpr int kHitRow = 33;                                  // padded row: conflict-free both ways
constexpr int kHitRows = 16;                                 // entries stored before a flush
constexpr int kWarpScratch = 2 * kHitRows * kHitRow + kHitRows;  // facm, wm, hit ids (floats)

template <int CP, int BATCH>
constexpr size_t blend_bwd_smem() {
    return sizeof(float) * (kStages * BATCH * (8 + CP) + kStages * BATCH + kBlendThreads * CP +
                            (kBlendThreads / 32) * kWarpScratch);
}

template <int CP, int BATCH, bool kVec>
__global__ void __launch_bounds__(kBlendThreads)
blend_bwd_kernel(const BlendArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* geo_sm = smem;                                                  // [kStages][BATCH][8]
    float* col_sm = geo_sm + kStages * BATCH * 8;                          // [kStages][BATCH][CP]
    int* ids_sm = reinterpret_cast<int*>(col_sm + kStages * BATCH * CP);   // [kStages][BATCH]
    float* vo_sm = reinterpret_cast<float*>(ids_sm + kStages * BATCH);     // [256][CP]
    __shared__ int s_max[kBlendThreads / 32];
    const int n_tiles = a.tiles_x * a.tiles_y;
    const int lin = blockIdx.y * n_tiles + blockIdx.x;
    const int gtile = a.tile_order ? __ldg(a.tile_order + lin) : lin;  // longest lists first
    const int view = gtile / n_tiles;
    const int tile = gtile - view * n_tiles;
    const int tile_y = tile / a.tiles_x, tile_x = tile - tile_y * a.tiles_x;
    int tx, ty;
    tile_pixel(tx, ty);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* facm = vo_sm + kBlendThreads * CP + warp * kWarpScratch;        // [kHitRows][33]
    float* wm = facm + kHitRows * kHitRow;                                 // [kHitRows][33]
    int* hit_g = reinterpret_cast<int*>(wm + kHitRows * kHitRow);          // [kHitRows]
    const int px = tile_x * GG_TILE + tx, py = tile_y * GG_TILE + ty;
    const bool inside = px < a.img_w && py < a.img_h;
    const float fpx = (float)px, fpy = (float)py;
    const float rx0 = (float)(tile_x * GG_TILE + ((warp & 1) << 3)), ry0 = (float)(tile_y * GG_TILE + ((warp >> 1) << 2));
    const float rx1 = rx0 + 7.0f, ry1 = ry0 + 3.0f;
    const int2 range = __ldg(reinterpret_cast<const int2*>(a.tile_ranges) + (long long)view * a.tiles_x * a.tiles_y + tile);
    const long long geo_base = (long long)view * a.geo_view_stride;
    const long long color_base = (long long)view * a.color_view_stride;
    const long long pix = ((long long)view * a.img_h + py) * a.img_w + px;

    // padding lanes of the staged colour rows are read by the dot product: keep them finite
    for (int k = threadIdx.x; k < kStages * BATCH * CP; k += kBlendThreads) col_sm[k] = 0.0f;

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
    {
        float4* row = reinterpret_cast<float4*>(vo_sm + threadIdx.x * CP);
#pragma unroll
        for (int q = 0; q < CP / 4; ++q) row[q] = make_float4(vo[4 * q], vo[4 * q + 1], vo[4 * q + 2], vo[4 * q + 3]);
    }
    float T = T_final;
    float R = T_final * bgdot;
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

    const float* vo_warp = vo_sm + warp * 32 * CP;
    int nhit = 0;
    auto reload_vo = [&]() {
        const float4* row = reinterpret_cast<const float4*>(vo_sm + threadIdx.x * CP);
#pragma unroll
        for (int q = 0; q < CP / 4; ++q) {
            const float4 v = row[q];
            vo[4 * q] = v.x; vo[4 * q + 1] = v.y; vo[4 * q + 2] = v.z; vo[4 * q + 3] = v.w;
        }
    };
    // phase B: lane = (stored entry j = lane & 15, pixel half h = lane >> 4); each half-warp sums
    // its 16 pixels, the halves are combined with one shuffle per value, then the 6 + C reds of an
    // entry are split between its two lanes
    auto flush = [&](int count) {
        __syncwarp();
        const int j = lane & (kHitRows - 1), h = lane >> 4;
        const bool live = j < count;
        const int g = live ? hit_g[j] : 0;
        float4 ga = make_float4(0.f, 0.f, 0.f, 0.f), gb = ga;
        if (live) {
            ga = __ldg(reinterpret_cast<const float4*>(a.geo) + 2 * (geo_base + g));
            gb = __ldg(reinterpret_cast<const float4*>(a.geo) + 2 * (geo_base + g) + 1);
        }
        float acc[CP];
#pragma unroll
        for (int c = 0; c < CP; ++c) acc[c] = 0.0f;
        float m0 = 0.f, mx = 0.f, my = 0.f, mxx = 0.f, mxy = 0.f, myy = 0.f;
        const float* fr = facm + j * kHitRow + 16 * h;
        const float* wr = wm + j * kHitRow + 16 * h;
        const float bx = ga.x - rx0, by = ga.y - ry0 - (float)(2 * h);
#pragma unroll kBwdBUnroll
        for (int p = 0; p < 16; ++p) {
            const float f = live ? fr[p] : 0.0f;
            const float wv = live ? wr[p] : 0.0f;
            const float4* vrow = reinterpret_cast<const float4*>(vo_warp + (16 * h + p) * CP);
#pragma unroll
            for (int q = 0; q < CP / 4; ++q) {
                const float4 v = vrow[q];
                acc[4 * q] = fmaf(f, v.x, acc[4 * q]);
                acc[4 * q + 1] = fmaf(f, v.y, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(f, v.z, acc[4 * q + 2]);
                acc[4 * q + 3] = fmaf(f, v.w, acc[4 * q + 3]);
            }
            const float dx = bx - (float)(p & 7), dy = by - (float)(p >> 3);
            const float wdx = wv * dx, wdy = wv * dy;
            m0 += wv; mx += wdx; my += wdy;
            mxx = fmaf(wdx, dx, mxx); mxy = fmaf(wdx, dy, mxy); myy = fmaf(wdy, dy, myy);
        }
#pragma unroll
        for (int c = 0; c < CP; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 16);
        m0 += __shfl_xor_sync(0xffffffffu, m0, 16);
        mx += __shfl_xor_sync(0xffffffffu, mx, 16);
        my += __shfl_xor_sync(0xffffffffu, my, 16);
        mxx += __shfl_xor_sync(0xffffffffu, mxx, 16);
        mxy += __shfl_xor_sync(0xffffffffu, mxy, 16);
        myy += __shfl_xor_sync(0xffffffffu, myy, 16);
        if (live) {
            float* vc = a.v_colors + (color_base + g) * (long long)a.color_stride;
            if (h == 0) {
                const float o = gb.y, A = 2.0f * ga.z, B = ga.w, C = 2.0f * gb.x;
                float* vg = a.v_geo + (geo_base + g) * 8;
                atomicAdd(vg + 0, -o * fmaf(A, mx, B * my));
                atomicAdd(vg + 1, -o * fmaf(B, mx, C * my));
                atomicAdd(vg + 2, -0.5f * o * mxx);
                atomicAdd(vg + 3, -o * mxy);
                atomicAdd(vg + 4, -0.5f * o * myy);
                atomicAdd(vg + 5, m0);
            }
            // lane h==0 takes the low channels (fewer, it also did the geometry), h==1 the rest
            constexpr int kSplit = (CP > 6) ? (CP - 6) / 2 : 0;
#pragma unroll
            for (int c = 0; c < CP; ++c)
                if (c < a.channels && ((c < kSplit) == (h == 0))) atomicAdd(vc + c, acc[c]);
        }
        __syncwarp();
    };

    auto stage = [&](int b) {
        const int first = range.x + b * BATCH;
        const int cnt = min(BATCH, bmax - first);
        const int buf = b % kStages;
        if (threadIdx.x < cnt) ids_sm[buf * BATCH + threadIdx.x] = __ldg(a.ids_sorted + first + threadIdx.x);
        stage_batch<CP, BATCH, kVec>(a, geo_base, color_base, first, cnt, geo_sm + buf * BATCH * 8,
                                     col_sm + buf * BATCH * CP);
        cp_async_commit();
    };
    stage(nb - 1);
    if (nb > 1) stage(nb - 2);
    for (int b = nb - 1; b >= 0; --b) {
        const int buf = b % kStages;
        if (b > 0) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncthreads();  // batch b is visible; everyone is done with the buffer batch b-2 reuses
        if (b > 1) stage(b - 2);
        const int first = range.x + b * BATCH;
        const int cnt = min(BATCH, bmax - first);
        // entries at or beyond the warp's own last contributor are dead for the whole warp
        const int e_hi = min(cnt, wmax - first);
        if (e_hi > 0) {
            const float* gbuf = geo_sm + buf * BATCH * 8;
            const float4* g4 = reinterpret_cast<const float4*>(gbuf);
            const float4* c4 = reinterpret_cast<const float4*>(col_sm + buf * BATCH * CP);
            const int* idb = ids_sm + buf * BATCH;
            unsigned mask[BATCH / 32];
            if (a.hit_words) {
                // entries the forward recorded as contributing for this warp (bwd batches are 64 wide,
                // forward records fwd_batch wide: pick the matching words)
                const int fb = a.fwd_batch;
                const int e_abs = b * BATCH;  // offset of this batch inside the tile's list
                const long long rec = (long long)(range.x / fb) + gtile + e_abs / fb;
                const uint32_t* src = a.hit_words + (rec * (kBlendThreads / 32) + warp) * (fb / 32) + (e_abs % fb) / 32;
#pragma unroll
                for (int k = 0; k < BATCH / 32; ++k) {
                    const int lo = k * 32;
                    unsigned w = lo < e_hi ? __ldg(src + k) : 0u;
                    if (e_hi - lo < 32 && e_hi > lo) w &= (1u << (e_hi - lo)) - 1u;
                    mask[k] = w;
                }
            } else {
                cull_batch<BATCH>(gbuf, e_hi, rx0, ry0, rx1, ry1, mask);
            }
            auto colour_dot = [&](int e) {
                float d0 = 0.0f, d1 = 0.0f;
#pragma unroll
                for (int q = 0; q < CP / 4; ++q) {
                    const float4 cc = c4[e * (CP / 4) + q];
                    d0 = fmaf(cc.x, vo[4 * q], d0);
                    d1 = fmaf(cc.y, vo[4 * q + 1], d1);
                    d0 = fmaf(cc.z, vo[4 * q + 2], d0);
                    d1 = fmaf(cc.w, vo[4 * q + 3], d1);
                }
                return d0 + d1;
            };
            // sequential part of one entry: transmittance, alpha gradient, row of the hit matrices
            auto record = [&](int e, bool valid, float alpha, float araw, float vis, float dot) {
                float fac = 0.0f, w = 0.0f;
                if (valid) {
                    const float ra = 1.0f / (1.0f - alpha);
                    T *= ra;  // transmittance in front of this entry
                    fac = alpha * T;
                    const float v_alpha = fmaf(dot, T, -R * ra);
                    R = fmaf(fac, dot, R);
                    // a clamped alpha passes no gradient to sigma / opacity
                    w = (araw <= kAlphaMax) ? vis * v_alpha : 0.0f;
                }
                facm[nhit * kHitRow + lane] = fac;
                wm[nhit * kHitRow + lane] = w;
                if (lane == 0) hit_g[nhit] = idb[e];
                if (++nhit == kHitRows) {
                    flush(kHitRows);
                    nhit = 0;
                    reload_vo();  // vo[] was dead across the flush: its registers held the accumulators
                }
            };
#pragma unroll
            for (int k = BATCH / 32 - 1; k >= 0; --k) {
                unsigned m = mask[k];
                while (m) {
                    // two surviving entries per round (independent alpha / dot chains)
                    const int bit0 = 31 - __clz(m);
                    m &= ~(1u << bit0);
                    const bool two = m != 0;
                    const int bit1 = two ? 31 - __clz(m) : bit0;
                    m &= ~(1u << bit1);
                    const int e0 = k * 32 + bit0, e1 = k * 32 + bit1;
                    const float4 ga0 = g4[2 * e0], gb0 = g4[2 * e0 + 1];
                    const float4 ga1 = g4[2 * e1], gb1 = g4[2 * e1 + 1];
                    // branch-free replay of the forward test (same arithmetic as blend_fwd_kernel)
                    const float s0 = eval_sigma(ga0.x - fpx, ga0.y - fpy, ga0.z, ga0.w, gb0.x);
                    const float s1 = eval_sigma(ga1.x - fpx, ga1.y - fpy, ga1.z, ga1.w, gb1.x);
                    const float vis0 = __expf(-s0), vis1 = __expf(-s1);
                    const float araw0 = gb0.y * vis0, araw1 = gb1.y * vis1;
                    const float al0 = fminf(kAlphaMax, araw0), al1 = fminf(kAlphaMax, araw1);
                    const bool v0 = (first + e0 < last) && !(s0 < 0.0f || s0 > gb0.z) && (al0 >= kAlphaMin);
                    const bool v1 = two && (first + e1 < last) && !(s1 < 0.0f || s1 > gb1.z) && (al1 >= kAlphaMin);
                    const bool any0 = __any_sync(0xffffffffu, v0), any1 = __any_sync(0xffffffffu, v1);
                    if (!(any0 || any1)) continue;
                    const float dot0 = v0 ? colour_dot(e0) : 0.0f;
                    const float dot1 = v1 ? colour_dot(e1) : 0.0f;
                    if (any0) record(e0, v0, al0, araw0, vis0, dot0);
                    if (any1) record(e1, v1, al1, araw1, vis1, dot1);
                }
            }
        }
    }
    if (nhit) flush(nhit);
}

// ---------------------------------------------------------------------------------------------
// Warp-autonomous backward (used whenever the forward's hit masks are available).
// Every warp owns one 8x4 pixel block and walks, back to front, ONLY the entries the forward
// recorded as contributing to that block.  It gathers their geo/channel rows itself (cp.async
// into a warp-private double buffer), so there is no CTA-wide staging, no __syncthreads and no
// coupling between the eight blocks of a tile: a warp's time is the sum of its own work instead
// of, per staged batch, the maximum over the tile's warps.  CTAs are 4 warps = half a tile.
// ---------------------------------------------------------------------------------------------
constexpr int kBwdWarps = 4;
constexpr int kChunk = 16;  // entries gathered per round (= kHitRows)

template <int CP>
constexpr size_t blend_bwd_warp_smem() {
    return sizeof(float) * kBwdWarps * (2 * kChunk * (8 + CP) + 2 * kHitRows * kHitRow + 32 * CP);
}

template <int CP, bool kVec>
__global__ void __launch_bounds__(kBwdWarps * 32, (CP <= 24) ? GG_BWD_MIN_BLOCKS : 1)
blend_bwd_warp_kernel(const BlendArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    constexpr int kRow = 8 + CP;                                   // staged row: geo then channels
    constexpr int kPerWarp = 2 * kChunk * kRow + 2 * kHitRows * kHitRow + 32 * CP;
    float* wsm = smem + wib * kPerWarp;
    float* rows = wsm;                                             // [2][kChunk][kRow]
    float* facm = rows + 2 * kChunk * kRow;                        // [kHitRows][33]
    float* wm = facm + kHitRows * kHitRow;                         // [kHitRows][33]
    float* vo_warp = wm + kHitRows * kHitRow;                      // [32][CP]

    const int n_tiles = a.tiles_x * a.tiles_y;
    const int lin = blockIdx.y * n_tiles + (blockIdx.x >> 1);
    const int gtile = a.tile_order ? __ldg(a.tile_order + lin) : lin;
    const int view = gtile / n_tiles;
    const int tile = gtile - view * n_tiles;
    const int tile_y = tile / a.tiles_x, tile_x = tile - tile_y * a.tiles_x;
    const int warp = (blockIdx.x & 1) * kBwdWarps + wib;           // 0..7: which 8x4 block of the tile
    const int tx = (lane & 7) | ((warp & 1) << 3), ty = (lane >> 3) | ((warp >> 1) << 2);
    const int px = tile_x * GG_TILE + tx, py = tile_y * GG_TILE + ty;
    const bool inside = px < a.img_w && py < a.img_h;
    const float fpx = (float)px, fpy = (float)py;
    const float rx0 = (float)(tile_x * GG_TILE + ((warp & 1) << 3)), ry0 = (float)(tile_y * GG_TILE + ((warp >> 1) << 2));
    const int2 range = __ldg(reinterpret_cast<const int2*>(a.tile_ranges) + gtile);
    const long long geo_base = (long long)view * a.geo_view_stride;
    const long long color_base = (long long)view * a.color_view_stride;
    const long long pix = ((long long)view * a.img_h + py) * a.img_w + px;

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
    {
        float4* row = reinterpret_cast<float4*>(vo_warp + lane * CP);
#pragma unroll
        for (int q = 0; q < CP / 4; ++q) row[q] = make_float4(vo[4 * q], vo[4 * q + 1], vo[4 * q + 2], vo[4 * q + 3]);
    }
    // zero the staged rows once: padding channels are read by the dot product
    for (int k = lane; k < 2 * kChunk * kRow; k += 32) rows[k] = 0.0f;
    float T = T_final;
    float R = T_final * bgdot;
    int wmax = last;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    __syncwarp();
    const int e_top = wmax - range.x;  // entries [0, e_top) of the tile's list can matter to this warp
    if (e_top <= 0) return;

    int nhit = 0;
    int fg = 0;  // Gaussian id of the stored entry (lane & 15) waiting in the hit matrices
    auto reload_vo = [&]() {
        const float4* row = reinterpret_cast<const float4*>(vo_warp + lane * CP);
#pragma unroll
        for (int q = 0; q < CP / 4; ++q) {
            const float4 v = row[q];
            vo[4 * q] = v.x; vo[4 * q + 1] = v.y; vo[4 * q + 2] = v.z; vo[4 * q + 3] = v.w;
        }
    };
    // phase B: lane = (stored entry j = lane & 15, pixel half h = lane >> 4), see blend_bwd_kernel
    auto flush = [&](int count) {
        __syncwarp();
        const int j = lane & (kHitRows - 1), h = lane >> 4;
        const bool live = j < count;
        const int g = live ? fg : 0;
        float4 ga = make_float4(0.f, 0.f, 0.f, 0.f), gb = ga;
        if (live) {
            ga = __ldg(reinterpret_cast<const float4*>(a.geo) + 2 * (geo_base + g));
            gb = __ldg(reinterpret_cast<const float4*>(a.geo) + 2 * (geo_base + g) + 1);
        }
        float m0 = 0.f, mx = 0.f, my = 0.f, mxx = 0.f, mxy = 0.f, myy = 0.f;
        const float* fr = facm + j * kHitRow + 16 * h;
        const float* wr = wm + j * kHitRow + 16 * h;
        const float bx = ga.x - rx0, by = ga.y - ry0 - (float)(2 * h);
        constexpr bool kMma = GG_BWD_MMA && (CP % 8 == 0);
        if constexpr (kMma) {
            // v_colour[16 entries x CP] = fac[16 x 32 pixels] . v_out[32 pixels x CP] on the tensor cores:
            // mma.sync m16n8k8 TF32 with the 3xTF32 split (hi.hi + lo.hi + hi.lo, fp32 accumulation): fp32-equivalent.
            const int g4 = lane >> 2, t4 = lane & 3;
            float d[CP / 8][4];
#pragma unroll
            for (int nt = 0; nt < CP / 8; ++nt) d[nt][0] = d[nt][1] = d[nt][2] = d[nt][3] = 0.0f;
            const bool row_lo = g4 < count, row_hi = g4 + 8 < count;   // rows beyond `count` hold stale values
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const int k0 = 8 * ks + t4;
                float af[4];
                af[0] = row_lo ? facm[g4 * kHitRow + k0] : 0.0f;
                af[1] = row_hi ? facm[(g4 + 8) * kHitRow + k0] : 0.0f;
                af[2] = row_lo ? facm[g4 * kHitRow + k0 + 4] : 0.0f;
                af[3] = row_hi ? facm[(g4 + 8) * kHitRow + k0 + 4] : 0.0f;
                uint32_t ah[4], al[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(ah[q]) : "f"(af[q]));
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(al[q]) : "f"(af[q] - __uint_as_float(ah[q])));
                }
#pragma unroll
                for (int nt = 0; nt < CP / 8; ++nt) {
                    const float b0f = vo_warp[k0 * CP + 8 * nt + g4], b1f = vo_warp[(k0 + 4) * CP + 8 * nt + g4];
                    uint32_t bh0, bh1, bl0, bl1;
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(bh0) : "f"(b0f));
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(bh1) : "f"(b1f));
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(bl0) : "f"(b0f - __uint_as_float(bh0)));
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(bl1) : "f"(b1f - __uint_as_float(bh1)));
#define GG_MMA_TF32(A, B0, B1)                                                                                        \
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, " \
                 "{%0, %1, %2, %3};"                                                                                  \
                 : "+f"(d[nt][0]), "+f"(d[nt][1]), "+f"(d[nt][2]), "+f"(d[nt][3])                                     \
                 : "r"(A[0]), "r"(A[1]), "r"(A[2]), "r"(A[3]), "r"(B0), "r"(B1))
                    GG_MMA_TF32(al, bh0, bh1);
                    GG_MMA_TF32(ah, bl0, bl1);
                    GG_MMA_TF32(ah, bh0, bh1);
#undef GG_MMA_TF32
                }
            }
            // lane holds entries g4 / g4 + 8, channels 8 nt + 2 t4 (+1)
            const int gid_lo = __shfl_sync(0xffffffffu, fg, g4), gid_hi = __shfl_sync(0xffffffffu, fg, g4 + 8);
            float* vc_lo = a.v_colors + (color_base + gid_lo) * (long long)a.color_stride;
            float* vc_hi = a.v_colors + (color_base + gid_hi) * (long long)a.color_stride;
#pragma unroll
            for (int nt = 0; nt < CP / 8; ++nt) {
                const int c0 = 8 * nt + 2 * t4;
                if (row_lo) {
                    if (c0 < a.channels) atomicAdd(vc_lo + c0, d[nt][0]);
                    if (c0 + 1 < a.channels) atomicAdd(vc_lo + c0 + 1, d[nt][1]);
                }
                if (row_hi) {
                    if (c0 < a.channels) atomicAdd(vc_hi + c0, d[nt][2]);
                    if (c0 + 1 < a.channels) atomicAdd(vc_hi + c0 + 1, d[nt][3]);
                }
            }
            // the six moments stay on the FP32 pipe (lane = entry j, pixel half h)
#pragma unroll
            for (int p = 0; p < 16; ++p) {
                const float wv = live ? wr[p] : 0.0f;
                const float dx = bx - (float)(p & 7), dy = by - (float)(p >> 3);
                const float wdx = wv * dx, wdy = wv * dy;
                m0 += wv; mx += wdx; my += wdy;
                mxx = fmaf(wdx, dx, mxx); mxy = fmaf(wdx, dy, mxy); myy = fmaf(wdy, dy, myy);
            }
            m0 += __shfl_xor_sync(0xffffffffu, m0, 16);
            mx += __shfl_xor_sync(0xffffffffu, mx, 16);
            my += __shfl_xor_sync(0xffffffffu, my, 16);
            mxx += __shfl_xor_sync(0xffffffffu, mxx, 16);
            mxy += __shfl_xor_sync(0xffffffffu, mxy, 16);
            myy += __shfl_xor_sync(0xffffffffu, myy, 16);
            if (live && h == 0) {
                const float o = gb.y, A = 2.0f * ga.z, B = ga.w, C = 2.0f * gb.x;
                float* vg = a.v_geo + (geo_base + g) * 8;
                atomicAdd(vg + 0, -o * fmaf(A, mx, B * my));
                atomicAdd(vg + 1, -o * fmaf(B, mx, C * my));
                atomicAdd(vg + 2, -0.5f * o * mxx);
                atomicAdd(vg + 3, -o * mxy);
                atomicAdd(vg + 4, -0.5f * o * myy);
                atomicAdd(vg + 5, m0);
            }
        } else {
        float acc[CP];
#pragma unroll
        for (int c = 0; c < CP; ++c) acc[c] = 0.0f;
#pragma unroll kBwdBUnroll
        for (int p = 0; p < 16; ++p) {
            const float f = live ? fr[p] : 0.0f;
            const float wv = live ? wr[p] : 0.0f;
            const float4* vrow = reinterpret_cast<const float4*>(vo_warp + (16 * h + p) * CP);
#pragma unroll
            for (int q = 0; q < CP / 4; ++q) {
                const float4 v = vrow[q];
                acc[4 * q] = fmaf(f, v.x, acc[4 * q]);
                acc[4 * q + 1] = fmaf(f, v.y, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(f, v.z, acc[4 * q + 2]);
                acc[4 * q + 3] = fmaf(f, v.w, acc[4 * q + 3]);
            }
            const float dx = bx - (float)(p & 7), dy = by - (float)(p >> 3);
            const float wdx = wv * dx, wdy = wv * dy;
            m0 += wv; mx += wdx; my += wdy;
            mxx = fmaf(wdx, dx, mxx); mxy = fmaf(wdx, dy, mxy); myy = fmaf(wdy, dy, myy);
        }
#pragma unroll
        for (int c = 0; c < CP; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 16);
        m0 += __shfl_xor_sync(0xffffffffu, m0, 16);
        mx += __shfl_xor_sync(0xffffffffu, mx, 16);
        my += __shfl_xor_sync(0xffffffffu, my, 16);
        mxx += __shfl_xor_sync(0xffffffffu, mxx, 16);
        mxy += __shfl_xor_sync(0xffffffffu, mxy, 16);
        myy += __shfl_xor_sync(0xffffffffu, myy, 16);
        if (live) {
            float* vc = a.v_colors + (color_base + g) * (long long)a.color_stride;
            if (h == 0) {
                const float o = gb.y, A = 2.0f * ga.z, B = ga.w, C = 2.0f * gb.x;
                float* vg = a.v_geo + (geo_base + g) * 8;
                atomicAdd(vg + 0, -o * fmaf(A, mx, B * my));
                atomicAdd(vg + 1, -o * fmaf(B, mx, C * my));
                atomicAdd(vg + 2, -0.5f * o * mxx);
                atomicAdd(vg + 3, -o * mxy);
                atomicAdd(vg + 4, -0.5f * o * myy);
                atomicAdd(vg + 5, m0);
            }
            constexpr int kSplit = (CP > 6) ? (CP - 6) / 2 : 0;
#pragma unroll
            for (int c = 0; c < CP; ++c)
                if (c < a.channels && ((c < kSplit) == (h == 0))) atomicAdd(vc + c, acc[c]);
        }
        }
        __syncwarp();
    };

    const int fb = a.fwd_batch;
    const long long rec0 = (long long)(range.x / fb) + gtile;
    auto hit_word = [&](int w) -> unsigned {
        // word w covers entries [32 w, 32 w + 32) of the tile's list
        const int e0 = w * 32;
        const long long rec = rec0 + e0 / fb;
        unsigned v = __ldg(a.hit_words + (rec * (kBlendThreads / 32) + warp) * (fb / 32) + (e0 % fb) / 32);
        if (e_top - e0 < 32) v &= (1u << (e_top - e0)) - 1u;
        return v;
    };
    // hit-entry iterator (warp-uniform): highest entry first, the next mask word prefetched
    int wi = (e_top - 1) >> 5;
    unsigned cur = hit_word(wi);
    unsigned nxt = wi > 0 ? hit_word(wi - 1) : 0u;
    auto take = [&](int& e) -> bool {
        while (cur == 0u) {
            if (wi == 0) return false;
            --wi;
            cur = nxt;
            nxt = wi > 0 ? hit_word(wi - 1) : 0u;
        }
        const int bit = 31 - __clz(cur);
        cur &= ~(1u << bit);
        e = wi * 32 + bit;
        return true;
    };
    // up to kChunk entries (they may span several mask words); lanes j and j+16 keep entry j
    auto build_chunk = [&](int& my_e) -> int {
        int k = 0, e = 0;
        while (k < kChunk && take(e)) {
            if ((lane & (kChunk - 1)) == k) my_e = e;
            ++k;
        }
        return k;
    };
    // gather the rows of a chunk: lane l serves entry j = l & 15 and copies half of its row
    auto gather = [&](int k, int my_e, int& my_g, int buf) {
        const int j = lane & (kChunk - 1), half = lane >> 4;
        if (j < k) {
            const int g = __ldg(a.ids_sorted + range.x + my_e);
            my_g = g;
            const float* grow = a.geo + (geo_base + g) * 8;
            const float* crow = a.colors + (color_base + g) * (long long)a.color_stride;
            float* dst = rows + (buf * kChunk + j) * kRow;
            if (kVec) {
                constexpr int kChunks16 = 2 + CP / 4;
                const int cchunks = 2 + ((a.channels + 3) >> 2);
                for (int q = half; q < kChunks16; q += 2) {
                    if (q < 2) cp_async16(dst + 4 * q, grow + 4 * q);
                    else if (q < cchunks) cp_async16(dst + 4 * q, crow + 4 * (q - 2));
                }
            } else {
                if (half == 0) { cp_async16(dst, grow); cp_async16(dst + 4, grow + 4); }
                for (int c = half; c < a.channels; c += 2) cp_async4(dst + 8 + c, crow + c);
            }
        }
        cp_async_commit();
    };

    int cur_e = 0, nxt_e = 0, cur_g = 0, nxt_g = 0;
    int cur_k = build_chunk(cur_e);
    int buf = 0;
    if (cur_k) gather(cur_k, cur_e, cur_g, buf);
    while (cur_k) {
        const int nxt_k = build_chunk(nxt_e);
        if (nxt_k) { gather(nxt_k, nxt_e, nxt_g, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncwarp();
        const float* rp = rows + buf * kChunk * kRow;  // row of the current entry
        const int rel_last = last - range.x;            // this pixel contributed to entries < rel_last
        for (int j = 0; j < cur_k; ++j, rp += kRow) {
            const int e = __shfl_sync(0xffffffffu, cur_e, j);
            const int g = __shfl_sync(0xffffffffu, cur_g, j);
            const float4 ga = *reinterpret_cast<const float4*>(rp);
            const float4 gb = *reinterpret_cast<const float4*>(rp + 4);
            const float s = eval_sigma(ga.x - fpx, ga.y - fpy, ga.z, ga.w, gb.x);
            const float vis = __expf(-s);
            const float araw = gb.y * vis;
            const float alpha = fminf(kAlphaMax, araw);
            const bool valid = (e < rel_last) && !(s < 0.0f || s > gb.z) && (alpha >= kAlphaMin);
            float fac = 0.0f, w = 0.0f;
            if (valid) {
                const float4* c4 = reinterpret_cast<const float4*>(rp + 8);
#if GG_BWD_DOT4
                float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f, d3 = 0.0f;   // four independent chains
#pragma unroll
                for (int q = 0; q < CP / 4; ++q) {
                    const float4 cc = c4[q];
                    d0 = fmaf(cc.x, vo[4 * q], d0);
                    d1 = fmaf(cc.y, vo[4 * q + 1], d1);
                    d2 = fmaf(cc.z, vo[4 * q + 2], d2);
                    d3 = fmaf(cc.w, vo[4 * q + 3], d3);
                }
                const float dot = (d0 + d1) + (d2 + d3);
#else
                float d0 = 0.0f, d1 = 0.0f;
#pragma unroll
                for (int q = 0; q < CP / 4; ++q) {
                    const float4 cc = c4[q];
                    d0 = fmaf(cc.x, vo[4 * q], d0);
                    d1 = fmaf(cc.y, vo[4 * q + 1], d1);
                    d0 = fmaf(cc.z, vo[4 * q + 2], d0);
                    d1 = fmaf(cc.w, vo[4 * q + 3], d1);
                }
                const float dot = d0 + d1;
#endif
                const float ra = 1.0f / (1.0f - alpha);
                T *= ra;
                fac = alpha * T;
                const float v_alpha = fmaf(dot, T, -R * ra);
                R = fmaf(fac, dot, R);
                w = (araw <= kAlphaMax) ? vis * v_alpha : 0.0f;
            }
            facm[nhit * kHitRow + lane] = fac;
            wm[nhit * kHitRow + lane] = w;
            if ((lane & (kHitRows - 1)) == nhit) fg = g;
            if (++nhit == kHitRows) {
                flush(kHitRows);
                nhit = 0;
                reload_vo();
            }
        }
        __syncwarp();
        cur_k = nxt_k;
        cur_e = nxt_e;
        cur_g = nxt_g;
        buf ^= 1;
    }
    if (nhit) flush(nhit);
}

// geo[g] = {x, y, A/2, B, C/2, o, tau, rcut2};  one thread per row
__global__ void __launch_bounds__(256)
pack_geo_kernel(long long rows, long long n, const float* __restrict__ xys, const float* __restrict__ conics,
                const float* __restrict__ opac, int opac_per_view, float* __restrict__ geo) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const float2 c = __ldg(reinterpret_cast<const float2*>(xys) + i);
    const float A = __ldg(conics + 3 * i), B = __ldg(conics + 3 * i + 1), C = __ldg(conics + 3 * i + 2);
    const float o = __ldg(opac + (opac_per_view ? i : i % n));
    float tau, rcut2;
    geo_tau_rcut(o, A, B, C, tau, rcut2);
    float4* dst = reinterpret_cast<float4*>(geo) + 2 * i;
    dst[0] = make_float4(c.x, c.y, 0.5f * A, B);
    dst[1] = make_float4(0.5f * C, o, tau, rcut2);
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
static int launch_blend_impl(bool backward, const BlendArgs& a, int n_views, bool vec, cudaStream_t st) {
    dim3 grid(a.tiles_x * a.tiles_y, n_views);
    size_t smem = sizeof(float) * (kStages * BATCH * (8 + CP) + (GG_FWD_MMA && CP % 8 == 0 ? (kBlendThreads / 32) * 8 * 32 : 0));
    if (backward && a.hit_words) {
        const size_t wsmem = blend_bwd_warp_smem<CP>();
        dim3 wgrid(2 * a.tiles_x * a.tiles_y, n_views);
        if (vec) {
            GG_CUDA(cudaFuncSetAttribute(blend_bwd_warp_kernel<CP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem));
            blend_bwd_warp_kernel<CP, true><<<wgrid, kBwdWarps * 32, wsmem, st>>>(a);
        } else {
            GG_CUDA(cudaFuncSetAttribute(blend_bwd_warp_kernel<CP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem));
            blend_bwd_warp_kernel<CP, false><<<wgrid, kBwdWarps * 32, wsmem, st>>>(a);
        }
        count_launch();
        return check_launch("blend_bwd_warp_kernel");
    }
    if (backward) {
        smem = blend_bwd_smem<CP, BATCH>();
        if (vec) {
            GG_CUDA(cudaFuncSetAttribute(blend_bwd_kernel<CP, BATCH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            blend_bwd_kernel<CP, BATCH, true><<<grid, kBlendThreads, smem, st>>>(a);
        } else {
            GG_CUDA(cudaFuncSetAttribute(blend_bwd_kernel<CP, BATCH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            blend_bwd_kernel<CP, BATCH, false><<<grid, kBlendThreads, smem, st>>>(a);
        }
    } else {
        if (vec) {
            GG_CUDA(cudaFuncSetAttribute(blend_fwd_kernel<CP, BATCH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            blend_fwd_kernel<CP, BATCH, true><<<grid, kBlendThreads, smem, st>>>(a);
        } else {
            GG_CUDA(cudaFuncSetAttribute(blend_fwd_kernel<CP, BATCH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            blend_fwd_kernel<CP, BATCH, false><<<grid, kBlendThreads, smem, st>>>(a);
        }
    }
    count_launch();
    return check_launch(backward ? "blend_bwd_kernel" : "blend_fwd_kernel");
}

// forward stages BATCH entries per round; the backward also keeps v_out and the per-warp hit
// matrices in shared memory, so it stages 64 (2 CTAs per SM at C <= 24; 3 CTAs with 32-entry
// batches measured slower: twice the barriers)
template <int CP, int BATCH>
static int launch_blend(bool backward, const BlendArgs& a, int n_views, bool vec, cudaStream_t st) {
    if (backward) return launch_blend_impl<CP, 64>(true, a, n_views, vec, st);
    return launch_blend_impl<CP, BATCH>(false, a, n_views, vec, st);
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

// entries staged per round by the forward kernel for a given channel count (see dispatch_blend)
static int fwd_batch_for(int channels) { return channels <= 32 ? 128 : 64; }

// number of uint32 words of the forward's hit-mask table for m intersections over num_tiles tiles
extern "C" size_t gg_blend_hit_words(long long m, long long num_tiles, int channels) {
    const long long fb = fwd_batch_for(channels);
    const long long records = m / fb + num_tiles + 2;
    return (size_t)(records * (kBlendThreads / 32) * (fb / 32));
}

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

// Workload counters of a binned batch (bench.py's roofline accounting, not part of a render): one thread per
// pixel walks its tile list with the forward's tests and counts stats[0] += entries visited up to the stop
// (K of SURVEY 8d, == gg_blend_fwd's pair_counter) and stats[1] += pairs that were blended.
namespace gg {
__global__ void __launch_bounds__(256)
pair_stats_kernel(int n_views, long long n, int img_h, int img_w, int tiles_x, int tiles_y,
                  const int32_t* __restrict__ ids_sorted, const int32_t* __restrict__ tile_ranges,
                  const float* __restrict__ geo, unsigned long long* __restrict__ stats) {
    const int n_tiles = tiles_x * tiles_y;
    const int view = blockIdx.y, tile = blockIdx.x;
    const int tile_y = tile / tiles_x, tile_x = tile - tile_y * tiles_x;
    int tx, ty;
    tile_pixel(tx, ty);
    const int px = tile_x * GG_TILE + tx, py = tile_y * GG_TILE + ty;
    unsigned long long visited = 0, blended = 0;
    if (px < img_w && py < img_h) {
        const int2 range = __ldg(reinterpret_cast<const int2*>(tile_ranges) + (long long)view * n_tiles + tile);
        const float fpx = (float)px, fpy = (float)py;
        float T = 1.0f;
        for (int k = range.x; k < range.y; ++k) {
            ++visited;
            const long long g = (long long)view * n + __ldg(ids_sorted + k);
            const float4 ga = __ldg(reinterpret_cast<const float4*>(geo) + 2 * g);
            const float4 gb = __ldg(reinterpret_cast<const float4*>(geo) + 2 * g + 1);
            const float s = eval_sigma(ga.x - fpx, ga.y - fpy, ga.z, ga.w, gb.x);
            const float alpha = fminf(kAlphaMax, gb.y * __expf(-s));
            if (s < 0.0f || s > gb.z || alpha < kAlphaMin) continue;
            const float next_T = T * (1.0f - alpha);
            if (next_T <= kTStop) break;
            T = next_T;
            ++blended;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        visited += __shfl_xor_sync(0xffffffffu, visited, o);
        blended += __shfl_xor_sync(0xffffffffu, blended, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (visited) atomicAdd(stats, visited);
        if (blended) atomicAdd(stats + 1, blended);
    }
}
}  // namespace gg

extern "C" int gg_blend_pair_stats(int n_views, long long n, int img_h, int img_w, int tiles_x, int tiles_y,
                                   const int32_t* ids_sorted, const int32_t* tile_ranges, const float* geo,
                                   unsigned long long* stats, void* stream) {
    GG_REQUIRE(n_views >= 1 && n >= 1 && img_h > 0 && img_w > 0, "gg_blend_pair_stats: bad sizes");
    GG_REQUIRE(tiles_x == (img_w + GG_TILE - 1) / GG_TILE && tiles_y == (img_h + GG_TILE - 1) / GG_TILE,
               "gg_blend_pair_stats: tile bounds must be ceil(size/16)");
    GG_REQUIRE(ids_sorted && tile_ranges && geo && stats, "gg_blend_pair_stats: null pointer");
    dim3 grid((unsigned)(tiles_x * tiles_y), (unsigned)n_views);
    gg::pair_stats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n_views, n, img_h, img_w, tiles_x, tiles_y, ids_sorted,
                                                                tile_ranges, geo, stats);
    count_launch();
    return check_launch("pair_stats_kernel");
}

extern "C" int gg_blend_fwd(int n_views, long long n, int channels, int color_stride, int colors_per_view,
                            int out_stride, int img_h, int img_w, int tiles_x, int tiles_y,
                            const int32_t* ids_sorted, const int32_t* tile_ranges, const int32_t* tile_order,
                            const float* geo, const float* colors, const float* bg, float* out, float* final_T,
                            int32_t* final_idx, unsigned long long* pair_counter, uint32_t* hit_words,
                            void* stream) {
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
    a.ids_sorted = ids_sorted; a.tile_ranges = tile_ranges; a.tile_order = tile_order; a.geo = geo; a.colors = colors; a.bg = bg;
    a.out = out; a.final_T = final_T; a.final_idx = final_idx; a.pair_counter = pair_counter;
    a.hit_words = hit_words; a.fwd_batch = fwd_batch_for(channels);
    return dispatch_blend(false, a, n_views, (cudaStream_t)stream);
}

extern "C" int gg_blend_bwd(int n_views, long long n, int channels, int color_stride, int colors_per_view,
                            int out_stride, int img_h, int img_w, int tiles_x, int tiles_y,
                            const int32_t* ids_sorted, const int32_t* tile_ranges, const int32_t* tile_order,
                            const float* geo, const float* colors, const float* bg, const float* final_T,
                            const int32_t* final_idx, const float* v_out, const uint32_t* hit_words, float* v_geo,
                            float* v_colors, void* stream) {
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
    a.ids_sorted = ids_sorted; a.tile_ranges = tile_ranges; a.tile_order = tile_order; a.geo = geo; a.colors = colors; a.bg = bg;
    a.final_T = const_cast<float*>(final_T); a.final_idx = const_cast<int32_t*>(final_idx);
    a.v_out = v_out; a.v_geo = v_geo; a.v_colors = v_colors;
    a.hit_words = const_cast<uint32_t*>(hit_words); a.fwd_batch = fwd_batch_for(channels);
    return dispatch_blend(true, a, n_views, (cudaStream_t)stream);
}
