// Fused per-(view, Gaussian) preparation for the multi-view entry point (render_views):
//   activation (exp / normalise / sigmoid) + EWA projection + SH->RGB (+0.5, clamp) + smallest-axis
//   normal + feature copy, written straight into the packed records the blend kernels gather:
//       geo [V*N, 8]   = {x, y, A/2, B, C/2, opacity, tau, rcut2}
//       chan[V*N, CP]  = {r, g, b, depth, nx, ny, nz, feature[0..D), 0-pad}
// and its exact backward (sum over views) to the raw model parameters.  This replaces, in one
// launch each way, the reference's ProjectGaussians + SphericalHarmonics + ~10 ATen elementwise
// kernels per view (nerfstudio/models/gaussian_splatting.py:699-731, :605-619, :742-780).
//
// Compiled with -fmad=false: radii / num_tiles_hit / depth bits must equal what the stand-alone
// projection kernel (and the CPU oracle) produce from the same activated inputs.
#include "gg_common.cuh"
#include "gg_geo.cuh"
#include "gg_math.cuh"
#include "gg_tma.cuh"
#include "gg_b200.h"

namespace gg {

constexpr int kPrepWarps = 4;
constexpr int kPrepThreads = kPrepWarps * 32;

struct PrepArgs {
    int n, n_views, feat_dim, cp, nb, deg_use;
    int img_h, img_w, tiles_x, tiles_y;
    float clip;
    const float* means;          // [N,3]
    const float* log_scales;     // [N,3]
    const float* quats;          // [N,4] raw
    const float* opacity_logit;  // [N]
    const float* sh;             // [N,nb,3]
    const float* features;       // [N,D]
    const float* viewmats;       // [V,12]
    const float* fullmats;       // [V,16]
    const float* intrins;        // [V,4]
    const float* positions;      // [V,3]
};

__device__ __forceinline__ Camera load_camera_regs(const PrepArgs& a, int view) {
    Camera cam;
#pragma unroll
    for (int k = 0; k < 12; ++k) cam.vm[k] = __ldg(a.viewmats + (size_t)view * 12 + k);
#pragma unroll
    for (int k = 0; k < 16; ++k) cam.fm[k] = __ldg(a.fullmats + (size_t)view * 16 + k);
    cam.fx = __ldg(a.intrins + 4 * view);
    cam.fy = __ldg(a.intrins + 4 * view + 1);
    cam.cx = __ldg(a.intrins + 4 * view + 2);
    cam.cy = __ldg(a.intrins + 4 * view + 3);
    return cam;
}

struct Activated {
    float p[3], ls[3], s[3], q_raw[4], qh[4], inv_qnorm, logit, opacity;
    int kmin;
};

__device__ __forceinline__ Activated load_activated(const PrepArgs& a, long long i) {
    Activated g;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        g.p[k] = __ldg(a.means + 3 * i + k);
        g.ls[k] = __ldg(a.log_scales + 3 * i + k);
        g.s[k] = expf(g.ls[k]);
    }
    const float4 q = __ldg(reinterpret_cast<const float4*>(a.quats) + i);
    g.q_raw[0] = q.x; g.q_raw[1] = q.y; g.q_raw[2] = q.z; g.q_raw[3] = q.w;
    const float qn = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    g.inv_qnorm = 1.0f / qn;
#pragma unroll
    for (int k = 0; k < 4; ++k) g.qh[k] = g.q_raw[k] / qn;
    g.logit = __ldg(a.opacity_logit + i);
    g.opacity = 1.0f / (1.0f + expf(-g.logit));
    // argmin of the scales, first minimum on ties (torch.min)
    g.kmin = 0;
    if (g.ls[1] < g.ls[g.kmin]) g.kmin = 1;
    if (g.ls[2] < g.ls[g.kmin]) g.kmin = 2;
    return g;
}

// (forcing 6 / 7 resident CTAs -- 80 / 72 registers, spills -- measured slower: 0.079 / 0.080 vs 0.073 ms at config 1)
__global__ void __launch_bounds__(kPrepThreads)
prepare_views_kernel(const PrepArgs a, float* __restrict__ geo, float* __restrict__ chan,
                     float* __restrict__ depths, int32_t* __restrict__ radii, int32_t* __restrict__ num_tiles_hit,
                     float* __restrict__ scales_out, float* __restrict__ quats_out, int slab_row, int phase,
                     int feat_bulk) {
    // phase 0: everything; 1: geometry only (geo, depths, radii, num_tiles_hit); 2: channel rows
    // only, from the radii / depths phase 1 left behind.  The split lets the host read the number
    // of intersections (needed to size the sort) while phase 2 keeps the GPU busy.
    extern __shared__ __align__(16) float sm[];
    const int view = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int fcols = feat_bulk ? a.feat_dim : 0;          // feature rows staged next to the SH slab
    float* slab = sm + (size_t)warp * 32 * (slab_row + fcols);
    float* fbuf = slab + 32 * slab_row;                    // [32][D]
    const long long first = ((long long)blockIdx.x * kPrepWarps + warp) * 32;
    if (first >= a.n) return;  // whole warp; no block-level barrier is used in this kernel
    const long long i = first + lane;
    const bool active = i < a.n;
    const int rows_here = (int)min((long long)32, a.n - first);
    const long long vrow = (long long)view * a.n + i;
    // ---- SH coefficients of the warp's 32 Gaussians: one contiguous span.  A full warp moves it with a single
    // bulk copy (TMA engine), issued BEFORE the projection arithmetic so that the biggest transfer of the kernel
    // (9.6 KB per warp) is in flight while the warp computes; the warp's feature rows (one contiguous span too)
    // ride on the same barrier, so that the row assembly below does not pay a second DRAM latency ----
    const int row = a.nb * 3;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + (size_t)kPrepWarps * 32 * (slab_row + fcols)) + warp;
    const bool bulk = rows_here == 32 && phase != 1;
    const bool fbulk = bulk && fcols > 0;
    if (bulk) {
        if (lane == 0) {
            mbar_init(bar, fbulk ? 2 : 1);
            bulk_load(slab, a.sh + first * row, (uint32_t)(32 * row * sizeof(float)), bar);
            if (fbulk) bulk_load(fbuf, a.features + first * fcols, (uint32_t)(32 * fcols * sizeof(float)), bar);
        }
        __syncwarp();
    }
    const Camera cam = load_camera_regs(a, view);

    Activated g;
    ProjOut o;
    o.tiles = 0;
    if (active && phase == 2) {
        g = load_activated(a, i);
        o.tiles = radii[vrow] > 0 ? 1 : 0;
        o.depth = depths[vrow];
    } else if (active) {
        g = load_activated(a, i);
        o = project_one(g.p, g.s, 1.0f, g.qh, cam, a.img_h, a.img_w, a.tiles_x, a.tiles_y, a.clip);
        depths[vrow] = o.depth;
        radii[vrow] = o.radius;
        num_tiles_hit[vrow] = o.tiles;
        const bool vis = o.tiles > 0;
        float tau = -1.0f, rcut2 = -1.0f;
        if (vis) geo_tau_rcut(g.opacity, o.conic[0], o.conic[1], o.conic[2], tau, rcut2);
        float4* gd = reinterpret_cast<float4*>(geo) + 2 * vrow;
        gd[0] = make_float4(o.ux, o.uy, 0.5f * o.conic[0], o.conic[1]);
        gd[1] = make_float4(0.5f * o.conic[2], vis ? g.opacity : 0.0f, tau, rcut2);
        if (view == 0 && scales_out) {
#pragma unroll
            for (int k = 0; k < 3; ++k) scales_out[3 * i + k] = g.s[k];
        }
        if (view == 0 && quats_out)
            reinterpret_cast<float4*>(quats_out)[i] = make_float4(g.qh[0], g.qh[1], g.qh[2], g.qh[3]);
    }
    const bool vis = active && o.tiles > 0;
    if (phase == 1) return;
    if (!__any_sync(0xffffffffu, vis)) {   // nothing of this warp reaches a tile list
        if (bulk) mbar_wait(bar, 0);       // the copy must have landed before this shared memory can be handed on
        return;
    }
    if (!bulk) {
        const float* gspan = a.sh + first * row;
        const int span = rows_here * row;
        for (int k = lane; k < span; k += 32) slab[k] = __ldg(gspan + k);
    }
    float rgb[3] = {0.f, 0.f, 0.f}, nrm[3] = {0.f, 0.f, 0.f};
    float Y[25];
    if (vis) {
        sh_basis(a.deg_use, g.p[0] - __ldg(a.positions + 3 * view), g.p[1] - __ldg(a.positions + 3 * view + 1),
                 g.p[2] - __ldg(a.positions + 3 * view + 2), Y);
        const Rot3 R = quat_to_rot(g.qh[0], g.qh[1], g.qh[2], g.qh[3]);
        nrm[0] = R.m[g.kmin]; nrm[1] = R.m[3 + g.kmin]; nrm[2] = R.m[6 + g.kmin];
    }
    if (bulk) mbar_wait(bar, 0); else __syncwarp();
    if (vis) {
        const int nuse = sh_num_bases(a.deg_use);
        const float* cf = slab + lane * row;
#pragma unroll
        for (int b = 0; b < 25; ++b) {  // fully unrolled: Y stays in registers
            if (b < nuse) {
#pragma unroll
                for (int c = 0; c < 3; ++c) rgb[c] = rgb[c] + Y[b] * cf[3 * b + c];
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) rgb[c] = fminf(1.0f, fmaxf(0.0f, rgb[c] + 0.5f));
    }
    __syncwarp();
    // ---- assemble the channel rows in shared memory, store them as one coalesced span ----
    {
        float* r = slab + lane * a.cp;
        if (vis) {
            r[0] = rgb[0]; r[1] = rgb[1]; r[2] = rgb[2]; r[3] = o.depth;
            r[4] = nrm[0]; r[5] = nrm[1]; r[6] = nrm[2];
            for (int d = 7 + a.feat_dim; d < a.cp; ++d) r[d] = 0.0f;
        } else {
            for (int d = 0; d < 7; ++d) r[d] = 0.0f;
            for (int d = 7 + a.feat_dim; d < a.cp; ++d) r[d] = 0.0f;
        }
        // features of the warp's 32 Gaussians are one contiguous span: coalesced loads, zero for
        // the culled rows
        const unsigned vismask = __ballot_sync(0xffffffffu, vis);
        const int D = a.feat_dim;
        const float* fspan = a.features + first * D;
        for (int k = lane; k < rows_here * D; k += 32) {
            const int l = k / D, d = k - l * D;
            slab[l * a.cp + 7 + d] = ((vismask >> l) & 1u) ? (fbulk ? fbuf[k] : __ldg(fspan + k)) : 0.0f;
        }
    }
    __syncwarp();
    {
        float* gspan = chan + ((long long)view * a.n + first) * a.cp;
        const int nvec = (rows_here * a.cp) >> 2;  // cp is a multiple of 4
        float4* g4 = reinterpret_cast<float4*>(gspan);
        const float4* s4 = reinterpret_cast<const float4*>(slab);
        for (int k = lane; k < nvec; k += 32) g4[k] = s4[k];
    }
}

// Channel rows for ALL views of a batch in one pass over the Gaussians (phase 2 of the fused
// path): the 300-byte SH row, the feature row and the quaternion are read once per Gaussian and
// reused for every view, instead of once per (view, Gaussian).  Per view only the depth is read and the
// CP-float row is written (coalesced through shared memory).
//
// Shared memory per warp: slab [32][row | 1] SH coefficients -- the ODD row stride puts coefficient k of the 32
// lanes' rows in 32 different banks (with the dense stride of 48 floats every read was a 16-way conflict, and the
// V x 48 reads per Gaussian were the whole kernel) -- and rbuf [32][cp], the channel rows: features, normal and
// padding are written once, each view only rewrites rgb + depth of its visible rows; rows culled in a view are
// stored as zeros straight from the copy-out loop.
__device__ __forceinline__ void cp_async_f32(float* smem_dst, const float* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gmem_src) : "memory");
}

__host__ __device__ __forceinline__ int chan_slab_stride(int row) { return row | 1; }

// rgb += sum_b Y[b] * coefficient row b, in the order of the per-view kernel (b ascending, channel inner); the
// basis count is a template parameter so that the unrolled loop carries no per-term predicate
template <int NUSE>
__device__ __forceinline__ void sh_dot(const float* __restrict__ Y, const float* __restrict__ cf, float (&rgb)[3]) {
#pragma unroll
    for (int b = 0; b < NUSE; ++b) {
#pragma unroll
        for (int c = 0; c < 3; ++c) rgb[c] = rgb[c] + Y[b] * cf[3 * b + c];
    }
}

__global__ void __launch_bounds__(kPrepThreads)
prepare_chan_kernel(const PrepArgs a, float* __restrict__ chan, const float* __restrict__ depths,
                    const int32_t* __restrict__ radii) {
    extern __shared__ __align__(16) float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = a.nb * 3, D = a.feat_dim, cp = a.cp;
    const int srow = chan_slab_stride(row);
    float* slab = sm + (size_t)warp * 32 * (srow + cp);   // [32][srow]  SH coefficients
    float* rbuf = slab + 32 * srow;                       // [32][cp]    channel rows
    const long long first = ((long long)blockIdx.x * kPrepWarps + warp) * 32;
    if (first >= a.n) return;
    const long long i = first + lane;
    const bool active = i < a.n;
    const int rows_here = (int)min((long long)32, a.n - first);
    // visibility of this Gaussian in each view (bit v); warps that are invisible everywhere stop here
    unsigned vismask = 0;
    if (active)
        for (int v = 0; v < a.n_views; ++v)
            if (radii[(long long)v * a.n + i] > 0) vismask |= 1u << (v & 31);
    const bool many_views = a.n_views > 32;  // the bit mask is only a hint then
    if (!many_views && !__any_sync(0xffffffffu, vismask != 0)) return;
    {   // SH coefficients: coalesced 4-byte async copies into the odd-stride rows
        const float* gspan = a.sh + first * row;
        int l = 0, j = lane;
        while (j >= row) { j -= row; ++l; }
        for (int k = lane; k < rows_here * row; k += 32) {
            cp_async_f32(slab + l * srow + j, gspan + k);
            j += 32;
            while (j >= row) { j -= row; ++l; }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    }
    {   // feature rows (one coalesced span) and the zero padding
        const float* fspan = a.features + first * D;
        for (int k = lane; k < rows_here * D; k += 32) {
            const int l = k / D, d = k - l * D;
            rbuf[l * cp + 7 + d] = __ldg(fspan + k);
        }
        for (int d = 7 + D; d < cp; ++d) rbuf[lane * cp + d] = 0.0f;
    }
    float p[3] = {0.f, 0.f, 0.f};
    float* r = rbuf + lane * cp;
    if (active) {
        const Activated g = load_activated(a, i);
        p[0] = g.p[0]; p[1] = g.p[1]; p[2] = g.p[2];
        const Rot3 R = quat_to_rot(g.qh[0], g.qh[1], g.qh[2], g.qh[3]);
        r[4] = R.m[g.kmin]; r[5] = R.m[3 + g.kmin]; r[6] = R.m[6 + g.kmin];
    }
    auto visible = [&](int v) -> bool {
        if (!many_views) return (vismask >> v) & 1u;
        return active && radii[(long long)v * a.n + i] > 0;
    };
    // depth of the next view is requested before this view's rows are built
    bool vis_next = visible(0);
    float dep_next = vis_next ? __ldg(depths + i) : 0.0f;
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncwarp();
    const int nuse = sh_num_bases(a.deg_use);
    const float* cf = slab + lane * srow;
    const int vec_per_row = cp >> 2;                 // cp is a multiple of 4
    const int nvec = rows_here * vec_per_row;
    const int row0 = lane / vec_per_row, piece0 = lane - row0 * vec_per_row;      // row / piece of float4 number `lane`
    const int row_step = 32 / vec_per_row, piece_step = 32 - row_step * vec_per_row;   // ... and of 32 pieces further
    for (int v = 0; v < a.n_views; ++v) {
        const bool vis = vis_next;
        const float dep = dep_next;
        if (v + 1 < a.n_views) {
            vis_next = visible(v + 1);
            dep_next = vis_next ? __ldg(depths + (long long)(v + 1) * a.n + i) : 0.0f;
        }
        const unsigned vm = __ballot_sync(0xffffffffu, vis);
        if (!vm) continue;
        if (vis) {
            float Y[25];
            sh_basis(a.deg_use, p[0] - __ldg(a.positions + 3 * v), p[1] - __ldg(a.positions + 3 * v + 1),
                     p[2] - __ldg(a.positions + 3 * v + 2), Y);
            float rgb[3] = {0.f, 0.f, 0.f};
            switch (nuse) {   // warp-uniform
                case 1: sh_dot<1>(Y, cf, rgb); break;
                case 4: sh_dot<4>(Y, cf, rgb); break;
                case 9: sh_dot<9>(Y, cf, rgb); break;
                case 16: sh_dot<16>(Y, cf, rgb); break;
                default: sh_dot<25>(Y, cf, rgb); break;
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) r[c] = fminf(1.0f, fmaxf(0.0f, rgb[c] + 0.5f));
            r[3] = dep;
        }
        __syncwarp();
        float4* g4 = reinterpret_cast<float4*>(chan + ((long long)v * a.n + first) * cp);
        const float4* s4 = reinterpret_cast<const float4*>(rbuf);
        // piece k of the span belongs to row k / vec_per_row: tracked incrementally (no division per piece)
        int l = row0, q = piece0;
        for (int k = lane; k < nvec; k += 32) {
            g4[k] = ((vm >> l) & 1u) ? s4[k] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            l += row_step; q += piece_step;
            if (q >= vec_per_row) { q -= vec_per_row; ++l; }
        }
        __syncwarp();
    }
}

// Sum over views of the vector-Jacobian product of prepare_views_kernel.
//   v_geo [V*N, 8] : v_x, v_y, v_A, v_B, v_C (w.r.t. the conic, not its halves), v_opacity
//   v_chan[V*N, CP]: v_rgb(3), v_depth, v_normal(3), v_feature(D)
__global__ void __launch_bounds__(kPrepThreads, 4)
prepare_views_bwd_kernel(const PrepArgs a, const float* __restrict__ geo, const float* __restrict__ chan,
                         const int32_t* __restrict__ radii, const float* __restrict__ v_geo,
                         const float* __restrict__ v_chan, float* __restrict__ v_means,
                         float* __restrict__ v_log_scales, float* __restrict__ v_quats,
                         float* __restrict__ v_opacity_logit, float* __restrict__ v_sh,
                         float* __restrict__ v_features, float* __restrict__ v_rgb_views) {
    // v_rgb_views != nullptr: the SH gradient is left to sh_grad_views_kernel; this kernel only emits its
    // per-view factor, the clamp-masked colour gradient [V*N, 3] (and needs no SH slab)
    extern __shared__ __align__(16) float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool defer_sh = v_rgb_views != nullptr;
    const int row = defer_sh ? 0 : a.nb * 3;
    const int D = a.feat_dim;
    const int Dp = D | 1;  // odd row stride: lane-per-row accumulation hits 32 different banks
    float* slab = sm + (size_t)warp * 32 * (row + Dp);  // [32][row] SH grads, then [32][Dp] feature grads
    float* fslab = slab + 32 * row;
    const long long first = ((long long)blockIdx.x * kPrepWarps + warp) * 32;
    if (first >= a.n) return;
    const long long i = first + lane;
    const bool active = i < a.n;
    const int rows_here = (int)min((long long)32, a.n - first);
    for (int k = lane; k < 32 * (row + Dp); k += 32) slab[k] = 0.0f;
    __syncwarp();

    Activated g;
    if (active) g = load_activated(a, i);
    float gm[3] = {0.f, 0.f, 0.f}, gs[3] = {0.f, 0.f, 0.f}, gq[4] = {0.f, 0.f, 0.f, 0.f}, go = 0.0f;
    const int nuse = sh_num_bases(a.deg_use);
    for (int view = 0; view < a.n_views; ++view) {
        const long long vrow = (long long)view * a.n + i;
        if (!(active && radii[vrow] > 0)) {
            if (defer_sh && active) {
                v_rgb_views[3 * vrow] = 0.0f; v_rgb_views[3 * vrow + 1] = 0.0f; v_rgb_views[3 * vrow + 2] = 0.0f;
            }
            continue;
        }
        const Camera cam = load_camera_regs(a, view);
        const float4 ga = __ldg(reinterpret_cast<const float4*>(geo) + 2 * vrow);
        const float4 gb = __ldg(reinterpret_cast<const float4*>(geo) + 2 * vrow + 1);
        const float4 va = __ldg(reinterpret_cast<const float4*>(v_geo) + 2 * vrow);
        const float4 vb = __ldg(reinterpret_cast<const float4*>(v_geo) + 2 * vrow + 1);
        const float* vc = v_chan + vrow * a.cp;
        const float4 c0 = __ldg(reinterpret_cast<const float4*>(vc));
        const float4 c1 = __ldg(reinterpret_cast<const float4*>(vc) + 1);
        const float conic[3] = {2.0f * ga.z, ga.w, 2.0f * gb.x};
        const float v_xy[2] = {va.x, va.y};
        const float v_conic[3] = {va.z, va.w, vb.x};
        go += vb.y;
        // normal channel: column kmin of R(qh)
        float vR[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) vR[k] = 0.0f;
        vR[g.kmin] = c1.x; vR[3 + g.kmin] = c1.y; vR[6 + g.kmin] = c1.z;
        const ProjGrad pg = project_bwd_one(g.p, g.s, 1.0f, g.qh, cam, a.img_h, a.img_w, conic, v_xy, c0.w, v_conic, vR);
#pragma unroll
        for (int k = 0; k < 3; ++k) { gm[k] += pg.v_mean[k]; gs[k] += pg.v_scale[k]; }
#pragma unroll
        for (int k = 0; k < 4; ++k) gq[k] += pg.v_quat[k];
        // SH: rgb = clamp(sh + 0.5, 0, 1); the clamp passes gradient strictly inside (0,1)
        const float* cr = chan + vrow * a.cp;
        float vrgb[3] = {c0.x, c0.y, c0.z};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = __ldg(cr + c);
            if (!(v > 0.0f && v < 1.0f)) vrgb[c] = 0.0f;
        }
        if (defer_sh) {
            v_rgb_views[3 * vrow] = vrgb[0]; v_rgb_views[3 * vrow + 1] = vrgb[1]; v_rgb_views[3 * vrow + 2] = vrgb[2];
        } else {
            float Y[25];
            sh_basis(a.deg_use, g.p[0] - __ldg(a.positions + 3 * view), g.p[1] - __ldg(a.positions + 3 * view + 1),
                     g.p[2] - __ldg(a.positions + 3 * view + 2), Y);
            float* sr = slab + lane * row;
            for (int b = 0; b < nuse; ++b) {
#pragma unroll
                for (int c = 0; c < 3; ++c) sr[3 * b + c] += Y[b] * vrgb[c];
            }
        }
        // feature gradients: columns 7.. of the row, fetched as aligned 16-byte pieces
        float* fr = fslab + lane * Dp;
        if (D > 0) fr[0] += c1.w;
        for (int q = 2; 4 * q < 7 + D; ++q) {
            const float4 c = __ldg(reinterpret_cast<const float4*>(vc) + q);
            const float e[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int d = 4 * q + t - 7;
                if (d < D) fr[d] += e[t];
            }
        }
    }
    if (active) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            v_means[3 * i + k] = gm[k];
            v_log_scales[3 * i + k] = gs[k] * g.s[k];
        }
        // qh = q / |q|
        const float dotq = g.qh[0] * gq[0] + g.qh[1] * gq[1] + g.qh[2] * gq[2] + g.qh[3] * gq[3];
        reinterpret_cast<float4*>(v_quats)[i] =
            make_float4((gq[0] - g.qh[0] * dotq) * g.inv_qnorm, (gq[1] - g.qh[1] * dotq) * g.inv_qnorm,
                        (gq[2] - g.qh[2] * dotq) * g.inv_qnorm, (gq[3] - g.qh[3] * dotq) * g.inv_qnorm);
        v_opacity_logit[i] = go * g.opacity * (1.0f - g.opacity);
    }
    __syncwarp();
    if (defer_sh) {
        // nothing to store for the SH table
    } else if (rows_here == 32) {
        // the warp's [32, 3*nb] gradient rows are one contiguous, 16-byte aligned span: one bulk store
        if (lane == 0) bulk_store_and_wait(v_sh + first * row, slab, (uint32_t)(32 * row * sizeof(float)));
        __syncwarp();
    } else {
        float* gspan = v_sh + first * row;
        for (int k = lane; k < rows_here * row; k += 32) gspan[k] = slab[k];
    }
    if (D > 0) {
        float* gspan = v_features + first * D;
        const int span = rows_here * D;
        for (int k = lane; k < span; k += 32) {
            const int l = k / D;
            gspan[k] = fslab[l * Dp + (k - l * D)];
        }
    }
}

// SH-coefficient gradient from its per-view factors: v_sh[i] = sum_v Y(dir_v(i)) (x) v_rgb[v][i].
// The gradient of the 3*nb coefficients is an outer product per view, so a data-parallel step only has
// to exchange the 3 floats per (view, Gaussian) -- all-gathered with the camera centres -- instead of
// all-reducing 3*nb floats per Gaussian (75 at degree 4); every rank then rebuilds the full sum here.
template <int NB>
__global__ void __launch_bounds__(kPrepThreads, 4)
sh_grad_views_kernel(int n, int n_views, int deg_use, const float* __restrict__ means,
                     const float* __restrict__ positions, const float* __restrict__ v_rgb, float* __restrict__ v_sh) {
    // Accumulators live in registers (3*NB per lane, all indices compile-time); shared memory is touched
    // once, to turn the lane-per-row result into one contiguous span for the bulk store.
    extern __shared__ __align__(16) float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int row = NB * 3;
    float* slab = sm + (size_t)warp * 32 * row;
    const long long first = ((long long)blockIdx.x * kPrepWarps + warp) * 32;
    if (first >= n) return;
    const long long i = first + lane;
    const bool active = i < n;
    const int rows_here = (int)min((long long)32, n - first);
    float acc[row];
#pragma unroll
    for (int k = 0; k < row; ++k) acc[k] = 0.0f;
    if (active) {
        const float px = __ldg(means + 3 * i), py = __ldg(means + 3 * i + 1), pz = __ldg(means + 3 * i + 2);
        const int nuse = sh_num_bases(deg_use);
        constexpr int kPre = NB > 16 ? 4 : 8;  // factors of kPre views are in flight before the first is used
        for (int v0 = 0; v0 < n_views; v0 += kPre) {
            float f[kPre][3];
#pragma unroll
            for (int u = 0; u < kPre; ++u) {
                const int v = v0 + u;
                const float* g = v_rgb + 3 * ((long long)(v < n_views ? v : 0) * n + i);
                f[u][0] = v < n_views ? __ldg(g) : 0.0f;
                f[u][1] = v < n_views ? __ldg(g + 1) : 0.0f;
                f[u][2] = v < n_views ? __ldg(g + 2) : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < kPre; ++u) {
                if (f[u][0] == 0.0f && f[u][1] == 0.0f && f[u][2] == 0.0f) continue;  // culled / clamped / unlit
                const int v = v0 + u;
                float Y[25];
                sh_basis(deg_use, px - __ldg(positions + 3 * v), py - __ldg(positions + 3 * v + 1),
                         pz - __ldg(positions + 3 * v + 2), Y);
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    if (b < nuse) {
                        acc[3 * b] += Y[b] * f[u][0]; acc[3 * b + 1] += Y[b] * f[u][1]; acc[3 * b + 2] += Y[b] * f[u][2];
                    }
                }
            }
        }
    }
    {
        float* sr = slab + lane * row;  // odd row stride for NB = 1, 9, 25; 2-way conflicts at most otherwise
#pragma unroll
        for (int k = 0; k < row; ++k) sr[k] = acc[k];
    }
    __syncwarp();
    if (rows_here == 32) {
        if (lane == 0) bulk_store_and_wait(v_sh + first * row, slab, (uint32_t)(32 * row * sizeof(float)));
        __syncwarp();
    } else {
        float* gspan = v_sh + first * row;
        for (int k = lane; k < rows_here * row; k += 32) gspan[k] = slab[k];
    }
}

static int fill_args(PrepArgs& a, int n, int n_views, int feat_dim, int cp, int degree, int degrees_to_use,
                     const float* means, const float* log_scales, const float* quats, const float* opacity_logit,
                     const float* sh, const float* features, const float* viewmats, const float* fullmats,
                     const float* intrins, const float* positions, int img_h, int img_w, int tiles_x, int tiles_y,
                     float clip) {
    GG_REQUIRE(n >= 1 && n_views >= 1, "gg_prepare_views: need n >= 1 and n_views >= 1");
    GG_REQUIRE(feat_dim >= 0 && feat_dim <= 64, "gg_prepare_views: feature dim must be in [0, 64]");
    GG_REQUIRE(cp % 4 == 0 && cp >= 7 + feat_dim && cp <= 72, "gg_prepare_views: cp must be a multiple of 4, >= 7 + D");
    GG_REQUIRE(degree >= 0 && degree <= 4 && degrees_to_use >= 0 && degrees_to_use <= degree,
               "gg_prepare_views: need 0 <= degrees_to_use <= degree <= 4");
    GG_REQUIRE(means && log_scales && quats && opacity_logit && sh && (features || feat_dim == 0) && viewmats &&
                   fullmats && intrins && positions,
               "gg_prepare_views: null input pointer");
    GG_REQUIRE(((uintptr_t)quats & 15) == 0 && ((uintptr_t)sh & 15) == 0, "gg_prepare_views: quats/sh misaligned");
    GG_REQUIRE(img_h > 0 && img_w > 0 && tiles_x > 0 && tiles_y > 0, "gg_prepare_views: bad image / tile bounds");
    a.n = n; a.n_views = n_views; a.feat_dim = feat_dim; a.cp = cp; a.nb = sh_num_bases(degree);
    a.deg_use = degrees_to_use; a.img_h = img_h; a.img_w = img_w; a.tiles_x = tiles_x; a.tiles_y = tiles_y;
    a.clip = clip; a.means = means; a.log_scales = log_scales; a.quats = quats; a.opacity_logit = opacity_logit;
    a.sh = sh; a.features = features; a.viewmats = viewmats; a.fullmats = fullmats; a.intrins = intrins;
    a.positions = positions;
    return GG_OK;
}

}  // namespace gg

using namespace gg;

extern "C" int gg_prepare_views(int n, int n_views, int feat_dim, int cp, int degree, int degrees_to_use,
                                const float* means, const float* log_scales, const float* quats,
                                const float* opacity_logit, const float* sh_coeffs, const float* features,
                                const float* viewmats, const float* fullmats, const float* intrins,
                                const float* positions, int img_h, int img_w, int tiles_x, int tiles_y,
                                float clip_thresh, float* geo, float* chan, float* depths, int32_t* radii,
                                int32_t* num_tiles_hit, float* scales_out, float* quats_out, int phase,
                                void* stream) {
    PrepArgs a;
    GG_REQUIRE(phase >= 0 && phase <= 2, "gg_prepare_views: phase must be 0, 1 or 2");
    const int rc = fill_args(a, n, n_views, feat_dim, cp, degree, degrees_to_use, means, log_scales, quats,
                             opacity_logit, sh_coeffs, features, viewmats, fullmats, intrins, positions, img_h, img_w,
                             tiles_x, tiles_y, clip_thresh);
    if (rc != GG_OK) return rc;
    GG_REQUIRE(geo && chan && depths && radii && num_tiles_hit, "gg_prepare_views: null output pointer");
    GG_REQUIRE(((uintptr_t)geo & 15) == 0 && ((uintptr_t)chan & 15) == 0, "gg_prepare_views: geo/chan misaligned");
    if (phase == 2 && n_views > 1) {  // one view: the per-view kernel below has the better occupancy
        const size_t smem2 = sizeof(float) * kPrepWarps * 32 * (size_t)(chan_slab_stride(a.nb * 3) + cp);
        GG_CUDA(cudaFuncSetAttribute(prepare_chan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        prepare_chan_kernel<<<div_up(n, kPrepThreads), kPrepThreads, smem2, (cudaStream_t)stream>>>(a, chan, depths,
                                                                                                   radii);
        count_launch();
        return check_launch("prepare_chan_kernel");
    }
    const int slab_row = a.nb * 3 > cp ? a.nb * 3 : cp;
    // the feature rows travel as a bulk copy too when their spans are 16-byte aligned (32 rows = 128 D bytes)
    const int feat_bulk = (feat_dim > 0 && phase != 1 && ((uintptr_t)features & 15) == 0) ? 1 : 0;
    const size_t smem = sizeof(float) * kPrepWarps * 32 * (size_t)(slab_row + (feat_bulk ? feat_dim : 0)) +
                        sizeof(uint64_t) * kPrepWarps;
    GG_CUDA(cudaFuncSetAttribute(prepare_views_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(div_up(n, kPrepThreads), n_views);
    prepare_views_kernel<<<grid, kPrepThreads, smem, (cudaStream_t)stream>>>(a, geo, chan, depths, radii,
                                                                             num_tiles_hit, scales_out, quats_out,
                                                                             slab_row, phase, feat_bulk);
    count_launch();
    return check_launch("prepare_views_kernel");
}

extern "C" int gg_prepare_views_bwd(int n, int n_views, int feat_dim, int cp, int degree, int degrees_to_use,
                                    const float* means, const float* log_scales, const float* quats,
                                    const float* opacity_logit, const float* features, const float* viewmats,
                                    const float* fullmats, const float* intrins, const float* positions, int img_h,
                                    int img_w, const float* geo, const float* chan, const int32_t* radii,
                                    const float* v_geo, const float* v_chan, float* v_means, float* v_log_scales,
                                    float* v_quats, float* v_opacity_logit, float* v_sh_coeffs, float* v_features,
                                    float* v_rgb_views, void* stream) {
    PrepArgs a;
    // the SH table itself is not read by the backward; pass a gradient buffer for the alignment check
    GG_REQUIRE(v_sh_coeffs || v_rgb_views, "gg_prepare_views_bwd: need v_sh_coeffs or v_rgb_views");
    const int rc = fill_args(a, n, n_views, feat_dim, cp, degree, degrees_to_use, means, log_scales, quats,
                             opacity_logit, v_rgb_views ? quats : v_sh_coeffs, features, viewmats, fullmats, intrins,
                             positions, img_h, img_w, 1, 1, 0.0f);
    if (rc != GG_OK) return rc;
    GG_REQUIRE(geo && chan && radii && v_geo && v_chan, "gg_prepare_views_bwd: null input pointer");
    GG_REQUIRE(v_means && v_log_scales && v_quats && v_opacity_logit && (v_features || feat_dim == 0),
               "gg_prepare_views_bwd: null output pointer");
    GG_REQUIRE(((uintptr_t)v_quats & 15) == 0 && ((uintptr_t)v_geo & 15) == 0 && ((uintptr_t)v_chan & 15) == 0,
               "gg_prepare_views_bwd: misaligned");
    const size_t smem = sizeof(float) * kPrepWarps * 32 * (size_t)((v_rgb_views ? 0 : a.nb * 3) + (feat_dim | 1));
    GG_CUDA(cudaFuncSetAttribute(prepare_views_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    prepare_views_bwd_kernel<<<div_up(n, kPrepThreads), kPrepThreads, smem, (cudaStream_t)stream>>>(
        a, geo, chan, radii, v_geo, v_chan, v_means, v_log_scales, v_quats, v_opacity_logit, v_sh_coeffs, v_features,
        v_rgb_views);
    count_launch();
    return check_launch("prepare_views_bwd_kernel");
}

extern "C" int gg_sh_grad_from_views(int n, int n_views, int degree, int degrees_to_use, const float* means,
                                     const float* positions, const float* v_rgb_views, float* v_sh_coeffs,
                                     void* stream) {
    GG_REQUIRE(n >= 1 && n_views >= 1, "gg_sh_grad_from_views: need n >= 1 and n_views >= 1");
    GG_REQUIRE(degree >= 0 && degree <= 4 && degrees_to_use >= 0 && degrees_to_use <= degree,
               "gg_sh_grad_from_views: need 0 <= degrees_to_use <= degree <= 4");
    GG_REQUIRE(means && positions && v_rgb_views && v_sh_coeffs, "gg_sh_grad_from_views: null pointer");
    GG_REQUIRE(((uintptr_t)v_sh_coeffs & 15) == 0, "gg_sh_grad_from_views: v_sh_coeffs misaligned");
    const int nb = sh_num_bases(degree);
    const size_t smem = sizeof(float) * kPrepWarps * 32 * (size_t)(nb * 3);
    auto launch = [&](auto kernel) -> int {
        GG_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kernel<<<div_up(n, kPrepThreads), kPrepThreads, smem, (cudaStream_t)stream>>>(
            n, n_views, degrees_to_use, means, positions, v_rgb_views, v_sh_coeffs);
        count_launch();
        return check_launch("sh_grad_views_kernel");
    };
    switch (degree) {
        case 0: return launch(sh_grad_views_kernel<1>);
        case 1: return launch(sh_grad_views_kernel<4>);
        case 2: return launch(sh_grad_views_kernel<9>);
        case 3: return launch(sh_grad_views_kernel<16>);
        default: return launch(sh_grad_views_kernel<25>);
    }
}
