// Tile-first binning without a host read: per-tile counts -> scan -> placement -> per-tile sort.
//
// Same result as gsplat 0.1.0's compute_cumulative_intersects + bin_and_sort_gaussians + get_tile_bin_edges
// (re-run inside every rasterize forward of the reference, nerfstudio/models/gaussian_splatting.py:735,747,759,
// 773): the Gaussian ids of every tile in (depth bits, id) order and the per-tile [start, end) ranges -- bit for
// bit what sorting the 64-bit tile|depth keys with ties broken by id gives.  The route differs:
//
//   1. tile_hist_kernel         each CTA owns a slice of one view's Gaussians and counts their tile entries in a
//                               shared-memory histogram (one row of T counters per CTA, stored to global)
//   2. tile_prefix_scan_kernel  per tile: exclusive prefix over the CTA rows; the last CTA to finish then scans
//                               the V*T tile totals -> tile_ranges, M and the capacity check (on the device),
//                               longest-first tile order
//   3. tile_place_kernel        same slices again: every entry takes the next slot of (its tile, its CTA) from a
//                               shared-memory cursor and stores (depth bits << 32 | id); no global atomics at all
//   4. tile_bucket_sort_kernel  one CTA per tile: bucket by a monotone function of the depth bits (shared-memory
//                               counting sort, a few entries per bucket), then every entry finds its rank inside
//                               its bucket by comparing the 64-bit records -- ties on depth fall to the id, so the
//                               result never depends on the placement order.  Tiles whose depths cluster in few
//                               buckets are handed to tile_sort_kernel (LSD radix on the varying depth bits, equal
//                               depths re-sorted on id) and tiles beyond shared memory to a bitonic network in
//                               global memory (tile_sort_big_kernel).
//   (tile_count_kernel / tile_scatter_kernel: global-atomic variants of 1 and 3 for tile grids that do not fit
//    a shared-memory histogram.)
//
// Nothing here needs the intersection count M on the host: buffers are sized by a caller-chosen capacity and
// info[] = {M, overflow, longest tile, 0} is written on the device (and copied to pinned host memory if asked).
// When M exceeds the capacity every tile range is (0,0) -- the blend kernels then render background only --
// and the overflow flag tells the caller to grow the buffers and repeat.
//
// Why not a radix sort: ranking is what a radix pass pays for (~50 G key-passes/s on this part whatever the
// formulation); the depth-first path needs 4 passes over V*N keys + 2 over M, an LSD sort per tile 4 over M.
// Here no entry is ever ranked against more than its bucket's handful of neighbours.
//
// Compiled with -fmad=false (tile_box() must reproduce the projection kernel's tile bbox).
#include "gg_common.cuh"
#include "gg_math.cuh"
#include "gg_b200.h"

namespace gg {

constexpr int kOrderBins2 = 1024;
constexpr int kBigBox = 16;       // bboxes with more tiles than this are spread over the warp
constexpr int kSortMaxSmem = 16384;  // longest segment the radix kernel sorts in shared memory
constexpr int kBucketMax = 24576;    // longest segment the bucket sort holds in shared memory

__device__ __forceinline__ int len_bucket(int len) { return min(kOrderBins2 - 1, len >> 3); }

struct BoxOf {
    int x0, y0, w, cnt, tile0;
};

// bbox of row i (= view * n + g) or an empty box
__device__ __forceinline__ BoxOf load_box(long long i, long long total, int n, const float* __restrict__ xys, int xy_stride,
                                          const int32_t* __restrict__ radii, int tiles_x, int tiles_y) {
    BoxOf b{0, 0, 1, 0, 0};
    if (i >= 0 && i < total) {
        const int r = radii[i];
        if (r > 0) {
            const float2 c = __ldg(reinterpret_cast<const float2*>(xys + i * xy_stride));
            const TileBox tb = tile_box(c.x, c.y, (float)r, tiles_x, tiles_y);
            const int view = (int)(i / n);
            b.x0 = tb.x0; b.y0 = tb.y0; b.w = max(tb.x1 - tb.x0, 1);
            b.cnt = max(tb.area(), 0);
            b.tile0 = view * tiles_x * tiles_y;
        }
    }
    return b;
}

// The three loads of a row, none depending on another (load_box reads the centre only when the radius is positive:
// two dependent DRAM latencies).  The shared-memory kernels below issue them one iteration ahead of their use.
struct RawRow {
    int r;
    float2 c;
    float d;
};
template <bool kDepth>
__device__ __forceinline__ RawRow load_raw(long long i, bool in_range, const float* __restrict__ xys, int xy_stride,
                                           const float* __restrict__ depths, const int32_t* __restrict__ radii) {
    RawRow w{0, make_float2(0.0f, 0.0f), 0.0f};
    if (in_range) {
        w.r = __ldg(radii + i);
        w.c = __ldg(reinterpret_cast<const float2*>(xys + i * xy_stride));
        if (kDepth) w.d = __ldg(depths + i);
    }
    return w;
}
// the box of a loaded row, tile ids local to the view
__device__ __forceinline__ BoxOf box_of(const RawRow& w, int tiles_x, int tiles_y) {
    BoxOf b{0, 0, 1, 0, 0};
    if (w.r > 0) {
        const TileBox tb = tile_box(w.c.x, w.c.y, (float)w.r, tiles_x, tiles_y);
        b.x0 = tb.x0; b.y0 = tb.y0; b.w = max(tb.x1 - tb.x0, 1);
        b.cnt = max(tb.area(), 0);
    }
    return b;
}

__global__ void __launch_bounds__(256)
tile_count_kernel(long long total, int n, const float* __restrict__ xys, int xy_stride,
                  const int32_t* __restrict__ radii, int tiles_x, int tiles_y, int* __restrict__ counts) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const BoxOf b = load_box(i, total, n, xys, xy_stride, radii, tiles_x, tiles_y);
    const bool big = b.cnt > kBigBox;
    if (!big) {
        for (int k = 0; k < b.cnt; ++k) {
            const int ty = b.y0 + k / b.w, tx = b.x0 + k % b.w;
            atomicAdd(counts + b.tile0 + ty * tiles_x + tx, 1);
        }
    }
    // a footprint of hundreds of tiles must not serialise on one thread: the warp shares it
    unsigned bigmask = __ballot_sync(0xffffffffu, big);
    while (bigmask) {
        const int src = __ffs(bigmask) - 1;
        bigmask &= bigmask - 1;
        const int x0 = __shfl_sync(0xffffffffu, b.x0, src), y0 = __shfl_sync(0xffffffffu, b.y0, src);
        const int w = __shfl_sync(0xffffffffu, b.w, src), cnt = __shfl_sync(0xffffffffu, b.cnt, src);
        const int tile0 = __shfl_sync(0xffffffffu, b.tile0, src);
        for (int k = lane; k < cnt; k += 32) {
            const int ty = y0 + k / w, tx = x0 + k % w;
            atomicAdd(counts + tile0 + ty * tiles_x + tx, 1);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Shared-memory variants of counting and placement.  grid = (blocks_per_view, views); CTA b of view v owns the
// Gaussians [b * per, (b+1) * per) of that view (per is a multiple of 32, identical in both kernels).
// ---------------------------------------------------------------------------------------------
constexpr int kHistThreads = 1024;
constexpr int kHistBlocksTotal = 296;            // 2 CTAs of 1024 threads per SM
constexpr int kHistMaxTiles = 40960;             // tiles per view whose counters fit a shared-memory histogram

// Every tile of every lane's box: GG_FOR_TILES(b, rec) { ... uses `tile` and `r` (the owner's record) ... }
// Small boxes are walked by their own lane; boxes of more than kBigBox tiles by the whole warp, 32 tiles per
// step (a footprint of hundreds of tiles must not serialise on one thread).  Must be reached by full warps.
#define GG_SMALL_BOX_LOOP(b, BODY)                                                          \
    if ((b).cnt <= kBigBox) {                                                               \
        for (int k_ = 0; k_ < (b).cnt; ++k_) {                                              \
            const int tile = (b).tile0 + ((b).y0 + k_ / (b).w) * tiles_x + (b).x0 + k_ % (b).w; \
            BODY                                                                            \
        }                                                                                   \
    }
#define GG_BIG_BOX_LOOP(b, rec, BODY)                                                       \
    for (unsigned bm_ = __ballot_sync(0xffffffffu, (b).cnt > kBigBox); bm_; bm_ &= bm_ - 1) { \
        const int src_ = __ffs(bm_) - 1;                                                    \
        const int x0_ = __shfl_sync(0xffffffffu, (b).x0, src_), y0_ = __shfl_sync(0xffffffffu, (b).y0, src_); \
        const int w_ = __shfl_sync(0xffffffffu, (b).w, src_), cnt_ = __shfl_sync(0xffffffffu, (b).cnt, src_); \
        const int t0_ = __shfl_sync(0xffffffffu, (b).tile0, src_);                          \
        const unsigned long long r = __shfl_sync(0xffffffffu, (rec), src_);                 \
        (void)r;                                                                            \
        for (int k_ = lane; k_ < cnt_; k_ += 32) {                                          \
            const int tile = t0_ + (y0_ + k_ / w_) * tiles_x + x0_ + k_ % w_;               \
            BODY                                                                            \
        }                                                                                   \
    }

__global__ void __launch_bounds__(kHistThreads)
tile_hist_kernel(int n, int per, const float* __restrict__ xys, int xy_stride, const int32_t* __restrict__ radii,
                 int tiles_x, int tiles_y, int* __restrict__ blockhist, unsigned* __restrict__ done) {
    extern __shared__ int sh_hist[];
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *done = 0u;  // ticket of tile_prefix_scan_kernel
    const int T = tiles_x * tiles_y;
    const int view = blockIdx.y, lane = threadIdx.x & 31;
    for (int t = threadIdx.x; t < T; t += kHistThreads) sh_hist[t] = 0;
    __syncthreads();
    const int lo = min(n, (int)blockIdx.x * per), hi = min(n, lo + per);
    const long long vbase = (long long)view * n;
    RawRow next = load_raw<false>(vbase + lo + threadIdx.x, lo + (int)threadIdx.x < hi, xys, xy_stride, nullptr, radii);
    for (int i0 = lo; i0 < hi; i0 += kHistThreads) {
        const RawRow cur = next;
        const int gn = i0 + kHistThreads + threadIdx.x;
        next = load_raw<false>(vbase + gn, gn < hi, xys, xy_stride, nullptr, radii);
        const BoxOf b = box_of(cur, tiles_x, tiles_y);
        GG_SMALL_BOX_LOOP(b, atomicAdd(&sh_hist[tile], 1);)
        GG_BIG_BOX_LOOP(b, 0ull, atomicAdd(&sh_hist[tile], 1);)
    }
    __syncthreads();
    int* row = blockhist + ((size_t)view * gridDim.x + blockIdx.x) * T;
    for (int t = threadIdx.x; t < T; t += kHistThreads) row[t] = sh_hist[t];
}

__global__ void __launch_bounds__(kHistThreads)
tile_place_kernel(int n, int per, const float* __restrict__ xys, int xy_stride, const float* __restrict__ depths,
                  const int32_t* __restrict__ radii, int tiles_x, int tiles_y, const int* __restrict__ blockhist,
                  const int* __restrict__ tile_start, const int32_t* __restrict__ info, long long capacity,
                  unsigned long long* __restrict__ pairs) {
    if (__ldg(info + 1)) return;  // over capacity: nothing is written, every range is (0,0)
    extern __shared__ int sh_cur[];
    const int T = tiles_x * tiles_y;
    const int view = blockIdx.y, lane = threadIdx.x & 31;
    const int* row = blockhist + ((size_t)view * gridDim.x + blockIdx.x) * T;
    // first slot of (tile, this CTA) = start of the tile + entries of the CTAs before this one
    for (int t = threadIdx.x; t < T; t += kHistThreads) sh_cur[t] = tile_start[(size_t)view * T + t] + row[t];
    __syncthreads();
    const int lo = min(n, (int)blockIdx.x * per), hi = min(n, lo + per);
    const long long vbase = (long long)view * n;
    RawRow next = load_raw<true>(vbase + lo + threadIdx.x, lo + (int)threadIdx.x < hi, xys, xy_stride, depths, radii);
    for (int i0 = lo; i0 < hi; i0 += kHistThreads) {
        const int g = i0 + threadIdx.x;
        const RawRow cur = next;
        const int gn = g + kHistThreads;
        next = load_raw<true>(vbase + gn, gn < hi, xys, xy_stride, depths, radii);
        const BoxOf b = box_of(cur, tiles_x, tiles_y);
        unsigned long long rec = 0ull;
        if (b.cnt > 0) rec = ((unsigned long long)__float_as_uint(cur.d) << 32) | (unsigned)g;
        GG_SMALL_BOX_LOOP(b, { const int pos = atomicAdd(&sh_cur[tile], 1); if (pos >= 0 && pos < capacity) pairs[pos] = rec; })
        GG_BIG_BOX_LOOP(b, rec, { const int pos = atomicAdd(&sh_cur[tile], 1); if (pos >= 0 && pos < capacity) pairs[pos] = r; })
    }
}

__device__ __forceinline__ int warp_incl_scan_i(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// One CTA of 1024 threads.  counts[t] in; out: tile_ranges, counts[t] <- start of tile t (the placement's cursor
// base), tile_order (tiles by descending length bucket), info = {M, overflow, longest, 0}, redo[t] <- 0.
__device__ __forceinline__ void scan_order_body(int num_tiles, long long capacity, int* counts, int32_t* tile_ranges,
                                                int32_t* tile_order, int32_t* info, int32_t* lists,
                                                int32_t* list_counts, int class_limit, int32_t* info_mapped) {
    __shared__ long long s_sum[32];
    __shared__ int s_max[32];
    __shared__ int s_wsum[32];
    __shared__ int bins[kOrderBins2];
    __shared__ int sm[kOrderBins2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    long long tot = 0;
    int mx = 0;
    for (int t = tid; t < num_tiles; t += 1024) {
        const int c = __ldcg(counts + t);   // written by other CTAs of the same launch (prefix kernel): bypass L1
        tot += c;
        mx = max(mx, c);
    }
    if (threadIdx.x == 0) { list_counts[0] = 0; list_counts[1] = 0; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        tot += __shfl_xor_sync(0xffffffffu, tot, o);
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) { s_sum[warp] = tot; s_max[warp] = mx; }
    bins[tid] = 0;
    __syncthreads();
    tot = 0; mx = 0;
#pragma unroll
    for (int w = 0; w < 32; ++w) { tot += s_sum[w]; mx = max(mx, s_max[w]); }
    const bool overflow = tot > capacity || tot >= (1ll << 31);
    if (tid == 0) {
        info[0] = (int32_t)min(tot, (long long)INT32_MAX);
        info[1] = overflow ? 1 : 0;
        info[2] = mx;
        info[3] = 0;
        if (info_mapped) {
            // the host's copy, stored straight into its pinned (device-mapped) record: no copy-engine hop on the
            // stream between the sort and the blend
            *reinterpret_cast<int4*>(info_mapped) = make_int4((int)min(tot, (long long)INT32_MAX), overflow ? 1 : 0, mx, 0);
        }
    }
    if (overflow) {
        for (int t = tid; t < num_tiles; t += 1024) {
            tile_ranges[2 * t] = 0; tile_ranges[2 * t + 1] = 0;
            tile_order[t] = t;
            counts[t] = 0;
        }
        return;
    }
    int carry = 0;
    for (int base = 0; base < num_tiles; base += 1024) {
        const int t = base + tid;
        const int c = t < num_tiles ? __ldcg(counts + t) : 0;
        const int incl = warp_incl_scan_i(c, lane);
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        int woff = 0, btot = 0;
#pragma unroll
        for (int w = 0; w < 32; ++w) {
            const int v = s_wsum[w];
            if (w < warp) woff += v;
            btot += v;
        }
        const int start = carry + woff + incl - c;
        if (t < num_tiles) {
            // empty tiles keep (0,0), as the reference's get_tile_bin_edges leaves them
            tile_ranges[2 * t] = c ? start : 0;
            tile_ranges[2 * t + 1] = c ? start + c : 0;
            counts[t] = start;
            atomicAdd(&bins[len_bucket(c)], 1);
            if (c > class_limit) {  // longer than the bucket-sort classes launched by this call
                const int which = c <= kSortMaxSmem ? 0 : 1;
                lists[which * num_tiles + atomicAdd(list_counts + which, 1)] = t;
            }
        }
        carry += btot;
        __syncthreads();
    }
    // bins[b] <- number of tiles in buckets above b (descending exclusive scan)
    sm[tid] = bins[kOrderBins2 - 1 - tid];
    __syncthreads();
    for (int o = 1; o < kOrderBins2; o <<= 1) {
        const int v = tid >= o ? sm[tid - o] : 0;
        __syncthreads();
        sm[tid] += v;
        __syncthreads();
    }
    const int start = sm[tid] - bins[kOrderBins2 - 1 - tid];
    __syncthreads();
    bins[kOrderBins2 - 1 - tid] = start;
    __syncthreads();
    for (int t = tid; t < num_tiles; t += 1024) {
        const int len = tile_ranges[2 * t + 1] - tile_ranges[2 * t];
        tile_order[atomicAdd(&bins[len_bucket(len)], 1)] = t;
    }
}

__global__ void __launch_bounds__(1024)
tile_scan_order_kernel(int num_tiles, long long capacity, int* counts, int32_t* tile_ranges, int32_t* tile_order,
                       int32_t* info, int32_t* lists, int32_t* list_counts, int class_limit, int32_t* info_mapped) {
    scan_order_body(num_tiles, capacity, counts, tile_ranges, tile_order, info, lists, list_counts, class_limit, info_mapped);
}

// Exclusive prefix over the rows of a view's counting CTAs, per tile: CTA = 32 tiles x 32 row groups (thread
// (x, y) sums rows y*rpg .. of tile x, the groups are scanned through shared memory, then every thread rewrites
// its rows as running offsets); total -> counts.  The CTA that finishes last runs the scan over the tiles (the
// rows and counts of the others are complete by then).
__global__ void __launch_bounds__(1024)
tile_prefix_scan_kernel(int num_tiles, int tiles_per_view, int blocks_per_view, long long capacity, int* blockhist,
                        int* counts, int32_t* tile_ranges, int32_t* tile_order, int32_t* info, int32_t* lists,
                        int32_t* list_counts, int class_limit, unsigned* done, int32_t* info_mapped) {
    __shared__ bool s_last;
    __shared__ int s_grp[32][33];
    constexpr int kMaxRows = 10;                      // rows per thread: 32 * 10 >= kHistBlocksTotal
    const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    const int gt = blockIdx.x * 32 + x;
    const int rpg = (blocks_per_view + 31) >> 5;      // rows per group (<= kMaxRows)
    const bool live = gt < num_tiles;
    const int view = live ? gt / tiles_per_view : 0, t = live ? gt - view * tiles_per_view : 0;
    int* p = blockhist + ((size_t)view * blocks_per_view + (size_t)y * rpg) * tiles_per_view + t;
    int c[kMaxRows];
    int sum = 0;
#pragma unroll
    for (int u = 0; u < kMaxRows; ++u) {
        c[u] = (live && u < rpg && y * rpg + u < blocks_per_view) ? p[(size_t)u * tiles_per_view] : 0;
        sum += c[u];
    }
    s_grp[y][x] = sum;
    __syncthreads();
    int run = 0, total = 0;
#pragma unroll
    for (int g = 0; g < 32; ++g) {
        const int v = s_grp[g][x];
        if (g < y) run += v;
        total += v;
    }
#pragma unroll
    for (int u = 0; u < kMaxRows; ++u) {
        if (live && u < rpg && y * rpg + u < blocks_per_view) p[(size_t)u * tiles_per_view] = run;
        run += c[u];
    }
    if (live && y == 0) counts[gt] = total;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x == 0) *done = 0u;  // armed for the next call
    __threadfence();
    scan_order_body(num_tiles, capacity, counts, tile_ranges, tile_order, info, lists, list_counts, class_limit, info_mapped);
}

__global__ void __launch_bounds__(256)
tile_scatter_kernel(long long total, int n, const float* __restrict__ xys, int xy_stride,
                    const float* __restrict__ depths, const int32_t* __restrict__ radii, int tiles_x, int tiles_y,
                    int* __restrict__ cursor, const int32_t* __restrict__ info, long long capacity,
                    unsigned long long* __restrict__ pairs) {
    if (__ldg(info + 1)) return;  // over capacity: nothing is written, every range is (0,0)
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const BoxOf b = load_box(i, total, n, xys, xy_stride, radii, tiles_x, tiles_y);
    unsigned long long rec = 0ull;
    if (b.cnt > 0) {
        const int view = (int)(i / n);
        const unsigned g = (unsigned)(i - (long long)view * n);
        rec = ((unsigned long long)__float_as_uint(depths[i]) << 32) | g;
    }
    const bool big = b.cnt > kBigBox;
    if (!big) {
        // slots of four entries are requested before the first record is stored: the atomics' round
        // trips overlap
        for (int k0 = 0; k0 < b.cnt; k0 += 4) {
            int pos[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int k = k0 + u;
                pos[u] = -1;
                if (k < b.cnt) {
                    const int ty = b.y0 + k / b.w, tx = b.x0 + k % b.w;
                    pos[u] = atomicAdd(cursor + b.tile0 + ty * tiles_x + tx, 1);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (pos[u] >= 0 && pos[u] < capacity) pairs[pos[u]] = rec;
        }
    }
    unsigned bigmask = __ballot_sync(0xffffffffu, big);
    while (bigmask) {
        const int src = __ffs(bigmask) - 1;
        bigmask &= bigmask - 1;
        const int x0 = __shfl_sync(0xffffffffu, b.x0, src), y0 = __shfl_sync(0xffffffffu, b.y0, src);
        const int w = __shfl_sync(0xffffffffu, b.w, src), cnt = __shfl_sync(0xffffffffu, b.cnt, src);
        const int tile0 = __shfl_sync(0xffffffffu, b.tile0, src);
        const unsigned long long r = __shfl_sync(0xffffffffu, rec, src);
        for (int k = lane; k < cnt; k += 32) {
            const int ty = y0 + k / w, tx = x0 + k % w;
            const int pos = atomicAdd(cursor + tile0 + ty * tiles_x + tx, 1);
            if (pos >= 0 && pos < capacity) pairs[pos] = r;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Per-tile sort in shared memory.  THREADS x IPT >= segment length.  Keys (depth bits) and values (ids) stay
// in registers between passes; a pass ranks the keys of each warp with one ballot per digit bit (stable in
// (warp, slot, lane) = index order), scans the per-warp digit counts and permutes through shared memory.
// ---------------------------------------------------------------------------------------------
struct SortArgs {
    const int32_t* tile_ranges;
    const int32_t* tile_order;
    unsigned long long* pairs;
    int32_t* ids_sorted;
    int num_tiles;
    // tiles the bucket sort does not finish: lists[0 .. num_tiles) for the radix kernel, lists[num_tiles .. 2 num_tiles)
    // for the bitonic kernel; list_counts[0], [1] = their lengths (reset by the scan, which also lists the tiles
    // that fit no shared-memory class)
    int32_t* lists;
    int32_t* list_counts;
};

// ---------------------------------------------------------------------------------------------
// Bucket sort of one tile (fast path).  NB = 2 * THREADS buckets over [min key, max key] of the tile, linear in
// the depth bits (a monotone map: float conversion, product with a positive scale and truncation all preserve
// <=), counting sort into shared memory with atomics (the order inside a bucket is arbitrary), then rank by
// comparison inside the bucket.  Work per entry ~ its bucket's population; when sum(pop^2) exceeds
// kRankCost * len (depths piled up in few buckets) the tile is flagged in redo[] and left to the radix kernel.
// ---------------------------------------------------------------------------------------------
constexpr int kRankCost = 64;

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
tile_bucket_sort_kernel(const SortArgs a, int len_lo, int len_hi) {
    constexpr int NB = 2 * THREADS;
    constexpr int NW = THREADS / 32;
    extern __shared__ __align__(16) unsigned char dyn_b[];
    unsigned long long* buf = reinterpret_cast<unsigned long long*>(dyn_b);   // [len_hi]
    int* bstart = reinterpret_cast<int*>(buf + len_hi);                        // [NB + 1]
    int* cursor = bstart + NB + 1;                                             // [NB]
    __shared__ uint32_t s_min[32], s_max[32];
    __shared__ int s_wsum[32];
    __shared__ float s_cost[32];

    const int tile = __ldg(a.tile_order + blockIdx.x);
    const int2 range = __ldg(reinterpret_cast<const int2*>(a.tile_ranges) + tile);
    const int len = range.y - range.x;
    if (len <= len_lo || len > len_hi) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long* src = a.pairs + range.x;

    uint32_t kmin = ~0u, kmax = 0u;
    for (int e = tid; e < len; e += THREADS) {
        const uint32_t k = (uint32_t)(src[e] >> 32);
        kmin = min(kmin, k); kmax = max(kmax, k);
    }
    kmin = __reduce_min_sync(0xffffffffu, kmin);
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    if (lane == 0) { s_min[warp] = kmin; s_max[warp] = kmax; }
    for (int d = tid; d < NB; d += THREADS) cursor[d] = 0;
    __syncthreads();
    kmin = __reduce_min_sync(0xffffffffu, lane < NW ? s_min[lane] : ~0u);
    kmax = __reduce_max_sync(0xffffffffu, lane < NW ? s_max[lane] : 0u);
    const float scale = (float)NB / ((float)(kmax - kmin) + 1.0f);
    auto bucket = [&](uint32_t k) { return min(NB - 1, (int)((float)(k - kmin) * scale)); };

    for (int e = tid; e < len; e += THREADS) atomicAdd(&cursor[bucket((uint32_t)(src[e] >> 32))], 1);
    __syncthreads();
    // exclusive scan of the populations: two adjacent buckets per thread
    const int c0 = cursor[2 * tid], c1 = cursor[2 * tid + 1];
    const int incl = warp_incl_scan_i(c0 + c1, lane);
    float cost = (float)c0 * (float)c0 + (float)c1 * (float)c1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cost += __shfl_xor_sync(0xffffffffu, cost, o);
    if (lane == 31) s_wsum[warp] = incl;
    if (lane == 0) s_cost[warp] = cost;
    __syncthreads();
    int woff = 0;
    cost = 0.0f;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        if (w < warp) woff += s_wsum[w];
        cost += s_cost[w];
    }
    if (cost > (float)kRankCost * (float)len) {  // block-uniform
        if (tid == 0) {
            const int which = len <= kSortMaxSmem ? 0 : 1;
            a.lists[which * a.num_tiles + atomicAdd(a.list_counts + which, 1)] = tile;
        }
        return;
    }
    const int ex = woff + incl - (c0 + c1);
    bstart[2 * tid] = ex; bstart[2 * tid + 1] = ex + c0;
    cursor[2 * tid] = ex; cursor[2 * tid + 1] = ex + c0;
    if (tid == THREADS - 1) bstart[NB] = ex + c0 + c1;
    __syncthreads();
    for (int e = tid; e < len; e += THREADS) {
        const unsigned long long p = src[e];
        buf[atomicAdd(&cursor[bucket((uint32_t)(p >> 32))], 1)] = p;
    }
    __syncthreads();
    int32_t* out = a.ids_sorted + range.x;
    for (int e = tid; e < len; e += THREADS) {
        const unsigned long long mine = buf[e];
        const int d = bucket((uint32_t)(mine >> 32));
        const int lo = bstart[d], hi = bstart[d + 1];
        int below = 0;
        for (int q = lo; q < hi; ++q) below += buf[q] < mine;
        out[lo + below] = (int32_t)(uint32_t)mine;
    }
}

template <int THREADS, int IPT>
constexpr size_t tile_sort_smem() {
    return sizeof(uint32_t) * (2 * THREADS * IPT + (THREADS / 32) * 256 + 256);
}

template <int THREADS, int IPT>
__device__ __forceinline__ void radix_sort_tiles(const SortArgs& a) {
    constexpr int CAP = THREADS * IPT;
    constexpr int NW = THREADS / 32;
    static_assert(IPT % 2 == 0 && 32 * IPT < 65536, "ranks are packed in 16 bits");
    extern __shared__ __align__(16) unsigned char dyn[];
    uint32_t* skeys = reinterpret_cast<uint32_t*>(dyn);  // [CAP]
    uint32_t* svals = skeys + CAP;                       // [CAP]
    uint32_t* whist = svals + CAP;                       // [NW][256]
    uint32_t* dbase = whist + NW * 256;                  // [256]
    __shared__ uint32_t s_red[4][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wbase = warp * 32 * IPT;
    const int n_todo = __ldcg(a.list_counts);
  for (int it = blockIdx.x; it < n_todo; it += gridDim.x) {
    __syncthreads();  // the previous tile of this CTA is done with shared memory
    const int tile = __ldcg(a.lists + it);
    const int2 range = __ldg(reinterpret_cast<const int2*>(a.tile_ranges) + tile);
    const int len = range.y - range.x;
    if (len <= 0 || len > CAP) continue;
    const unsigned long long* src = a.pairs + range.x;

    uint32_t key[IPT], val[IPT];
    uint32_t kor = 0u, kand = ~0u, vor = 0u, vand = ~0u;
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
        const int idx = wbase + j * 32 + lane;
        key[j] = ~0u; val[j] = ~0u;   // padding: never ranked, never placed
        if (idx < len) {
            const unsigned long long p = src[idx];
            key[j] = (uint32_t)(p >> 32); val[j] = (uint32_t)p;
            kor |= key[j]; kand &= key[j]; vor |= val[j]; vand &= val[j];
        }
    }
    // bits on which the tile's keys / ids differ: only those are sorted
    kor = __reduce_or_sync(0xffffffffu, kor); kand = __reduce_and_sync(0xffffffffu, kand);
    vor = __reduce_or_sync(0xffffffffu, vor); vand = __reduce_and_sync(0xffffffffu, vand);
    if (lane == 0) { s_red[0][warp] = kor; s_red[1][warp] = kand; s_red[2][warp] = vor; s_red[3][warp] = vand; }
    __syncthreads();
    {
        const bool in = lane < NW;
        kor = __reduce_or_sync(0xffffffffu, in ? s_red[0][lane] : 0u);
        kand = __reduce_and_sync(0xffffffffu, in ? s_red[1][lane] : ~0u);
        vor = __reduce_or_sync(0xffffffffu, in ? s_red[2][lane] : 0u);
        vand = __reduce_and_sync(0xffffffffu, in ? s_red[3][lane] : ~0u);
    }
    const int khi = 32 - __clz(kor ^ kand);  // 0: all depths equal
    const int vhi = 32 - __clz(vor ^ vand);

    bool in_smem = false;  // shared memory holds the current order (registers may be stale)
    // phase 0: depth bits.  phase 1 (only if two entries share a depth): id bits, then phase 2: depth bits again
    for (int phase = 0; phase < 3; ++phase) {
        if (phase == 1) {
            if (!in_smem) {  // no depth bit varies: stage the loaded order
#pragma unroll
                for (int j = 0; j < IPT; ++j) {
                    const int idx = wbase + j * 32 + lane;
                    skeys[idx] = key[j]; svals[idx] = val[j];
                }
                in_smem = true;
            }
            __syncthreads();
            bool tie = false;
            for (int e = tid; e + 1 < len; e += THREADS) tie |= skeys[e] == skeys[e + 1];
            if (!__syncthreads_or(tie)) break;
        }
        const bool by_val = phase == 1;
        const int hi = by_val ? vhi : khi;
        const int np = (hi + 7) >> 3;
        int shift = 0;
        for (int p = 0; p < np; ++p) {
            const int bits = hi / np + (p < hi % np ? 1 : 0);
            const uint32_t mask = (1u << bits) - 1u;
            const int nb = 1 << bits;
            if (in_smem) {  // registers <- current order
#pragma unroll
                for (int j = 0; j < IPT; ++j) {
                    const int idx = wbase + j * 32 + lane;
                    key[j] = skeys[idx]; val[j] = svals[idx];
                }
            }
            for (int k = tid; k < NW * 256; k += THREADS) whist[k] = 0u;
            __syncthreads();
            uint32_t rk[IPT / 2];
#pragma unroll
            for (int j = 0; j < IPT; ++j) {
                if (wbase + j * 32 >= len) continue;  // warp-uniform: nothing but padding in this slot
                const bool valid = wbase + j * 32 + lane < len;
                const uint32_t d = ((by_val ? val[j] : key[j]) >> shift) & mask;
                uint32_t peers = __ballot_sync(0xffffffffu, valid);  // padding is neither ranked nor placed
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    if (b < bits) {  // warp-uniform
                        const bool bit = (d >> b) & 1u;
                        const uint32_t v = __ballot_sync(0xffffffffu, bit);
                        peers &= bit ? v : ~v;
                    }
                }
                const int leader = __ffs(peers) - 1;
                uint32_t prev = 0u;
                if (valid && lane == leader) {
                    prev = whist[warp * 256 + d];
                    whist[warp * 256 + d] = prev + (uint32_t)__popc(peers);
                }
                prev = __shfl_sync(0xffffffffu, prev, leader);
                const uint32_t r = prev + (uint32_t)__popc(peers & ((1u << lane) - 1u));
                if (j & 1) rk[j >> 1] |= r << 16; else rk[j >> 1] = r;
                __syncwarp();
            }
            __syncthreads();
            // digit d: exclusive offsets of the warps, total of the tile
            for (int d = tid; d < nb; d += THREADS) {
                uint32_t run = 0u;
#pragma unroll 4
                for (int w = 0; w < NW; ++w) {
                    const uint32_t t = whist[w * 256 + d];
                    whist[w * 256 + d] = run;
                    run += t;
                }
                dbase[d] = run;
            }
            __syncthreads();
            if (warp == 0) {  // exclusive scan of the (up to 256) digit totals: 8 per lane
                uint32_t loc[8];
                uint32_t s = 0u;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int d = lane * 8 + q;
                    loc[q] = d < nb ? dbase[d] : 0u;
                    s += loc[q];
                }
                uint32_t ex = (uint32_t)warp_incl_scan_i((int)s, lane) - s;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int d = lane * 8 + q;
                    if (d < nb) dbase[d] = ex;
                    ex += loc[q];
                }
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < IPT; ++j) {
                if (wbase + j * 32 + lane >= len) continue;
                const uint32_t d = ((by_val ? val[j] : key[j]) >> shift) & mask;
                const uint32_t r = (j & 1) ? (rk[j >> 1] >> 16) : (rk[j >> 1] & 0xffffu);
                const uint32_t pos = dbase[d] + whist[warp * 256 + d] + r;
                skeys[pos] = key[j];
                svals[pos] = val[j];
            }
            in_smem = true;
            __syncthreads();
            shift += bits;
        }
    }
    if (!in_smem) {  // single entry or nothing to do: registers are the answer
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
            const int idx = wbase + j * 32 + lane;
            if (idx < len) a.ids_sorted[range.x + idx] = (int32_t)val[j];
        }
    } else {
        for (int e = tid; e < len; e += THREADS) a.ids_sorted[range.x + e] = (int32_t)svals[e];
    }
  }
}

// Segments longer than the shared-memory classes: bitonic network over the 64-bit records (depth << 32 | id
// order = the wanted order) in global memory, all comparators ascending (the "flip" form), so that virtual
// +inf padding beyond the segment never moves and the length need not be a power of two.
__device__ __forceinline__ void bitonic_sort_tiles(const SortArgs& a) {
    const int n_todo = __ldcg(a.list_counts + 1);
  for (int it = blockIdx.x; it < n_todo; it += gridDim.x) {
    const int tile = __ldcg(a.lists + a.num_tiles + it);
    const int2 range = __ldg(reinterpret_cast<const int2*>(a.tile_ranges) + tile);
    const int len = range.y - range.x;
    if (len <= 0) continue;
    unsigned long long* seg = a.pairs + range.x;
    long long np2 = 1;
    while (np2 < len) np2 <<= 1;
    const long long half = np2 >> 1;
    for (long long k = 2; k <= np2; k <<= 1) {
        const long long hk = k >> 1;
        for (long long i = threadIdx.x; i < half; i += 1024) {
            const long long blk = i / hk, off = i - blk * hk;
            const long long lo = blk * k + off, hi = blk * k + (k - 1 - off);
            if (hi < len) {
                const unsigned long long x = seg[lo], y = seg[hi];
                if (x > y) { seg[lo] = y; seg[hi] = x; }
            }
        }
        __syncthreads();
        for (long long j = k >> 2; j >= 1; j >>= 1) {
            for (long long i = threadIdx.x; i < half; i += 1024) {
                const long long lo = 2 * j * (i / j) + (i % j), hi = lo + j;
                if (hi < len) {
                    const unsigned long long x = seg[lo], y = seg[hi];
                    if (x > y) { seg[lo] = y; seg[hi] = x; }
                }
            }
            __syncthreads();
        }
    }
    for (int e = threadIdx.x; e < len; e += 1024) a.ids_sorted[range.x + e] = (int32_t)(uint32_t)seg[e];
  }
}

// The tiles the bucket sort left behind (normally none): radix list, then bitonic list, persistent CTAs.
__global__ void __launch_bounds__(1024, 1)
tile_sort_fallback_kernel(const SortArgs a) {
    radix_sort_tiles<1024, 16>(a);
    __syncthreads();
    bitonic_sort_tiles(a);
}

struct Bin2Layout {
    size_t info, done, list_counts, counts, lists, blockhist, pairs, bytes;
};

// CTAs per view and Gaussians per CTA of the shared-memory counting / placement kernels
static void hist_blocks(int n, int n_views, long long tiles_per_view, int& blocks, int& per) {
    // every placing CTA keeps one partially written 128-byte line open per tile: beyond ~half the L2 of open lines
    // the records are evicted half-filled and fetched again (config 2: 2.2 GB written for 1.1 GB of records), so
    // large tile grids run one CTA per SM instead of two (measured 33.9 -> 32.4 ms of binning per step there)
    const int total_blocks = (long long)kHistBlocksTotal * tiles_per_view * 128 > (64ll << 20) ? kHistBlocksTotal / 2
                                                                                                : kHistBlocksTotal;
    int b = total_blocks / n_views;
    if (b < 1) b = 1;
    const int by_work = (n + kHistThreads - 1) / kHistThreads;
    if (b > by_work) b = by_work;
    per = (((n + b - 1) / b) + 31) & ~31;
    blocks = (n + per - 1) / per;
}

static Bin2Layout bin2_layout(int n_views, long long tiles_per_view, long long capacity) {
    Bin2Layout L;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 255) & ~(size_t)255; return r; };
    const size_t num_tiles = (size_t)(n_views > 0 ? n_views : 1) * (size_t)(tiles_per_view > 0 ? tiles_per_view : 1);
    L.info = take(sizeof(int32_t) * 4);
    L.done = take(sizeof(unsigned));
    L.list_counts = take(sizeof(int32_t) * 2);
    L.counts = take(sizeof(int32_t) * num_tiles);
    L.lists = take(sizeof(int32_t) * 2 * num_tiles);
    // one row of counters per counting CTA: at most kHistBlocksTotal rows over all views (or one per view)
    const size_t rows = (size_t)(n_views > kHistBlocksTotal ? n_views : kHistBlocksTotal);
    L.blockhist = take(tiles_per_view <= kHistMaxTiles ? sizeof(int32_t) * rows * (size_t)tiles_per_view : 4);
    L.pairs = take(sizeof(unsigned long long) * (size_t)(capacity > 0 ? capacity : 1));
    L.bytes = o;
    return L;
}

}  // namespace gg

using namespace gg;

extern "C" size_t gg_bin_tiles_scratch_bytes(int n_views, long long tiles_per_view, long long capacity) {
    return bin2_layout(n_views, tiles_per_view, capacity).bytes;
}

extern "C" int gg_bin_tiles(int n, int n_views, const float* xys, int xy_stride, const float* depths,
                            const int32_t* radii, int tiles_x, int tiles_y, long long capacity, void* scratch,
                            size_t scratch_bytes, int32_t* ids_sorted, int32_t* tile_ranges, int32_t* tile_order,
                            int32_t* info_dev, int32_t* info_host, int longest_hint, void* stream) {
    GG_REQUIRE(n >= 1 && n_views >= 1 && (long long)n * n_views < (1ll << 31), "gg_bin_tiles: bad sizes");
    GG_REQUIRE(tiles_x >= 1 && tiles_y >= 1 && (long long)n_views * tiles_x * tiles_y < (1ll << 30),
               "gg_bin_tiles: too many tiles");
    GG_REQUIRE(capacity >= 0 && capacity < (1ll << 31), "gg_bin_tiles: capacity out of range");
    GG_REQUIRE(xys && depths && radii && scratch && tile_ranges && tile_order && (ids_sorted || capacity == 0),
               "gg_bin_tiles: null pointer");
    GG_REQUIRE((xy_stride == 2 || xy_stride == 8) && ((uintptr_t)xys & 7) == 0, "gg_bin_tiles: bad xys");
    GG_REQUIRE(((uintptr_t)scratch & 255) == 0 && ((uintptr_t)tile_ranges & 7) == 0, "gg_bin_tiles: misaligned");
    const long long total = (long long)n * n_views;
    const int T = tiles_x * tiles_y;
    const int num_tiles = n_views * T;
    const Bin2Layout L = bin2_layout(n_views, T, capacity);
    GG_REQUIRE(scratch_bytes >= L.bytes, "gg_bin_tiles: scratch too small");
    unsigned char* s = reinterpret_cast<unsigned char*>(scratch);
    int* counts = reinterpret_cast<int*>(s + L.counts);
    int32_t* info = info_dev ? info_dev : reinterpret_cast<int32_t*>(s + L.info);
    unsigned* done = reinterpret_cast<unsigned*>(s + L.done);
    int32_t* list_counts = reinterpret_cast<int32_t*>(s + L.list_counts);
    int32_t* lists = reinterpret_cast<int32_t*>(s + L.lists);
    int* blockhist = reinterpret_cast<int*>(s + L.blockhist);
    unsigned long long* pairs = reinterpret_cast<unsigned long long*>(s + L.pairs);
    cudaStream_t st = (cudaStream_t)stream;
    const bool smem_path = T <= kHistMaxTiles;
    // pinned host memory is mapped into the device's address space (UVA): the scan kernel stores the record there
    // itself.  Anything else (pageable memory) gets the stream-ordered copy at the end.
    int32_t* info_mapped = nullptr;
    if (info_host && ((uintptr_t)info_host & 15) == 0) {
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, info_host) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer)
            info_mapped = reinterpret_cast<int32_t*>(at.devicePointer);
        else
            (void)cudaGetLastError();
    }
    int blocks = 1, per = n;
    hist_blocks(n, n_views, T, blocks, per);
    const size_t hist_smem = sizeof(int) * (size_t)T;
    int launches = 0;
    // the 1024-thread bucket class (8193 .. 24576 entries) is launched only when such tiles are expected; a tile
    // beyond the launched classes is listed for the fallback kernel by the scan, so the hint never costs exactness
    const bool long_class = longest_hint <= 0 || longest_hint > 6144;
    const int class_limit = long_class ? kBucketMax : 8192;
    if (smem_path) {
        GG_CUDA(cudaFuncSetAttribute(tile_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_smem));
        tile_hist_kernel<<<dim3(blocks, n_views), kHistThreads, hist_smem, st>>>(n, per, xys, xy_stride, radii, tiles_x,
                                                                                 tiles_y, blockhist, done);
        tile_prefix_scan_kernel<<<div_up(num_tiles, 32), 1024, 0, st>>>(num_tiles, T, blocks, capacity, blockhist,
                                                                          counts, tile_ranges, tile_order, info, lists,
                                                                          list_counts, class_limit, done, info_mapped);
        launches += 2;
    } else {
        GG_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * (size_t)num_tiles, st));
        tile_count_kernel<<<div_up(total, 256), 256, 0, st>>>(total, n, xys, xy_stride, radii, tiles_x, tiles_y, counts);
        tile_scan_order_kernel<<<1, 1024, 0, st>>>(num_tiles, capacity, counts, tile_ranges, tile_order, info, lists,
                                                   list_counts, class_limit, info_mapped);
        launches += 2;
    }
    if (capacity > 0) {
        if (smem_path) {
            GG_CUDA(cudaFuncSetAttribute(tile_place_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_smem));
            tile_place_kernel<<<dim3(blocks, n_views), kHistThreads, hist_smem, st>>>(
                n, per, xys, xy_stride, depths, radii, tiles_x, tiles_y, blockhist, counts, info, capacity, pairs);
        } else {
            tile_scatter_kernel<<<div_up(total, 256), 256, 0, st>>>(total, n, xys, xy_stride, depths, radii, tiles_x,
                                                                    tiles_y, counts, info, capacity, pairs);
        }
        SortArgs a{tile_ranges, tile_order, pairs, ids_sorted, num_tiles, lists, list_counts};
        // tile_order lists the tiles by descending length bucket (8 entries wide, saturating at 8184), so the
        // tiles of a class sit among the first capacity / (class minimum) positions
        auto grid_for = [&](long long min_len) {
            const long long g = capacity / min_len + 1;
            return (unsigned)(g < num_tiles ? g : num_tiles);
        };
        auto bucket_smem = [](int threads, int cap) { return sizeof(unsigned long long) * (size_t)cap + sizeof(int) * (4 * (size_t)threads + 1); };
        {
            const size_t sm = bucket_smem(256, 2048);
            GG_CUDA(cudaFuncSetAttribute(tile_bucket_sort_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)bucket_smem(256, 8192)));
            tile_bucket_sort_kernel<256><<<(unsigned)num_tiles, 256, sm, st>>>(a, 0, 2048);
            tile_bucket_sort_kernel<256><<<grid_for(2048), 256, bucket_smem(256, 8192), st>>>(a, 2048, 8192);
        }
        if (long_class) {
            const size_t sm = bucket_smem(1024, kBucketMax);
            GG_CUDA(cudaFuncSetAttribute(tile_bucket_sort_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            tile_bucket_sort_kernel<1024><<<grid_for(8184), 1024, sm, st>>>(a, 8192, kBucketMax);
            ++launches;
        }
        {
            constexpr size_t sm = tile_sort_smem<1024, 16>();
            GG_CUDA(cudaFuncSetAttribute(tile_sort_fallback_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            const unsigned g = (unsigned)(num_tiles < 148 ? num_tiles : 148);
            tile_sort_fallback_kernel<<<g, 1024, sm, st>>>(a);
        }
        launches += 4;
    }
    count_launch(launches);
    if (info_host && !info_mapped)
        GG_CUDA(cudaMemcpyAsync(info_host, info, sizeof(int32_t) * 4, cudaMemcpyDeviceToHost, st));
    return check_launch("gg_bin_tiles");
}
