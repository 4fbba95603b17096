// Tile-first binning without a host read: count -> scan -> scatter -> per-tile sort.
//
// Same result as gsplat 0.1.0's compute_cumulative_intersects + bin_and_sort_gaussians + get_tile_bin_edges
// (re-run inside every rasterize forward of the reference, nerfstudio/models/gaussian_splatting.py:735,747,759,
// 773): the Gaussian ids of every tile in (depth bits, id) order and the per-tile [start, end) ranges -- bit for
// bit what sorting the 64-bit tile|depth keys with ties broken by id gives.  The route differs:
//
//   1. tile_count_kernel       every visible (view, Gaussian) adds 1 to the counter of each tile of its bbox
//   2. tile_scan_order_kernel  one CTA: exclusive scan of the V*T counters -> tile_ranges, per-tile write
//                              cursors, M and the capacity check (on the device), longest-first tile order
//   3. tile_scatter_kernel     every entry takes a slot of its tile's segment (atomic cursor) and stores
//                              (depth bits << 32 | id); the order inside a segment is arbitrary at this point
//   4. tile_sort_kernel        one CTA per tile sorts its segment in shared memory: LSD radix on the depth
//                              bits that actually vary inside the tile; equal depths are detected afterwards
//                              and, only then, the tile is re-sorted on (id, depth) -- the final order never
//                              depends on the scatter order.  Segments beyond 16384 entries take a bitonic
//                              network in global memory (tile_sort_big_kernel).
//
// Nothing here needs the intersection count M on the host: buffers are sized by a caller-chosen capacity and
// info[] = {M, overflow, longest tile, 0} is written on the device (and copied to pinned host memory if asked).
// When M exceeds the capacity every tile range is (0,0) -- the blend kernels then render background only --
// and the overflow flag tells the caller to grow the buffers and repeat.
//
// Traffic: 16 B per (view, Gaussian) read twice, 8 B per entry written + read, 4 B per entry written --
// against ~70 B per entry for the 2-pass global radix sort of the depth-first path plus its 4-5 N-sized passes;
// and 6 kernel launches instead of ~20.
//
// Compiled with -fmad=false (tile_box() must reproduce the projection kernel's tile bbox).
#include "gg_common.cuh"
#include "gg_math.cuh"
#include "gg_b200.h"

namespace gg {

constexpr int kOrderBins2 = 1024;
constexpr int kBigBox = 16;       // bboxes with more tiles than this are spread over the warp
constexpr int kSortMaxSmem = 16384;  // longest segment sorted in shared memory

__device__ __forceinline__ int len_bucket(int len) { return min(kOrderBins2 - 1, len >> 3); }

struct BoxOf {
    int x0, y0, w, cnt, tile0;
};

// bbox of row i (= view * n + g) or an empty box
__device__ __forceinline__ BoxOf load_box(long long i, long long total, int n, const float* __restrict__ xys, int xy_stride,
                                          const int32_t* __restrict__ radii, int tiles_x, int tiles_y) {
    BoxOf b{0, 0, 1, 0, 0};
    if (i < total) {
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

__device__ __forceinline__ int warp_incl_scan_i(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// One CTA of 1024 threads.  counts[t] in; out: tile_ranges, counts[t] <- start of tile t (the scatter's cursor),
// tile_order (tiles by descending length bucket), info = {M, overflow, longest, 0}.
__global__ void __launch_bounds__(1024)
tile_scan_order_kernel(int num_tiles, long long capacity, int* __restrict__ counts, int32_t* __restrict__ tile_ranges,
                       int32_t* __restrict__ tile_order, int32_t* __restrict__ info) {
    __shared__ long long s_sum[32];
    __shared__ int s_max[32];
    __shared__ int s_wsum[32];
    __shared__ int bins[kOrderBins2];
    __shared__ int sm[kOrderBins2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    long long tot = 0;
    int mx = 0;
    for (int t = tid; t < num_tiles; t += 1024) {
        const int c = counts[t];
        tot += c;
        mx = max(mx, c);
    }
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
        const int c = t < num_tiles ? counts[t] : 0;
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
};

template <int THREADS, int IPT>
constexpr size_t tile_sort_smem() {
    return sizeof(uint32_t) * (2 * THREADS * IPT + (THREADS / 32) * 256 + 256);
}

template <int THREADS, int IPT>
__global__ void __launch_bounds__(THREADS, (THREADS >= 1024) ? 1 : ((THREADS >= 256) ? 3 : 8))
tile_sort_kernel(const SortArgs a, int len_lo, int len_hi) {
    constexpr int CAP = THREADS * IPT;
    constexpr int NW = THREADS / 32;
    static_assert(IPT % 2 == 0 && 32 * IPT < 65536, "ranks are packed in 16 bits");
    extern __shared__ __align__(16) unsigned char dyn[];
    uint32_t* skeys = reinterpret_cast<uint32_t*>(dyn);  // [CAP]
    uint32_t* svals = skeys + CAP;                       // [CAP]
    uint32_t* whist = svals + CAP;                       // [NW][256]
    uint32_t* dbase = whist + NW * 256;                  // [256]
    __shared__ uint32_t s_red[4][32];

    const int tile = __ldg(a.tile_order + blockIdx.x);
    const int2 range = __ldg(reinterpret_cast<const int2*>(a.tile_ranges) + tile);
    const int len = range.y - range.x;
    if (len <= len_lo || len > len_hi) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long* src = a.pairs + range.x;
    const int wbase = warp * 32 * IPT;

    uint32_t key[IPT], val[IPT];
    uint32_t kor = 0u, kand = ~0u, vor = 0u, vand = ~0u;
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
        const int idx = wbase + j * 32 + lane;
        // padding sorts behind every real entry in every pass (all digit bits set) and starts behind them
        key[j] = ~0u; val[j] = ~0u;
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
                const uint32_t d = ((by_val ? val[j] : key[j]) >> shift) & mask;
                uint32_t peers = 0xffffffffu;
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
                if (lane == leader) {
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
        return;
    }
    for (int e = tid; e < len; e += THREADS) a.ids_sorted[range.x + e] = (int32_t)svals[e];
}

// Segments longer than the shared-memory classes: bitonic network over the 64-bit records (depth << 32 | id
// order = the wanted order) in global memory, all comparators ascending (the "flip" form), so that virtual
// +inf padding beyond the segment never moves and the length need not be a power of two.
__global__ void __launch_bounds__(1024)
tile_sort_big_kernel(const SortArgs a, int len_lo) {
    const int tile = __ldg(a.tile_order + blockIdx.x);
    const int2 range = __ldg(reinterpret_cast<const int2*>(a.tile_ranges) + tile);
    const int len = range.y - range.x;
    if (len <= len_lo) return;
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

struct Bin2Layout {
    size_t counts, info, pairs, bytes;
};

static Bin2Layout bin2_layout(long long num_tiles, long long capacity) {
    Bin2Layout L;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 255) & ~(size_t)255; return r; };
    L.info = take(sizeof(int32_t) * 4);
    L.counts = take(sizeof(int32_t) * (size_t)(num_tiles > 0 ? num_tiles : 1));
    L.pairs = take(sizeof(unsigned long long) * (size_t)(capacity > 0 ? capacity : 1));
    L.bytes = o;
    return L;
}

}  // namespace gg

using namespace gg;

extern "C" size_t gg_bin_tiles_scratch_bytes(long long num_tiles, long long capacity) {
    return bin2_layout(num_tiles, capacity).bytes;
}

extern "C" int gg_bin_tiles(int n, int n_views, const float* xys, int xy_stride, const float* depths,
                            const int32_t* radii, int tiles_x, int tiles_y, long long capacity, void* scratch,
                            size_t scratch_bytes, int32_t* ids_sorted, int32_t* tile_ranges, int32_t* tile_order,
                            int32_t* info_dev, int32_t* info_host, void* stream) {
    GG_REQUIRE(n >= 1 && n_views >= 1 && (long long)n * n_views < (1ll << 31), "gg_bin_tiles: bad sizes");
    GG_REQUIRE(tiles_x >= 1 && tiles_y >= 1 && (long long)n_views * tiles_x * tiles_y < (1ll << 30),
               "gg_bin_tiles: too many tiles");
    GG_REQUIRE(capacity >= 0 && capacity < (1ll << 31), "gg_bin_tiles: capacity out of range");
    GG_REQUIRE(xys && depths && radii && scratch && tile_ranges && tile_order && (ids_sorted || capacity == 0),
               "gg_bin_tiles: null pointer");
    GG_REQUIRE((xy_stride == 2 || xy_stride == 8) && ((uintptr_t)xys & 7) == 0, "gg_bin_tiles: bad xys");
    GG_REQUIRE(((uintptr_t)scratch & 255) == 0 && ((uintptr_t)tile_ranges & 7) == 0, "gg_bin_tiles: misaligned");
    const long long total = (long long)n * n_views;
    const int num_tiles = n_views * tiles_x * tiles_y;
    const Bin2Layout L = bin2_layout(num_tiles, capacity);
    GG_REQUIRE(scratch_bytes >= L.bytes, "gg_bin_tiles: scratch too small");
    unsigned char* s = reinterpret_cast<unsigned char*>(scratch);
    int* counts = reinterpret_cast<int*>(s + L.counts);
    int32_t* info = info_dev ? info_dev : reinterpret_cast<int32_t*>(s + L.info);
    unsigned long long* pairs = reinterpret_cast<unsigned long long*>(s + L.pairs);
    cudaStream_t st = (cudaStream_t)stream;
    GG_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * (size_t)num_tiles, st));
    tile_count_kernel<<<div_up(total, 256), 256, 0, st>>>(total, n, xys, xy_stride, radii, tiles_x, tiles_y, counts);
    tile_scan_order_kernel<<<1, 1024, 0, st>>>(num_tiles, capacity, counts, tile_ranges, tile_order, info);
    int launches = 2;
    if (capacity > 0) {
        tile_scatter_kernel<<<div_up(total, 256), 256, 0, st>>>(total, n, xys, xy_stride, depths, radii, tiles_x,
                                                                tiles_y, counts, info, capacity, pairs);
        SortArgs a{tile_ranges, tile_order, pairs, ids_sorted, num_tiles};
        // tile_order lists the tiles by descending length bucket (8 entries wide, saturating at 8184), so the
        // tiles of a class sit among the first capacity / (class minimum) positions
        auto grid_for = [&](long long min_len) {
            const long long g = capacity / min_len + 1;
            return (unsigned)(g < num_tiles ? g : num_tiles);
        };
        {
            constexpr size_t sm = tile_sort_smem<128, 8>();
            GG_CUDA(cudaFuncSetAttribute(tile_sort_kernel<128, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            tile_sort_kernel<128, 8><<<(unsigned)num_tiles, 128, sm, st>>>(a, 0, 1024);
        }
        {
            constexpr size_t sm = tile_sort_smem<256, 16>();
            GG_CUDA(cudaFuncSetAttribute(tile_sort_kernel<256, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            tile_sort_kernel<256, 16><<<grid_for(1024), 256, sm, st>>>(a, 1024, 4096);
        }
        {
            constexpr size_t sm = tile_sort_smem<1024, 16>();
            GG_CUDA(cudaFuncSetAttribute(tile_sort_kernel<1024, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            tile_sort_kernel<1024, 16><<<grid_for(4096), 1024, sm, st>>>(a, 4096, kSortMaxSmem);
        }
        tile_sort_big_kernel<<<grid_for(8184), 1024, 0, st>>>(a, kSortMaxSmem);
        launches += 5;
    }
    count_launch(launches);
    if (info_host) GG_CUDA(cudaMemcpyAsync(info_host, info, sizeof(int32_t) * 4, cudaMemcpyDeviceToHost, st));
    return check_launch("gg_bin_tiles");
}
