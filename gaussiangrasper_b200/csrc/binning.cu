// Tile binning: prefix sum of intersection counts, 64-bit tile|depth key emission, in-house
// onesweep LSD radix sort (64-bit keys, 32-bit payload) and per-tile ranges.
//
// Replaces gsplat 0.1.0 utils.compute_cumulative_intersects / bin_and_sort_gaussians
// (torch.cumsum + map_gaussian_to_intersects + torch.sort + gather + get_tile_bin_edges), which
// the reference re-runs inside every RasterizeGaussians / NDRasterizeGaussians.forward
// (nerfstudio/models/gaussian_splatting.py:735,747,759,773).  All outputs are integers and are
// bit-exact against the CPU oracle; ties in the sort are broken by ascending Gaussian id because
// emission is id-ordered and every pass of the LSD sort is stable.
//
// Compiled with -fmad=false (tile_box() must reproduce the projection kernel's tile bbox).
#include "gg_common.cuh"
#include "gg_math.cuh"
#include "gg_b200.h"

#include <atomic>

namespace gg {

// ---------------------------------------------------------------------------------------------
// Inclusive int32 prefix sum: reduce -> spine -> downsweep.  4096 items per block.
// ---------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanIpt = 16;
constexpr int kScanTile = kScanThreads * kScanIpt;

__device__ __forceinline__ int warp_incl_scan(int v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// exclusive scan of one int per thread across a 256-thread block; returns exclusive prefix and
// (through total) the block total.  sm must hold 8 ints.
__device__ __forceinline__ int block_excl_scan_256(int v, int* sm, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int incl = warp_incl_scan(v);
    if (lane == 31) sm[warp] = incl;
    __syncthreads();
    int woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const int t = sm[w];
        if (w < warp) woff += t;
        tot += t;
    }
    total = tot;
    __syncthreads();
    return woff + incl - v;
}

__global__ void __launch_bounds__(kScanThreads)
scan_reduce_kernel(long long n, const int32_t* __restrict__ in, int32_t* __restrict__ block_sums) {
    __shared__ int sm[8];
    const long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanIpt;
    int s = 0;
    if (base + kScanIpt <= n) {
        const int4* p = reinterpret_cast<const int4*>(in + base);
#pragma unroll
        for (int k = 0; k < kScanIpt / 4; ++k) { const int4 v = __ldg(p + k); s += v.x + v.y + v.z + v.w; }
    } else {
        for (int k = 0; k < kScanIpt; ++k) if (base + k < n) s += __ldg(in + base + k);
    }
    int total;
    block_excl_scan_256(s, sm, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block: exclusive scan of block_sums in place, grand total to *total_out (and *total_out2)
__global__ void __launch_bounds__(kScanThreads)
scan_spine_kernel(int nb, int32_t* __restrict__ block_sums, int32_t* __restrict__ total_out) {
    __shared__ int sm[8];
    int carry = 0;
    for (int base = 0; base < nb; base += kScanThreads) {
        const int i = base + threadIdx.x;
        const int v = i < nb ? block_sums[i] : 0;
        int total;
        const int ex = block_excl_scan_256(v, sm, total);
        if (i < nb) block_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void __launch_bounds__(kScanThreads)
scan_down_kernel(long long n, const int32_t* __restrict__ in, const int32_t* __restrict__ block_offsets,
                 int32_t* __restrict__ out) {
    __shared__ int sm[8];
    const long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanIpt;
    int v[kScanIpt];
    const bool full = base + kScanIpt <= n;
    if (full) {
        const int4* p = reinterpret_cast<const int4*>(in + base);
#pragma unroll
        for (int k = 0; k < kScanIpt / 4; ++k) {
            const int4 t = __ldg(p + k);
            v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < kScanIpt; ++k) v[k] = (base + k < n) ? __ldg(in + base + k) : 0;
    }
    int s = 0;
#pragma unroll
    for (int k = 0; k < kScanIpt; ++k) { s += v[k]; v[k] = s; }
    int total;
    const int ex = block_excl_scan_256(s, sm, total) + block_offsets[blockIdx.x];
    if (full) {
        int4* p = reinterpret_cast<int4*>(out + base);
#pragma unroll
        for (int k = 0; k < kScanIpt / 4; ++k)
            p[k] = make_int4(v[4 * k] + ex, v[4 * k + 1] + ex, v[4 * k + 2] + ex, v[4 * k + 3] + ex);
    } else {
#pragma unroll
        for (int k = 0; k < kScanIpt; ++k) if (base + k < n) out[base + k] = v[k] + ex;
    }
}

// ---------------------------------------------------------------------------------------------
// Key emission.  One thread per (view, Gaussian); entry k of Gaussian g goes to cum[g-1] + k with
// tiles visited row-major inside the bbox.  key = (view * T + tile) << 32 | depth bits.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
emit_keys_kernel(int n, int n_views, const float* __restrict__ xys, int xy_stride, const float* __restrict__ depths,
                 const int32_t* __restrict__ radii, const int32_t* __restrict__ cum, int tiles_x, int tiles_y,
                 int64_t* __restrict__ keys, int32_t* __restrict__ ids) {
    const long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= (long long)n * n_views) return;
    const int r = radii[gi];
    if (r <= 0) return;
    const int view = (int)(gi / n);
    const int g = (int)(gi - (long long)view * n);
    const float2 c = __ldg(reinterpret_cast<const float2*>(xys + gi * xy_stride));
    const TileBox tb = tile_box(c.x, c.y, (float)r, tiles_x, tiles_y);
    long long cur = gi == 0 ? 0 : cum[gi - 1];
    const long long end = cum[gi];  // never write past this Gaussian's slot, whatever the caller's counts say
    const uint64_t dbits = (uint64_t)__float_as_uint(depths[gi]);
    const long long tile0 = (long long)view * tiles_x * tiles_y;
    for (int ty = tb.y0; ty < tb.y1; ++ty)
        for (int tx = tb.x0; tx < tb.x1 && cur < end; ++tx) {
            const uint64_t tile = (uint64_t)(tile0 + (long long)ty * tiles_x + tx);
            keys[cur] = (int64_t)((tile << 32) | dbits);
            ids[cur] = g;
            ++cur;
        }
}

// ---------------------------------------------------------------------------------------------
// Onesweep LSD radix sort, 8-bit digits, 256 threads x 16 keys per tile.
//   1 upfront kernel  : digit histograms of every pass
//   P pass kernels    : rank (warp match), decoupled look-back across tiles, shared-memory
//                       reorder, coalesced scatter
// Digits are balanced: a b-bit key takes ceil(b/8) passes of b/passes (+1) bits each -- an 11-bit tile
// key sorts as 6+5 bits, whose 64/32-bin scatters write 4-8x longer runs than an 8+3 split.
// Tile status words are 64 bit: [generation:32 | flag:2 | count:30]; a per-process generation
// counter makes stale words from earlier passes/calls invisible, so the status array is zeroed
// only once, by whoever allocates the workspace.  Generations carry bit 31, which the high word of
// nothing else this workspace ever holds (keys, ids, counts) has.
// ---------------------------------------------------------------------------------------------
constexpr int kRsThreads = 256;
constexpr int kRsIpt = 16;
constexpr int kRsTile = kRsThreads * kRsIpt;  // 4096
constexpr int kRadix = 256;
constexpr int kMaxPasses = 8;
constexpr uint32_t kFlagAgg = 1u << 30;
constexpr uint32_t kFlagIncl = 1u << 31;
constexpr uint32_t kValueMask = (1u << 30) - 1;
constexpr int kLookback = 4;   // status words fetched per look-back round

struct DigitPlan {
    int passes;
    int shift[kMaxPasses];
    uint32_t mask[kMaxPasses];
};

static DigitPlan digit_plan(int key_bits) {
    DigitPlan P;
    P.passes = (key_bits + 7) / 8;
    const int base = key_bits / P.passes, rem = key_bits % P.passes;
    int shift = 0;
    for (int p = 0; p < kMaxPasses; ++p) {
        const int bits = p < P.passes ? base + (p < rem ? 1 : 0) : 0;
        P.shift[p] = shift;
        P.mask[p] = (1u << bits) - 1u;
        shift += bits;
    }
    return P;
}

__global__ void __launch_bounds__(256)
radix_hist_kernel(const uint64_t* __restrict__ keys, long long m, const DigitPlan plan, uint32_t* __restrict__ ghist) {
    const int passes = plan.passes;
    // Block-shared histograms, plain shared-memory atomics.  Digits on which a whole warp agrees
    // (the high bytes of tile|depth keys are nearly constant) would serialise 32-fold on one
    // address, so those take a match.all fast path: one add of the warp's population.
    __shared__ uint32_t sh[kMaxPasses * kRadix];
    for (int i = threadIdx.x; i < passes * kRadix; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long rounds = (m + stride - 1) / stride;
    // kHistBatch independent loads are in flight before the first key is used: the warp votes and the
    // shared-memory atomics below keep the compiler from overlapping the loads of successive rounds,
    // and one exposed memory latency per key is what a round would cost otherwise
    constexpr int kHistBatch = 8;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long r0 = 0; r0 < rounds; r0 += kHistBatch) {
        uint64_t kb[kHistBatch];
        bool vb[kHistBatch];
#pragma unroll
        for (int j = 0; j < kHistBatch; ++j) {
            const long long i = (r0 + j) * stride + tid;
            vb[j] = (r0 + j < rounds) && i < m;
            kb[j] = vb[j] ? __ldg(keys + i) : 0ull;
        }
#pragma unroll
        for (int j = 0; j < kHistBatch; ++j) {
            const bool valid = vb[j];
            const unsigned active = __ballot_sync(0xffffffffu, valid);
            if (valid) {
                const uint64_t k = kb[j];
                const int leader = __ffs(active) - 1;
                const uint32_t pop = (uint32_t)__popc(active);
#pragma unroll
                for (int p = 0; p < kMaxPasses; ++p) {  // unrolled: the plan is indexed statically
                    if (p < passes) {
                        const uint32_t d = (uint32_t)(k >> plan.shift[p]) & plan.mask[p];
                        int same;
                        __match_all_sync(active, d, &same);
                        if (same) {
                            if (lane == leader) atomicAdd(&sh[p * kRadix + d], pop);
                        } else {
                            atomicAdd(&sh[p * kRadix + d], 1u);
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * kRadix; i += blockDim.x) {
        const uint32_t v = sh[i];
        if (v) atomicAdd(&ghist[i], v);
    }
}

__global__ void __launch_bounds__(kRsThreads, 3)
radix_pass_kernel(const uint64_t* __restrict__ kin, const uint32_t* __restrict__ vin, uint64_t* __restrict__ kout,
                  uint32_t* __restrict__ vout, long long m, int shift, uint32_t dmask, const uint32_t* __restrict__ ghist_excl,
                  volatile uint64_t* tile_state, uint32_t* ticket, uint32_t generation) {
    // everything lives in dynamic shared memory (59.5 KB > the 48 KB static limit)
    extern __shared__ __align__(16) unsigned char dyn[];
    uint64_t* keys_sm = reinterpret_cast<uint64_t*>(dyn);
    uint32_t* vals_sm = reinterpret_cast<uint32_t*>(dyn + sizeof(uint64_t) * kRsTile);
    uint32_t(*warp_hist)[kRadix] = reinterpret_cast<uint32_t(*)[kRadix]>(vals_sm + kRsTile);
    uint32_t* local_start = reinterpret_cast<uint32_t*>(warp_hist + kRsThreads / 32);
    uint32_t* global_base = local_start + kRadix;
    int* scan_sm = reinterpret_cast<int*>(global_base + kRadix);
    uint32_t& s_tile = *reinterpret_cast<uint32_t*>(scan_sm + 8);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
#pragma unroll
    for (int w = 0; w < kRsThreads / 32; ++w) warp_hist[w][tid] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const long long base = (long long)tile * kRsTile;
    const long long wbase = base + (long long)warp * (32 * kRsIpt);

    uint64_t key[kRsIpt];
#pragma unroll
    for (int j = 0; j < kRsIpt; ++j) {
        const long long idx = wbase + j * 32 + lane;
        key[j] = idx < m ? __ldg(kin + idx) : ~0ull;
    }
    // stable rank inside the warp: items are ordered (j, lane).  All 16 match.any are issued first
    // (they are independent and have a long latency); only the counter updates form a chain.
    uint32_t peers[kRsIpt];
#pragma unroll
    for (int j = 0; j < kRsIpt; ++j) {
        // lanes holding the same digit, from one ballot per digit bit: match.any serialises over the
        // distinct values of a warp (~30 for an 8-bit digit of spread keys) and bounded the whole pass
        const uint32_t d = (uint32_t)(key[j] >> shift) & dmask;
        uint32_t same = 0xffffffffu;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            if ((dmask >> b) & 1u) {  // warp-uniform
                const bool bit = (d >> b) & 1u;
                const uint32_t v = __ballot_sync(0xffffffffu, bit);
                same &= bit ? v : ~v;
            }
        }
        peers[j] = same;
    }
    uint32_t rank[kRsIpt];
#pragma unroll
    for (int j = 0; j < kRsIpt; ++j) {
        const uint32_t d = (uint32_t)(key[j] >> shift) & dmask;
        const uint32_t mask = peers[j];
        const int leader = __ffs(mask) - 1;
        uint32_t prev = 0;
        if (lane == leader) {
            prev = warp_hist[warp][d];
            warp_hist[warp][d] = prev + (uint32_t)__popc(mask);
        }
        prev = __shfl_sync(0xffffffffu, prev, leader);
        rank[j] = prev + (uint32_t)__popc(mask & ((1u << lane) - 1u));
        __syncwarp();
    }
    __syncthreads();
    // thread tid owns digit tid: warp-exclusive offsets, tile count, look-back
    uint32_t count = 0;
#pragma unroll
    for (int w = 0; w < kRsThreads / 32; ++w) {
        const uint32_t t = warp_hist[w][tid];
        warp_hist[w][tid] = count;
        count += t;
    }
    const uint64_t gen = (uint64_t)generation << 32;
    volatile uint64_t* my_state = tile_state + (size_t)tile * kRadix + tid;
    uint32_t excl = 0;
    if (tile == 0) {
        *my_state = gen | kFlagIncl | count;
    } else {
        *my_state = gen | kFlagAgg | count;
        // Windowed look-back: kLookback status words are fetched at once (independent loads), then
        // consumed front to back until one is not published yet or carries an inclusive prefix.
        // A serial walk costs one L2 round trip per predecessor, i.e. ~0.3 us x (tiles in flight),
        // which is what bounds a pass on the few-hundred-tile sorts of this workload.
        long long prev = (long long)tile - 1;
        bool found = false;
        while (!found) {
            uint64_t w[kLookback];
#pragma unroll
            for (int i = 0; i < kLookback; ++i) {
                const long long idx = prev - i;
                w[i] = idx >= 0 ? tile_state[(size_t)idx * kRadix + tid] : (gen | kFlagIncl);
            }
            int consumed = 0;
#pragma unroll
            for (int i = 0; i < kLookback; ++i) {
                if (found || consumed < i) continue;  // stop at the first word that is not ready
                const uint64_t sw = w[i];
                if ((sw >> 32) != generation || ((uint32_t)sw & (kFlagAgg | kFlagIncl)) == 0) continue;
                excl += (uint32_t)sw & kValueMask;
                ++consumed;
                if ((uint32_t)sw & kFlagIncl) found = true;
            }
            prev -= consumed;
        }
        *my_state = gen | kFlagIncl | ((excl + count) & kValueMask);
    }
    int total;
    const uint32_t lstart = (uint32_t)block_excl_scan_256((int)count, scan_sm, total);
    // global digit offsets: exclusive scan of this pass's 256-bin histogram, redone by every tile
    // (256 values; cheaper than a separate launch per sort)
    const uint32_t gstart = (uint32_t)block_excl_scan_256((int)ghist_excl[tid], scan_sm, total);
    local_start[tid] = lstart;
    global_base[tid] = gstart + excl - lstart;
    __syncthreads();
    // reorder through shared memory so that the scatter writes runs of consecutive addresses
#pragma unroll
    for (int j = 0; j < kRsIpt; ++j) {
        const uint32_t d = (uint32_t)(key[j] >> shift) & dmask;
        const uint32_t pos = local_start[d] + warp_hist[warp][d] + rank[j];
        const long long idx = wbase + j * 32 + lane;
        keys_sm[pos] = key[j];
        vals_sm[pos] = idx < m ? __ldg(vin + idx) : 0u;
    }
    __syncthreads();
    const long long remain = m - base;
    const int valid = remain < kRsTile ? (int)remain : kRsTile;
#pragma unroll
    for (int k = 0; k < kRsIpt; ++k) {
        const int pos = k * kRsThreads + tid;
        if (pos < valid) {
            const uint64_t kk = keys_sm[pos];
            const uint32_t d = (uint32_t)(kk >> shift) & dmask;
            const uint32_t gpos = global_base[d] + (uint32_t)pos;
            kout[gpos] = kk;
            vout[gpos] = vals_sm[pos];
        }
    }
}

// tile_ranges[t] = (first, last+1) of the run of sorted keys whose high word is t; must be
// zero-initialised (empty tiles stay (0,0)).
__global__ void __launch_bounds__(256)
tile_ranges_kernel(long long m, const int64_t* __restrict__ keys_sorted, int shift,
                   int32_t* __restrict__ tile_ranges) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int32_t t = (int32_t)(keys_sorted[i] >> shift);
    if (i == 0 || (int32_t)(keys_sorted[i - 1] >> shift) != t) tile_ranges[2 * (size_t)t] = (int32_t)i;
    if (i == m - 1 || (int32_t)(keys_sorted[i + 1] >> shift) != t) tile_ranges[2 * (size_t)t + 1] = (int32_t)(i + 1);
}

// ---------------------------------------------------------------------------------------------
// Depth-first binning (used by the internal path; the gsplat-compatible entry points above keep
// the reference's one-sort formulation).  The (tile, depth, id) order is obtained as
//   1. stable sort of the V*N Gaussians by (view, depth)            -- 4-5 passes over N items
//   2. emission of the tile entries in that order                    -- key = view*T + tile only
//   3. stable sort of the M entries by that tile key                 -- 2-3 passes over M items
// which yields bit for bit the same sorted id list as sorting M 64-bit (tile|depth) keys with
// ties broken by id, at a third of the M-sized passes.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
depth_keys_kernel(long long total, long long n, const float* __restrict__ depths, int64_t* __restrict__ keys,
                  int32_t* __restrict__ rows) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const uint64_t view = (uint64_t)(i / n);
    keys[i] = (int64_t)((view << 32) | (uint64_t)__float_as_uint(depths[i]));
    rows[i] = (int32_t)i;
}

__global__ void __launch_bounds__(256)
gather_counts_kernel(long long total, const int32_t* __restrict__ order, const int32_t* __restrict__ num_tiles_hit,
                     int32_t* __restrict__ counts) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    counts[i] = __ldg(num_tiles_hit + order[i]);
}

// Warp-cooperative emission: a warp owns 32 consecutive slots of the depth order; their tile
// entries form one contiguous output span, written 32 entries per step with coalesced stores (a
// thread-per-Gaussian loop scatters 8+4 byte writes and stalls on the store queue).  Each output
// entry finds its Gaussian by a 5-step binary search over the warp's exclusive offsets.
__global__ void __launch_bounds__(256)
emit_tiles_sorted_kernel(long long total, int n, const int32_t* __restrict__ order, const float* __restrict__ xys,
                         int xy_stride, const int32_t* __restrict__ radii, const int32_t* __restrict__ cum,
                         int tiles_x, int tiles_y, int64_t* __restrict__ keys, int32_t* __restrict__ ids) {
    __shared__ int s_off[8][33];
    __shared__ int4 s_box[8][32];  // x0, y0, width, view * tiles
    __shared__ int s_g[8][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long i0 = i - lane;
    if (i0 >= total) return;
    int cnt = 0, g = 0;
    int4 box = make_int4(0, 0, 1, 0);
    if (i < total) {
        const long long row = order[i];
        const int r = radii[row];
        if (r > 0) {
            const int view = (int)(row / n);
            g = (int)(row - (long long)view * n);
            const float2 c = __ldg(reinterpret_cast<const float2*>(xys + row * xy_stride));
            const TileBox tb = tile_box(c.x, c.y, (float)r, tiles_x, tiles_y);
            cnt = tb.area();
            box = make_int4(tb.x0, tb.y0, max(tb.x1 - tb.x0, 1), view * tiles_x * tiles_y);
        }
    }
    const int incl = warp_incl_scan(cnt);
    s_off[warp][lane] = incl - cnt;
    if (lane == 31) s_off[warp][32] = incl;
    s_box[warp][lane] = box;
    s_g[warp][lane] = g;
    __syncwarp();
    const long long wbase = i0 == 0 ? 0 : (long long)cum[i0 - 1];
    const long long ilast = min(i0 + 31, total - 1);
    // the span this warp may write is fixed by the scanned counts, whatever the recomputed boxes say
    const int total_w = min(s_off[warp][32], (int)((long long)cum[ilast] - wbase));
    for (int j = lane; j < total_w; j += 32) {
        int lo = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1)
            if (s_off[warp][lo + step] <= j) lo += step;
        const int k = j - s_off[warp][lo];
        const int4 bx = s_box[warp][lo];
        const int ty = bx.y + k / bx.z, tx = bx.x + k - (k / bx.z) * bx.z;
        keys[wbase + j] = (int64_t)bx.w + (int64_t)ty * tiles_x + tx;
        ids[wbase + j] = s_g[warp][lo];
    }
}

// ---------------------------------------------------------------------------------------------
// Longest-first tile order for the blend kernels: CTAs are dispatched in blockIdx order, so
// visiting the tiles by descending list length keeps the tail of the launch short (the lists
// vary by two orders of magnitude).  Counting sort on 1024 length buckets; the order inside a
// bucket is whatever the atomics give (it only affects scheduling, never results).
// ---------------------------------------------------------------------------------------------
constexpr int kOrderBins = 1024;
__device__ __forceinline__ int tile_len_bucket(int len) { return min(kOrderBins - 1, len >> 3); }

__global__ void __launch_bounds__(256)
tile_len_hist_kernel(long long num_tiles, const int32_t* __restrict__ tile_ranges, int* __restrict__ bins) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_tiles) return;
    const int2 r = __ldg(reinterpret_cast<const int2*>(tile_ranges) + t);
    atomicAdd(&bins[tile_len_bucket(r.y - r.x)], 1);
}
// bins[b] <- number of tiles in buckets above b (descending exclusive scan); one block of 1024
__global__ void __launch_bounds__(kOrderBins) tile_order_scan_kernel(int* __restrict__ bins) {
    __shared__ int sm[kOrderBins];
    const int b = threadIdx.x;
    sm[b] = bins[kOrderBins - 1 - b];  // reversed: ascending index = descending length
    __syncthreads();
    for (int o = 1; o < kOrderBins; o <<= 1) {
        const int v = b >= o ? sm[b - o] : 0;
        __syncthreads();
        sm[b] += v;
        __syncthreads();
    }
    bins[kOrderBins - 1 - b] = sm[b] - bins[kOrderBins - 1 - b];
}
__global__ void __launch_bounds__(256)
tile_order_scatter_kernel(long long num_tiles, const int32_t* __restrict__ tile_ranges, int* __restrict__ cursor,
                          int32_t* __restrict__ order) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_tiles) return;
    const int2 r = __ldg(reinterpret_cast<const int2*>(tile_ranges) + t);
    order[atomicAdd(&cursor[tile_len_bucket(r.y - r.x)], 1)] = (int32_t)t;
}

// single-block variant for up to a few thousand tiles: histogram, scan and scatter in one launch
__global__ void __launch_bounds__(kOrderBins)
tile_order_small_kernel(int num_tiles, const int32_t* __restrict__ tile_ranges, int32_t* __restrict__ order) {
    __shared__ int bins[kOrderBins];
    __shared__ int sm[kOrderBins];
    const int b = threadIdx.x;
    bins[b] = 0;
    __syncthreads();
    for (int t = b; t < num_tiles; t += kOrderBins) {
        const int2 r = __ldg(reinterpret_cast<const int2*>(tile_ranges) + t);
        atomicAdd(&bins[tile_len_bucket(r.y - r.x)], 1);
    }
    __syncthreads();
    sm[b] = bins[kOrderBins - 1 - b];
    __syncthreads();
    for (int o = 1; o < kOrderBins; o <<= 1) {
        const int v = b >= o ? sm[b - o] : 0;
        __syncthreads();
        sm[b] += v;
        __syncthreads();
    }
    const int start = sm[b] - bins[kOrderBins - 1 - b];
    __syncthreads();
    bins[kOrderBins - 1 - b] = start;
    __syncthreads();
    for (int t = b; t < num_tiles; t += kOrderBins) {
        const int2 r = __ldg(reinterpret_cast<const int2*>(tile_ranges) + t);
        order[atomicAdd(&bins[tile_len_bucket(r.y - r.x)], 1)] = t;
    }
}

static std::atomic<uint32_t> g_generation{1};

struct SortLayout {
    size_t off_hist, off_ticket, off_state, off_keys, off_vals, total;
    long long tiles;
};

static SortLayout sort_layout(long long m) {
    SortLayout L;
    L.tiles = (m + kRsTile - 1) / kRsTile;
    if (L.tiles < 1) L.tiles = 1;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 255) & ~(size_t)255; return r; };
    L.off_hist = take(sizeof(uint32_t) * kMaxPasses * kRadix);
    L.off_ticket = take(sizeof(uint32_t) * kMaxPasses);
    L.off_state = take(sizeof(uint64_t) * (size_t)L.tiles * kRadix);
    L.off_keys = take(sizeof(uint64_t) * (size_t)(m > 0 ? m : 1));
    L.off_vals = take(sizeof(uint32_t) * (size_t)(m > 0 ? m : 1));
    L.total = o;
    return L;
}

}  // namespace gg

using namespace gg;

extern "C" size_t gg_cumsum_workspace_bytes(long long n) {
    return sizeof(int32_t) * (size_t)(div_up(n > 0 ? n : 1, kScanTile) + 1);
}

extern "C" int gg_cumsum(long long n, const int32_t* in, int32_t* out, int32_t* total_dev, void* workspace,
                         size_t workspace_bytes, void* stream) {
    GG_REQUIRE(n >= 1 && n < (1ll << 40), "gg_cumsum: need n >= 1");
    GG_REQUIRE(in && out && workspace, "gg_cumsum: null pointer");
    GG_REQUIRE(workspace_bytes >= gg_cumsum_workspace_bytes(n), "gg_cumsum: workspace too small");
    GG_REQUIRE(((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0, "gg_cumsum: arrays must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = div_up(n, kScanTile);
    int32_t* sums = reinterpret_cast<int32_t*>(workspace);
    scan_reduce_kernel<<<nb, kScanThreads, 0, st>>>(n, in, sums);
    scan_spine_kernel<<<1, kScanThreads, 0, st>>>(nb, sums, total_dev);
    scan_down_kernel<<<nb, kScanThreads, 0, st>>>(n, in, sums, out);
    count_launch(3);
    return check_launch("gg_cumsum");
}

static int map_to_intersects(int n, int n_views, const float* xys, int xy_stride, const float* depths,
                             const int32_t* radii, const int32_t* cum_tiles_hit, int tiles_x, int tiles_y,
                             int64_t* keys, int32_t* ids, void* stream) {
    GG_REQUIRE(n >= 1 && n_views >= 1, "gg_map_to_intersects: need n >= 1");
    GG_REQUIRE(xys && depths && radii && cum_tiles_hit && keys && ids, "gg_map_to_intersects: null pointer");
    GG_REQUIRE(((uintptr_t)xys & 7) == 0, "gg_map_to_intersects: xys misaligned");
    GG_REQUIRE((long long)n_views * tiles_x * tiles_y < (1ll << 31), "gg_map_to_intersects: too many tiles");
    const long long total = (long long)n * n_views;
    emit_keys_kernel<<<div_up(total, 256), 256, 0, (cudaStream_t)stream>>>(n, n_views, xys, xy_stride, depths, radii,
                                                                           cum_tiles_hit, tiles_x, tiles_y, keys, ids);
    count_launch();
    return check_launch("emit_keys_kernel");
}

extern "C" int gg_map_to_intersects(int n, int n_views, const float* xys, const float* depths, const int32_t* radii,
                                    const int32_t* cum_tiles_hit, int tiles_x, int tiles_y, int64_t* keys,
                                    int32_t* ids, void* stream) {
    return map_to_intersects(n, n_views, xys, 2, depths, radii, cum_tiles_hit, tiles_x, tiles_y, keys, ids, stream);
}

// same, reading the pixel centres from the packed geo records ([V*n, 8], x and y first)
extern "C" int gg_map_to_intersects_geo(int n, int n_views, const float* geo, const float* depths,
                                        const int32_t* radii, const int32_t* cum_tiles_hit, int tiles_x, int tiles_y,
                                        int64_t* keys, int32_t* ids, void* stream) {
    return map_to_intersects(n, n_views, geo, 8, depths, radii, cum_tiles_hit, tiles_x, tiles_y, keys, ids, stream);
}

extern "C" size_t gg_sort_workspace_bytes(long long m) { return sort_layout(m).total; }

extern "C" int gg_sort_pairs(long long m, int key_bits, const int64_t* keys_in, const int32_t* vals_in,
                             int64_t* keys_out, int32_t* vals_out, void* workspace, size_t workspace_bytes,
                             void* stream) {
    GG_REQUIRE(m >= 0 && m < (1ll << 30), "gg_sort_pairs: need 0 <= m < 2^30");
    GG_REQUIRE(key_bits >= 1 && key_bits <= 64, "gg_sort_pairs: key_bits out of range");
    if (m == 0) return GG_OK;
    GG_REQUIRE(keys_in && vals_in && keys_out && vals_out && workspace, "gg_sort_pairs: null pointer");
    const SortLayout L = sort_layout(m);
    GG_REQUIRE(workspace_bytes >= L.total, "gg_sort_pairs: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
    uint32_t* ghist = reinterpret_cast<uint32_t*>(ws + L.off_hist);
    uint32_t* tickets = reinterpret_cast<uint32_t*>(ws + L.off_ticket);
    uint64_t* state = reinterpret_cast<uint64_t*>(ws + L.off_state);
    uint64_t* kalt = reinterpret_cast<uint64_t*>(ws + L.off_keys);
    uint32_t* valt = reinterpret_cast<uint32_t*>(ws + L.off_vals);
    const DigitPlan plan = digit_plan(key_bits);
    const int passes = plan.passes;

    const size_t dyn = (sizeof(uint64_t) + sizeof(uint32_t)) * kRsTile +
                       sizeof(uint32_t) * ((kRsThreads / 32) * kRadix + 2 * kRadix) + sizeof(int) * 8 + 16;
    GG_CUDA(cudaFuncSetAttribute(radix_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    // histograms + tickets are contiguous at the start of the workspace
    GG_CUDA(cudaMemsetAsync(ws, 0, L.off_state, st));
    // few, fat blocks: every block ends with passes*256 global atomics onto the same bins
    int hist_blocks = div_up(m, 256 * 16);
    if (hist_blocks > 148 * 2) hist_blocks = 148 * 2;
    radix_hist_kernel<<<hist_blocks, 256, 0, st>>>(reinterpret_cast<const uint64_t*>(keys_in), m, plan, ghist);
    const uint64_t* kin = reinterpret_cast<const uint64_t*>(keys_in);
    const uint32_t* vin = reinterpret_cast<const uint32_t*>(vals_in);
    for (int p = 0; p < passes; ++p) {
        // the last pass must land in keys_out: passes-1-p even -> out, odd -> alt
        const bool to_out = ((passes - 1 - p) & 1) == 0;
        uint64_t* ko = to_out ? reinterpret_cast<uint64_t*>(keys_out) : kalt;
        uint32_t* vo = to_out ? reinterpret_cast<uint32_t*>(vals_out) : valt;
        const uint32_t gen = g_generation.fetch_add(1);
        radix_pass_kernel<<<(unsigned)L.tiles, kRsThreads, dyn, st>>>(kin, vin, ko, vo, m, plan.shift[p], plan.mask[p],
                                                                      ghist + p * kRadix, state, tickets + p,
                                                                      gen | 0x80000000u);
        kin = ko;
        vin = vo;
    }
    count_launch(1 + passes);
    return check_launch("gg_sort_pairs");
}

static int tile_ranges_impl(long long m, const int64_t* keys_sorted, int key_shift, long long num_tiles,
                            int32_t* tile_ranges, void* stream);

extern "C" int gg_tile_ranges(long long m, const int64_t* keys_sorted, long long num_tiles, int32_t* tile_ranges,
                              void* stream) {
    return tile_ranges_impl(m, keys_sorted, 32, num_tiles, tile_ranges, stream);
}

// same for keys that hold the tile id in their low bits (depth-first binning)
extern "C" int gg_tile_ranges_lowkey(long long m, const int64_t* keys_sorted, long long num_tiles,
                                     int32_t* tile_ranges, void* stream) {
    return tile_ranges_impl(m, keys_sorted, 0, num_tiles, tile_ranges, stream);
}

static int tile_ranges_impl(long long m, const int64_t* keys_sorted, int key_shift, long long num_tiles,
                            int32_t* tile_ranges, void* stream) {
    GG_REQUIRE(m >= 0 && num_tiles >= 1, "gg_tile_ranges: bad sizes");
    GG_REQUIRE(tile_ranges && (m == 0 || keys_sorted), "gg_tile_ranges: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    GG_CUDA(cudaMemsetAsync(tile_ranges, 0, sizeof(int32_t) * 2 * (size_t)num_tiles, st));
    if (m > 0) {
        tile_ranges_kernel<<<div_up(m, 256), 256, 0, st>>>(m, keys_sorted, key_shift, tile_ranges);
        count_launch();
    }
    return check_launch("tile_ranges_kernel");
}

extern "C" size_t gg_tile_order_workspace_bytes(void) { return sizeof(int) * kOrderBins; }

extern "C" int gg_tile_order(long long num_tiles, const int32_t* tile_ranges, int32_t* tile_order, void* workspace,
                             size_t workspace_bytes, void* stream) {
    GG_REQUIRE(num_tiles >= 1 && num_tiles < (1ll << 31), "gg_tile_order: bad tile count");
    GG_REQUIRE(tile_ranges && tile_order && workspace, "gg_tile_order: null pointer");
    GG_REQUIRE(workspace_bytes >= gg_tile_order_workspace_bytes(), "gg_tile_order: workspace too small");
    GG_REQUIRE(((uintptr_t)tile_ranges & 7) == 0, "gg_tile_order: tile_ranges misaligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (num_tiles <= 8 * kOrderBins) {
        tile_order_small_kernel<<<1, kOrderBins, 0, st>>>((int)num_tiles, tile_ranges, tile_order);
        count_launch();
        return check_launch("tile_order_small_kernel");
    }
    int* bins = reinterpret_cast<int*>(workspace);
    GG_CUDA(cudaMemsetAsync(bins, 0, sizeof(int) * kOrderBins, st));
    tile_len_hist_kernel<<<div_up(num_tiles, 256), 256, 0, st>>>(num_tiles, tile_ranges, bins);
    tile_order_scan_kernel<<<1, kOrderBins, 0, st>>>(bins);
    tile_order_scatter_kernel<<<div_up(num_tiles, 256), 256, 0, st>>>(num_tiles, tile_ranges, bins, tile_order);
    count_launch(3);
    return check_launch("gg_tile_order");
}

extern "C" int gg_depth_keys(long long n, int n_views, const float* depths, int64_t* keys, int32_t* rows,
                             void* stream) {
    GG_REQUIRE(n >= 1 && n_views >= 1 && n * n_views < (1ll << 31), "gg_depth_keys: bad sizes");
    GG_REQUIRE(depths && keys && rows, "gg_depth_keys: null pointer");
    const long long total = n * n_views;
    depth_keys_kernel<<<div_up(total, 256), 256, 0, (cudaStream_t)stream>>>(total, n, depths, keys, rows);
    count_launch();
    return check_launch("depth_keys_kernel");
}

extern "C" int gg_gather_counts(long long total, const int32_t* order, const int32_t* num_tiles_hit, int32_t* counts,
                                void* stream) {
    GG_REQUIRE(total >= 1, "gg_gather_counts: bad size");
    GG_REQUIRE(order && num_tiles_hit && counts, "gg_gather_counts: null pointer");
    gather_counts_kernel<<<div_up(total, 256), 256, 0, (cudaStream_t)stream>>>(total, order, num_tiles_hit, counts);
    count_launch();
    return check_launch("gather_counts_kernel");
}

extern "C" int gg_emit_tiles_sorted(int n, int n_views, const int32_t* order, const float* xys, int xy_stride,
                                    const int32_t* radii, const int32_t* cum_sorted, int tiles_x, int tiles_y,
                                    int64_t* keys, int32_t* ids, void* stream) {
    GG_REQUIRE(n >= 1 && n_views >= 1, "gg_emit_tiles_sorted: need n >= 1");
    GG_REQUIRE(order && xys && radii && cum_sorted && keys && ids, "gg_emit_tiles_sorted: null pointer");
    GG_REQUIRE((xy_stride == 2 || xy_stride == 8) && ((uintptr_t)xys & 7) == 0, "gg_emit_tiles_sorted: bad xys");
    GG_REQUIRE((long long)n_views * tiles_x * tiles_y < (1ll << 31), "gg_emit_tiles_sorted: too many tiles");
    const long long total = (long long)n * n_views;
    emit_tiles_sorted_kernel<<<div_up(total, 256), 256, 0, (cudaStream_t)stream>>>(
        total, n, order, xys, xy_stride, radii, cum_sorted, tiles_x, tiles_y, keys, ids);
    count_launch();
    return check_launch("emit_tiles_sorted_kernel");
}

// ---------------------------------------------------------------------------------------------
// Whole depth-first binning in two host calls (the sequences above, enqueued from C so that the
// host spends microseconds, not a Python call per kernel, between launches).
// ---------------------------------------------------------------------------------------------
namespace gg {

struct BeginLayout {
    size_t dkeys, rows, dkeys_sorted, order, counts, cum, total_dev, scan, sort, bytes;
};

static BeginLayout begin_layout(long long total) {
    BeginLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 255) & ~(size_t)255; return r; };
    const size_t t = (size_t)(total > 0 ? total : 1);
    // the sort workspace comes first: its look-back words must sit at offsets that do not move
    // with `total` more than the workspace itself does (the caller zero-fills the scratch once)
    L.sort = take(gg_sort_workspace_bytes(total));
    L.scan = take(gg_cumsum_workspace_bytes(total));
    L.total_dev = take(sizeof(int32_t));
    L.dkeys = take(sizeof(int64_t) * t);
    L.dkeys_sorted = take(sizeof(int64_t) * t);
    L.rows = take(sizeof(int32_t) * t);
    L.order = take(sizeof(int32_t) * t);
    L.counts = take(sizeof(int32_t) * t);
    L.cum = take(sizeof(int32_t) * t);
    L.bytes = o;
    return L;
}

struct FinishLayout {
    size_t sort, order_ws, keys, keys_sorted, ids, bytes;
};

static FinishLayout finish_layout(long long m) {
    FinishLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 255) & ~(size_t)255; return r; };
    const size_t t = (size_t)(m > 0 ? m : 1);
    L.sort = take(gg_sort_workspace_bytes(m));
    L.order_ws = take(gg_tile_order_workspace_bytes());
    L.keys = take(sizeof(int64_t) * t);
    L.keys_sorted = take(sizeof(int64_t) * t);
    L.ids = take(sizeof(int32_t) * t);
    L.bytes = o;
    return L;
}

constexpr int kMaxDevices = 64;
static cudaEvent_t g_count_event[kMaxDevices] = {};

}  // namespace gg

extern "C" size_t gg_bin_begin_scratch_bytes(long long total) { return begin_layout(total).bytes; }
extern "C" size_t gg_bin_finish_scratch_bytes(long long m) { return finish_layout(m).bytes; }

extern "C" int gg_bin_begin(int n, int n_views, const float* depths, const int32_t* num_tiles_hit, void* scratch,
                            size_t scratch_bytes, int32_t* m_host, void* stream) {
    GG_REQUIRE(n >= 1 && n_views >= 1 && (long long)n * n_views < (1ll << 31), "gg_bin_begin: bad sizes");
    GG_REQUIRE(depths && num_tiles_hit && scratch && m_host, "gg_bin_begin: null pointer");
    const long long total = (long long)n * n_views;
    const BeginLayout L = begin_layout(total);
    GG_REQUIRE(scratch_bytes >= L.bytes, "gg_bin_begin: scratch too small");
    GG_REQUIRE(((uintptr_t)scratch & 255) == 0, "gg_bin_begin: scratch must be 256-byte aligned");
    unsigned char* s = reinterpret_cast<unsigned char*>(scratch);
    int64_t* dkeys = reinterpret_cast<int64_t*>(s + L.dkeys);
    int64_t* dkeys_sorted = reinterpret_cast<int64_t*>(s + L.dkeys_sorted);
    int32_t* rows = reinterpret_cast<int32_t*>(s + L.rows);
    int32_t* order = reinterpret_cast<int32_t*>(s + L.order);
    int32_t* counts = reinterpret_cast<int32_t*>(s + L.counts);
    int32_t* cum = reinterpret_cast<int32_t*>(s + L.cum);
    int32_t* total_dev = reinterpret_cast<int32_t*>(s + L.total_dev);
    int view_bits = 0;
    while ((1 << view_bits) < n_views) ++view_bits;
    int rc;
    if ((rc = gg_depth_keys(n, n_views, depths, dkeys, rows, stream))) return rc;
    if ((rc = gg_sort_pairs(total, 32 + view_bits, dkeys, rows, dkeys_sorted, order, s + L.sort,
                            gg_sort_workspace_bytes(total), stream)))
        return rc;
    if ((rc = gg_gather_counts(total, order, num_tiles_hit, counts, stream))) return rc;
    if ((rc = gg_cumsum(total, counts, cum, total_dev, s + L.scan, gg_cumsum_workspace_bytes(total), stream))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    GG_CUDA(cudaMemcpyAsync(m_host, total_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    int dev = 0;
    GG_CUDA(cudaGetDevice(&dev));
    GG_REQUIRE(dev >= 0 && dev < kMaxDevices, "gg_bin_begin: device index out of range");
    if (!g_count_event[dev]) GG_CUDA(cudaEventCreateWithFlags(&g_count_event[dev], cudaEventDisableTiming));
    GG_CUDA(cudaEventRecord(g_count_event[dev], st));
    return GG_OK;
}

extern "C" int gg_bin_wait(void) {
    int dev = 0;
    GG_CUDA(cudaGetDevice(&dev));
    GG_REQUIRE(dev >= 0 && dev < kMaxDevices && g_count_event[dev], "gg_bin_wait: no gg_bin_begin on this device");
    GG_CUDA(cudaEventSynchronize(g_count_event[dev]));
    return GG_OK;
}

extern "C" int gg_bin_finish(int n, int n_views, long long m, const float* xys, int xy_stride, const int32_t* radii,
                             int tiles_x, int tiles_y, const void* begin_scratch, void* scratch, size_t scratch_bytes,
                             int32_t* ids_sorted, int32_t* tile_ranges, int32_t* tile_order, void* stream) {
    GG_REQUIRE(n >= 1 && n_views >= 1 && m >= 1 && m < (1ll << 30), "gg_bin_finish: bad sizes");
    GG_REQUIRE(xys && radii && begin_scratch && scratch && ids_sorted && tile_ranges && tile_order,
               "gg_bin_finish: null pointer");
    const long long total = (long long)n * n_views;
    const BeginLayout B = begin_layout(total);
    const FinishLayout L = finish_layout(m);
    GG_REQUIRE(scratch_bytes >= L.bytes, "gg_bin_finish: scratch too small");
    GG_REQUIRE(((uintptr_t)scratch & 255) == 0, "gg_bin_finish: scratch must be 256-byte aligned");
    const unsigned char* b = reinterpret_cast<const unsigned char*>(begin_scratch);
    unsigned char* s = reinterpret_cast<unsigned char*>(scratch);
    const int32_t* order = reinterpret_cast<const int32_t*>(b + B.order);
    const int32_t* cum = reinterpret_cast<const int32_t*>(b + B.cum);
    int64_t* keys = reinterpret_cast<int64_t*>(s + L.keys);
    int64_t* keys_sorted = reinterpret_cast<int64_t*>(s + L.keys_sorted);
    int32_t* ids = reinterpret_cast<int32_t*>(s + L.ids);
    const long long num_tiles = (long long)n_views * tiles_x * tiles_y;
    int tile_bits = 1;
    while ((1ll << tile_bits) < num_tiles) ++tile_bits;
    int rc;
    if ((rc = gg_emit_tiles_sorted(n, n_views, order, xys, xy_stride, radii, cum, tiles_x, tiles_y, keys, ids, stream)))
        return rc;
    if ((rc = gg_sort_pairs(m, tile_bits, keys, ids, keys_sorted, ids_sorted, s + L.sort, gg_sort_workspace_bytes(m),
                            stream)))
        return rc;
    if ((rc = gg_tile_ranges_lowkey(m, keys_sorted, num_tiles, tile_ranges, stream))) return rc;
    return gg_tile_order(num_tiles, tile_ranges, tile_order, s + L.order_ws, gg_tile_order_workspace_bytes(), stream);
}
