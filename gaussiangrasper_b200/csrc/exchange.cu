// Gradient exchange of a view-sharded training step over NVLink 5 / NVSwitch, as ONE kernel over symmetric
// (peer-mapped) memory instead of two NCCL collectives (BASELINE north_star "per-Gaussian gradients are
// all-reduced over NVLink").
//
// Every rank holds, at the same offset of a symmetric allocation,
//   bucket  : its own leaf gradients except the SH coefficients' (N x (11 + D) floats, GradientBucket layout)
//   rgb_all : [world, V, N, 3] per-view colour gradients -- the factor the SH gradient is rebuilt from
//             (distributed.FactoredExchange); a rank has written only its own slot.
// gg_nvls_exchange then does, in one launch per rank,
//   * all-gather of the rgb slots : the rank copies its slot into the same slot of every peer
//                                   (multimem.st on the multicast address: one store, the switch replicates it);
//   * all-reduce of the bucket    : two-shot through the switch -- the rank owns slice `rank` of the bucket, pulls
//                                   the SUM of all ranks' copies of it with multimem.ld_reduce (the reduction happens
//                                   in the NVSwitch), and multimem.st's the result back into every rank's copy.
//     Per GPU and direction that moves (world-1)/world of the bucket once: the bandwidth-optimal schedule, with
//     no intermediate buffer and no reduction arithmetic on the SMs.
// The same schedule also runs over the peers' mapped addresses with plain ld.global / st.global (W loads in flight
// per element, fixed summation order: every rank computes the same bits) -- and that is the DEFAULT: on this pool's
// NVSwitch boxes it measured faster than the multimem form at 2 ranks and, for buckets of config 1's size, at 8
// (profiles/r02_exchange_tuning.txt); reductions of >= 80 MB over >= 4 ranks take the multimem form (fewer bytes per
// link).  GG_NVLS_MODE=0 / 1 / 2 force multimem / peer / peer loads + multimem stores.  The caller brackets the launch with cross-rank barriers (the symmetric-memory signal
// pads): one before (every rank's backward has written its bucket and slot), one after (every slice is back).
#include "gg_common.cuh"
#include "gg_b200.h"
#include <cstdlib>

namespace gg {

constexpr int kMaxRanks = 16;

struct ExchangeArgs {
    int rank, world, parts;        // parts: bit 0 = gather the rgb slots, bit 1 = reduce the bucket
    float* bucket_mc;              // multicast address of the bucket, or nullptr
    float* bucket_peer[kMaxRanks]; // every rank's bucket as mapped here (own included)
    long long bucket_vec4;         // float4 elements of the bucket
    float* rgb_mc;
    float* rgb_peer[kMaxRanks];
    long long slot_vec4;           // float4 elements of one rank's rgb slot
};

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(mc)
                 : "memory");
    return v;
}

__device__ __forceinline__ void multimem_st(float* mc, const float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

// kMode 0: multimem.ld_reduce + multimem.st; 1: peer loads + peer stores (no multicast mapping needed);
//       2: peer loads (the W slices arrive over W-1 independent NVLink paths) + multimem.st (one store, replicated)
template <int kMode>
__global__ void __launch_bounds__(512)
nvls_exchange_kernel(const ExchangeArgs a) {
    constexpr bool kMulticast = kMode != 1;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    // ---- all-gather: my rgb slot -> the same slot on every rank -------------------------------------
    if (a.parts & 1) {
        const long long base = (long long)a.rank * a.slot_vec4;
        const float4* src = reinterpret_cast<const float4*>(a.rgb_peer[a.rank]) + base;
        for (long long i = tid; i < a.slot_vec4; i += stride) {
            const float4 v = src[i];
            if (kMulticast) {
                multimem_st(a.rgb_mc + 4 * (base + i), v);
            } else {
                for (int p = 0; p < a.world; ++p)
                    if (p != a.rank) reinterpret_cast<float4*>(a.rgb_peer[p])[base + i] = v;
            }
        }
    }
    // ---- all-reduce: slice `rank` of the bucket, summed over the ranks, written back to all of them ----
    if (a.parts & 2) {
        const long long per = (a.bucket_vec4 + a.world - 1) / a.world;
        const long long lo = (long long)a.rank * per;
        const long long hi = lo + per < a.bucket_vec4 ? lo + per : a.bucket_vec4;
        if (kMode == 0) {
            // a round trip through the switch per load: keep kUnroll of them in flight per thread
            constexpr int kUnroll = 4;
            long long i = lo + tid;
            for (; i + (kUnroll - 1) * stride < hi; i += kUnroll * stride) {
                float4 v[kUnroll];
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) v[u] = multimem_ld_reduce_add(a.bucket_mc + 4 * (i + u * stride));
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) multimem_st(a.bucket_mc + 4 * (i + u * stride), v[u]);
            }
            for (; i < hi; i += stride) multimem_st(a.bucket_mc + 4 * i, multimem_ld_reduce_add(a.bucket_mc + 4 * i));
        } else {
            for (long long i = lo + tid; i < hi; i += stride) {
                // fixed summation order (rank 0, 1, ...) so that every run and every rank gives the same bits
                float4 v[kMaxRanks];
#pragma unroll
                for (int p = 0; p < kMaxRanks; ++p)
                    if (p < a.world) v[p] = reinterpret_cast<const float4*>(a.bucket_peer[p])[i];   // W loads in flight
                float4 s = v[0];
#pragma unroll
                for (int p = 1; p < kMaxRanks; ++p)
                    if (p < a.world) { s.x += v[p].x; s.y += v[p].y; s.z += v[p].z; s.w += v[p].w; }
                if (kMode == 2) {
                    multimem_st(a.bucket_mc + 4 * i, s);
                } else {
                    for (int p = 0; p < a.world; ++p) reinterpret_cast<float4*>(a.bucket_peer[p])[i] = s;
                }
            }
        }
    }
}

}  // namespace gg

using namespace gg;

extern "C" int gg_nvls_exchange(int rank, int world, float* bucket_multicast, float* const* bucket_peers,
                                long long bucket_floats, float* rgb_multicast, float* const* rgb_peers,
                                long long rgb_slot_floats, int parts, void* stream) {
    GG_REQUIRE(world >= 2 && world <= kMaxRanks && rank >= 0 && rank < world, "gg_nvls_exchange: 2..16 ranks");
    GG_REQUIRE(bucket_peers && rgb_peers, "gg_nvls_exchange: null peer table");
    GG_REQUIRE(parts >= 1 && parts <= 3, "gg_nvls_exchange: parts is 1 (gather), 2 (reduce) or 3 (both)");
    GG_REQUIRE(bucket_floats >= 0 && rgb_slot_floats >= 0 && bucket_floats % 4 == 0 && rgb_slot_floats % 4 == 0,
               "gg_nvls_exchange: sizes must be multiples of 4 floats");
    GG_REQUIRE((bucket_multicast == nullptr) == (rgb_multicast == nullptr),
               "gg_nvls_exchange: both buffers or neither must have a multicast mapping");
    ExchangeArgs a{};
    a.rank = rank; a.world = world; a.parts = parts;
    a.bucket_mc = bucket_multicast; a.rgb_mc = rgb_multicast;
    a.bucket_vec4 = bucket_floats / 4; a.slot_vec4 = rgb_slot_floats / 4;
    for (int p = 0; p < world; ++p) {
        GG_REQUIRE(bucket_peers[p] && rgb_peers[p], "gg_nvls_exchange: null peer address");
        GG_REQUIRE(((uintptr_t)bucket_peers[p] & 15) == 0 && ((uintptr_t)rgb_peers[p] & 15) == 0,
                   "gg_nvls_exchange: peer buffers must be 16-byte aligned");
        a.bucket_peer[p] = bucket_peers[p];
        a.rgb_peer[p] = rgb_peers[p];
    }
    GG_REQUIRE(((uintptr_t)bucket_multicast & 15) == 0 && ((uintptr_t)rgb_multicast & 15) == 0,
               "gg_nvls_exchange: multicast addresses must be 16-byte aligned");
    if (a.bucket_vec4 == 0 && a.slot_vec4 == 0) return GG_OK;
    // a copy-shaped kernel: enough CTAs to keep every NVLink port busy; one CTA per SM when the two halves run as
    // two concurrent launches, two when one launch does both
    int blocks = parts == 3 ? 148 * 2 : 148;
    if (const char* e = getenv("GG_NVLS_BLOCKS")) {   // tuning knob (profiles/r02_exchange_tuning.txt)
        const int v = atoi(e);
        if (v >= 1 && v <= 148 * 16) blocks = v;
    }
    // mode 0: multimem.ld_reduce + multimem.st, 1: peer loads and stores, 2: peer loads + multimem.st.
    // Default: peer loads and stores -- measured fastest at 2 ranks (0.150 vs 0.229 ms) and, for config 1's 54 MB
    // bucket, at 8 ranks (0.294 vs 0.312 / 0.323 ms), profiles/r02_exchange_tuning.txt.  The peer form moves
    // 2 (W-1)/W of the bucket per GPU and direction, the multimem form 1x (the switch adds and replicates): with
    // many ranks and a large bucket the bytes decide -- configs[3]'s 108 MB bucket at 8 ranks took 0.84 ms through
    // multimem and 1.09 ms through peer loads -- so reductions of >= 80 MB over >= 4 ranks go through the switch.
    // GG_NVLS_MODE overrides (the multimem forms need a multicast mapping).
    int mode = 1;
    if (bucket_multicast && (parts & 2) && world >= 4 && bucket_floats * 4ll >= (80ll << 20)) mode = 0;
    if (const char* e = getenv("GG_NVLS_MODE")) {
        const int v = atoi(e);
        if (v == 1 || (bucket_multicast && (v == 0 || v == 2))) mode = v;
    }
    if (mode == 0)
        nvls_exchange_kernel<0><<<blocks, 512, 0, (cudaStream_t)stream>>>(a);
    else if (mode == 2)
        nvls_exchange_kernel<2><<<blocks, 512, 0, (cudaStream_t)stream>>>(a);
    else
        nvls_exchange_kernel<1><<<blocks, 512, 0, (cudaStream_t)stream>>>(a);
    count_launch();
    return check_launch("nvls_exchange_kernel");
}
