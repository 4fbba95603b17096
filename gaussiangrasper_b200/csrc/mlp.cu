// The CLIP up-projection of the reference model applied to a whole feature map:
//   fea_up = MLP(32 -> 128 -> 512), Linear - ReLU - Linear  (nerfstudio/models/gaussian_splatting.py:198-213, :294)
//   render.sh applies it to every pixel of every rendered view (nerfstudio/pipelines/base_pipeline.py:408):
//   42.8 GFLOP per 640x480 frame -- the one dense contraction next to the rasterizer path.
//
// One fused kernel on the 5th-generation tensor cores (tcgen05.mma, kind::tf32, accumulators in TMEM):
//   * fp32-equivalent accuracy by the 3xTF32 split: x = hi + lo (both TF32), x.w ~ hi.hi + hi.lo + lo.hi, fp32
//     accumulation in TMEM; the dropped lo.lo term is 2^-22 relative.
//   * per 128-pixel tile: X [128x32] is split and written to shared memory in the UMMA canonical K-major layout,
//     layer 1 is 12 MMAs (M128 N128 K8, A and B from shared memory) into TMEM columns 0..127; the epilogue warps read
//     the accumulator rows with tcgen05.ld, add the bias, apply ReLU, split again and write H [128x128] (hi, lo)
//     back into TMEM with tcgen05.st: layer 2 takes its A operand FROM TMEM (a row is a lane, K runs along the
//     columns -- exactly what a thread-per-row epilogue produces), so the hidden activations never touch shared
//     memory, let alone HBM;
//   * layer 2 runs in 16 column chunks of 32 outputs: the packed (pre-split, pre-laid-out) W2 chunk arrives by one
//     bulk copy (cp.async.bulk, TMA engine) into a 4-deep ring, 48 MMAs (M128 N32 K8) fill one of four TMEM
//     accumulators while the epilogue warps drain the others (bias add, 128-byte row stores);
//   * warp-specialised, persistent (one CTA per SM): warps 0-3 epilogue, warp 4 one MMA-issuing thread, warp 5 stages
//     the next tile's X and feeds the ring; everything is handed over through mbarriers (tcgen05.commit for the
//     tensor-core side), there is no CTA-wide barrier in the tile loop;
//   * the output [P, 512] fp32 is written exactly once (629 MB per 640x480 frame: the HBM floor of the op).
// Shared memory: X 32 KB + W1 32 KB (resident) + W2 ring 128 KB; TMEM: all 512 columns.
#include "gg_common.cuh"
#include "gg_tma.cuh"
#include "gg_b200.h"

namespace gg {

constexpr int kMlpIn = 32, kMlpHid = 128, kMlpOut = 512;
constexpr int kTileM = 128;                 // pixels per tile = UMMA M
constexpr int kChunkN = 32;                 // layer-2 outputs per chunk = UMMA N of layer 2
constexpr int kChunks = kMlpOut / kChunkN;  // 16
constexpr int kRing = 4;                    // W2 chunk slots in shared memory == layer-2 accumulators in TMEM
constexpr int kSlabA = kTileM * 16;         // bytes of one K-core-column (4 tf32) of a 128-row operand
constexpr int kSlabW2 = kChunkN * 16;       // ... of a 32-row W2 chunk
constexpr int kXBytes = (kMlpIn / 4) * kSlabA;               // 16 KB per half (hi / lo)
constexpr int kW1Bytes = (kMlpIn / 4) * (kMlpHid * 16);      // 16 KB per half
constexpr int kW2ChunkBytes = 2 * (kMlpHid / 4) * kSlabW2;   // 32 KB (hi + lo)
constexpr int kMlpSmem = 2 * kXBytes + 2 * kW1Bytes + kRing * kW2ChunkBytes + (kMlpHid + kMlpOut) * 4 + 256 + 1024;
// TMEM columns: layer-1 accumulator | H hi | H lo (the A operand of layer 2 lives in TMEM) | layer-2 accumulators
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kD1Col = 0, kHhiCol = 128, kHloCol = 256, kD2Col = 384;
constexpr int kEpiThreads = 128, kMlpThreads = 192;   // warps 0-3 epilogue, warp 4 MMA issue, warp 5 loads / X staging

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// UMMA shared-memory descriptor, K-major, no swizzle: 8-row x 16-byte core matrices; SBO = distance between 8-row
// groups, LBO = distance between the two core matrices of one K step (cute/arch/mma_sm100_desc.hpp SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

// instruction descriptor: D = F32, A = B = TF32, both K-major, dense (InstrDescriptor of the same header)
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] . B[smem]
__device__ __forceinline__ void umma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// D[tmem] (+)= A[tmem] . B[smem]: A rows are TMEM lanes, its K elements consecutive 32-bit columns
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// a wait that cannot hang the GPU: a barrier that never completes is a bug in this file, reported as a trap
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, unsigned parity) {
    for (unsigned spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 28)) __trap();
}

#define GG_R32(r) \
    r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8], r[9], r[10], r[11], r[12], r[13], r[14], r[15], r[16], r[17], \
        r[18], r[19], r[20], r[21], r[22], r[23], r[24], r[25], r[26], r[27], r[28], r[29], r[30], r[31]

// 32 consecutive fp32 accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 consecutive 32-bit columns of this thread's TMEM lane <- registers
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(v[i]);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}

// Weights -> the two TF32 halves in the exact shared-memory layout of the kernel (done once per weight set):
//   w1p [2][8 slabs][128 rows][4]                 hi / lo of W1 [128, 32]
//   w2p [16 chunks][2][32 slabs][32 rows][4]      hi / lo of W2 [512, 128], one contiguous 32 KB block per chunk
__global__ void __launch_bounds__(256)
mlp_pack_kernel(const float* __restrict__ w1, const float* __restrict__ w2, float* __restrict__ w1p, float* __restrict__ w2p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < kMlpHid * kMlpIn) {
        const int n = i / kMlpIn, k = i - n * kMlpIn;
        const float x = w1[i], hi = to_tf32(x), lo = to_tf32(x - hi);
        const int o = ((k >> 2) * kMlpHid + n) * 4 + (k & 3);
        w1p[o] = hi;
        w1p[(kMlpIn / 4) * kMlpHid * 4 + o] = lo;
    }
    if (i < kMlpOut * kMlpHid) {
        const int n = i / kMlpHid, k = i - n * kMlpHid;
        const float x = w2[i], hi = to_tf32(x), lo = to_tf32(x - hi);
        const int c = n / kChunkN, nn = n - c * kChunkN;
        const int half = (kMlpHid / 4) * kChunkN * 4;   // floats of one half of a chunk
        const int o = c * 2 * half + ((k >> 2) * kChunkN + nn) * 4 + (k & 3);
        w2p[o] = hi;
        w2p[o + half] = lo;
    }
}

// Barrier slots (shared memory, 64-bit each)
enum : int {
    kBarW1 = 0,                 // W1 has landed (once per CTA)
    kBarXFull = 1,              // the tile's X halves are staged                      (32 arrivals: the staging warp)
    kBarXEmpty = 2,             // layer 1 of the tile has consumed X                   (tcgen05.commit)
    kBarD1Full = 3,             // layer-1 accumulator complete                          (tcgen05.commit)
    kBarD1Empty = 4,            // ... and read by every epilogue thread               (128 arrivals)
    kBarHFull = 5,              // H hi / lo written to TMEM                            (128 arrivals)
    kBarW2Full = 6,             // [kRing] chunk weights landed                          (bulk-copy bytes)
    kBarW2Empty = 6 + kRing,    // [kRing] chunk weights consumed                        (tcgen05.commit)
    kBarD2Full = 6 + 2 * kRing, // [kRing] layer-2 accumulator complete                  (tcgen05.commit)
    kBarD2Empty = 6 + 3 * kRing,// [kRing] ... and drained by every epilogue thread     (128 arrivals)
    kNumBars = 6 + 4 * kRing
};

__global__ void __launch_bounds__(kMlpThreads, 1)
mlp_up_kernel(long long n_rows, const float* __restrict__ x, long long x_stride, const float* __restrict__ w1p,
              const float* __restrict__ b1, const float* __restrict__ w2p, const float* __restrict__ b2,
              float* __restrict__ y) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* x_hi = base;
    unsigned char* x_lo = base + kXBytes;
    unsigned char* w1_s = base + 2 * kXBytes;                       // hi then lo, as packed
    unsigned char* ring = base + 2 * kXBytes + 2 * kW1Bytes;        // kRing slots of {W2 chunk hi, lo}
    float* b1_s = reinterpret_cast<float*>(ring + kRing * kW2ChunkBytes);
    float* b2_s = b1_s + kMlpHid;
    uint64_t* bars = reinterpret_cast<uint64_t*>(b2_s + kMlpOut);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < kMlpHid; i += kMlpThreads) b1_s[i] = b1[i];
    for (int i = tid; i < kMlpOut; i += kMlpThreads) b2_s[i] = b2[i];
    if (tid == 0) {
        mbar_init(bars + kBarW1, 1);
        mbar_init(bars + kBarXFull, 32);
        mbar_init(bars + kBarXEmpty, 1);
        mbar_init(bars + kBarD1Full, 1);
        mbar_init(bars + kBarD1Empty, kEpiThreads);
        mbar_init(bars + kBarHFull, kEpiThreads);
        for (int i = 0; i < kRing; ++i) {
            mbar_init(bars + kBarW2Full + i, 1);
            mbar_init(bars + kBarW2Empty + i, 1);
            mbar_init(bars + kBarD2Full + i, 1);
            mbar_init(bars + kBarD2Empty + i, kEpiThreads);
        }
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const long long n_tiles = (n_rows + kTileM - 1) / kTileM;
    // A "fresh" barrier passes a wait on parity 1: the first round of every *Empty barrier is free.

    if (warp < 4) {
        // ===================== epilogue warps: thread = row = TMEM lane ==========================================
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        unsigned it = 0;          // tiles done by this CTA
        unsigned g = 0;           // chunks done by this CTA (ring position)
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const long long row = tile * kTileM + tid;
            // ---- epilogue 1: bias + ReLU, split, H hi / lo back into TMEM as the A operand of layer 2 --------------
            mbar_wait_bounded(bars + kBarD1Full, it & 1);
            tc_fence_after();
#pragma unroll 1
            for (int cb = 0; cb < kMlpHid / 32; ++cb) {
                float v[32], hi[32], lo[32];
                tmem_ld32(tmem + lane_base + kD1Col + cb * 32, v);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float h = fmaxf(v[j] + b1_s[cb * 32 + j], 0.0f);
                    hi[j] = to_tf32(h);
                    lo[j] = to_tf32(h - hi[j]);
                }
                tmem_st32(tmem + lane_base + kHhiCol + cb * 32, hi);
                tmem_st32(tmem + lane_base + kHloCol + cb * 32, lo);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
            tc_fence_before();
            mbar_arrive(bars + kBarD1Empty);
            mbar_arrive(bars + kBarHFull);
            // ---- epilogue 2: 16 chunks of 32 outputs: bias add, one 128-byte row segment per thread ----------------
#pragma unroll 1
            for (int c = 0; c < kChunks; ++c, ++g) {
                const int slot = g % kRing;
                mbar_wait_bounded(bars + kBarD2Full + slot, (g / kRing) & 1);
                tc_fence_after();
                float v[32];
                tmem_ld32(tmem + lane_base + kD2Col + slot * kChunkN, v);
                tc_fence_before();
                mbar_arrive(bars + kBarD2Empty + slot);      // the accumulator may be overwritten
                if (row < n_rows) {
                    float4* dst = reinterpret_cast<float4*>(y + row * kMlpOut + c * kChunkN);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        float4 o;
                        o.x = v[4 * q] + b2_s[c * kChunkN + 4 * q];
                        o.y = v[4 * q + 1] + b2_s[c * kChunkN + 4 * q + 1];
                        o.z = v[4 * q + 2] + b2_s[c * kChunkN + 4 * q + 2];
                        o.w = v[4 * q + 3] + b2_s[c * kChunkN + 4 * q + 3];
                        __stcs(dst + q, o);   // streaming store: the 629 MB output must not evict the weights from L2
                    }
                }
            }
        }
    } else if (warp == 4) {
        // ===================== MMA issuer: one thread =============================================================
        if (lane == 0) {
            constexpr uint32_t kIdesc1 = umma_idesc_tf32(kTileM, kMlpHid);
            constexpr uint32_t kIdesc2 = umma_idesc_tf32(kTileM, kChunkN);
            mbar_wait_bounded(bars + kBarW1, 0);
            unsigned it = 0, g = 0;
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                // ---- layer 1: D1 = Xhi W1hi + Xhi W1lo + Xlo W1hi ------------------------------------------------
                mbar_wait_bounded(bars + kBarXFull, it & 1);
                mbar_wait_bounded(bars + kBarD1Empty, (it & 1) ^ 1);
                tc_fence_after();
                {
                    const uint32_t a_addr[3] = {smem_u32(x_hi), smem_u32(x_hi), smem_u32(x_lo)};
                    const uint32_t b_addr[3] = {smem_u32(w1_s), smem_u32(w1_s) + kW1Bytes, smem_u32(w1_s)};
                    uint32_t acc = 0;
#pragma unroll
                    for (int t = 0; t < 3; ++t)
#pragma unroll
                        for (int s = 0; s < kMlpIn / 8; ++s) {
                            umma_tf32_ss(tmem + kD1Col, umma_desc(a_addr[t] + s * 2 * kSlabA, kSlabA, 128),
                                         umma_desc(b_addr[t] + s * 2 * (kMlpHid * 16), kMlpHid * 16, 128), kIdesc1, acc);
                            acc = 1;
                        }
                }
                umma_commit(bars + kBarD1Full);
                umma_commit(bars + kBarXEmpty);
                // ---- layer 2: 16 chunks, A = H from TMEM, B = the ring slot ---------------------------------------
                mbar_wait_bounded(bars + kBarHFull, it & 1);
#pragma unroll 1
                for (int c = 0; c < kChunks; ++c, ++g) {
                    const int slot = g % kRing;
                    const unsigned ph = (g / kRing) & 1;
                    mbar_wait_bounded(bars + kBarW2Full + slot, ph);
                    mbar_wait_bounded(bars + kBarD2Empty + slot, ph ^ 1);
                    tc_fence_after();
                    const uint32_t w_hi = smem_u32(ring) + slot * kW2ChunkBytes, w_lo = w_hi + kW2ChunkBytes / 2;
                    const uint32_t d = tmem + kD2Col + slot * kChunkN;
                    const uint64_t bd_hi = umma_desc(w_hi, kSlabW2, 128), bd_lo = umma_desc(w_lo, kSlabW2, 128);
                    constexpr uint64_t kStep = (2 * kSlabW2) >> 4;   // descriptor address field advances by one K step
#pragma unroll
                    for (int s = 0; s < kMlpHid / 8; ++s) {
                        umma_tf32_ts(d, tmem + kHhiCol + 8 * s, bd_hi + s * kStep, kIdesc2, s > 0);
                        umma_tf32_ts(d, tmem + kHhiCol + 8 * s, bd_lo + s * kStep, kIdesc2, 1);
                        umma_tf32_ts(d, tmem + kHloCol + 8 * s, bd_hi + s * kStep, kIdesc2, 1);
                    }
                    umma_commit(bars + kBarD2Full + slot);
                    umma_commit(bars + kBarW2Empty + slot);
                }
            }
        }
    } else {
        // ===================== staging warp: X split (all lanes), W1 / W2 bulk copies (lane 0) =====================
        auto stage_x = [&](long long tile, unsigned it) {
            mbar_wait_bounded(bars + kBarXEmpty, (it & 1) ^ 1);
#pragma unroll 1
            for (int r4 = 0; r4 < 4; ++r4) {
                const int r = lane + 32 * r4;
                const long long row = tile * kTileM + r;
                const float* src = x + row * x_stride;
#pragma unroll
                for (int s = 0; s < kMlpIn / 4; ++s) {
                    float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
                    if (row < n_rows) { v0 = __ldg(src + 4 * s); v1 = __ldg(src + 4 * s + 1); v2 = __ldg(src + 4 * s + 2); v3 = __ldg(src + 4 * s + 3); }
                    float4 hi, lo;
                    hi.x = to_tf32(v0); lo.x = to_tf32(v0 - hi.x);
                    hi.y = to_tf32(v1); lo.y = to_tf32(v1 - hi.y);
                    hi.z = to_tf32(v2); lo.z = to_tf32(v2 - hi.z);
                    hi.w = to_tf32(v3); lo.w = to_tf32(v3 - hi.w);
                    *reinterpret_cast<float4*>(x_hi + s * kSlabA + r * 16) = hi;
                    *reinterpret_cast<float4*>(x_lo + s * kSlabA + r * 16) = lo;
                }
            }
            proxy_fence();
            mbar_arrive(bars + kBarXFull);
        };
        if (lane == 0) bulk_load(w1_s, w1p, 2 * kW1Bytes, bars + kBarW1);
        unsigned it = 0, g = 0;
        long long tile = blockIdx.x;
        if (tile < n_tiles) stage_x(tile, 0);
        for (; tile < n_tiles; tile += gridDim.x, ++it) {
#pragma unroll 1
            for (int c = 0; c < kChunks; ++c, ++g) {
                if (c == kRing && tile + gridDim.x < n_tiles) stage_x(tile + gridDim.x, it + 1);   // next tile's X, early
                if (lane == 0) {
                    const int slot = g % kRing;
                    mbar_wait_bounded(bars + kBarW2Empty + slot, ((g / kRing) & 1) ^ 1);
                    bulk_load(ring + slot * kW2ChunkBytes, reinterpret_cast<const unsigned char*>(w2p) + (size_t)c * kW2ChunkBytes,
                              kW2ChunkBytes, bars + kBarW2Full + slot);
                }
                __syncwarp();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(kTmemCols));
}

}  // namespace gg

using namespace gg;

extern "C" size_t gg_mlp_packed_floats(void) {
    return (size_t)2 * kMlpHid * kMlpIn + (size_t)2 * kMlpOut * kMlpHid;
}

extern "C" int gg_mlp_pack_weights(const float* w1, const float* w2, float* packed, void* stream) {
    GG_REQUIRE(w1 && w2 && packed, "gg_mlp_pack_weights: null pointer");
    GG_REQUIRE(((uintptr_t)packed & 15) == 0, "gg_mlp_pack_weights: packed buffer must be 16-byte aligned");
    mlp_pack_kernel<<<div_up(kMlpOut * kMlpHid, 256), 256, 0, (cudaStream_t)stream>>>(w1, w2, packed, packed + 2 * kMlpHid * kMlpIn);
    count_launch();
    return check_launch("mlp_pack_kernel");
}

extern "C" int gg_mlp_up(long long n_rows, const float* x, long long x_stride, const float* packed, const float* b1,
                         const float* b2, float* y, void* stream) {
    GG_REQUIRE(n_rows >= 0 && x_stride >= kMlpIn, "gg_mlp_up: bad sizes");
    if (n_rows == 0) return GG_OK;
    GG_REQUIRE(x && packed && b1 && b2 && y, "gg_mlp_up: null pointer");
    GG_REQUIRE(((uintptr_t)packed & 15) == 0 && ((uintptr_t)y & 15) == 0, "gg_mlp_up: packed weights / output must be 16-byte aligned");
    static bool configured = false;
    if (!configured) {
        GG_CUDA(cudaFuncSetAttribute(mlp_up_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMlpSmem));
        configured = true;
    }
    const long long n_tiles = (n_rows + kTileM - 1) / kTileM;
    const int blocks = (int)(n_tiles < 148 ? n_tiles : 148);   // persistent: one CTA per SM
    mlp_up_kernel<<<blocks, kMlpThreads, kMlpSmem, (cudaStream_t)stream>>>(n_rows, x, x_stride, packed, b1,
                                                                          packed + 2 * kMlpHid * kMlpIn, b2, y);
    count_launch();
    return check_launch("mlp_up_kernel");
}
