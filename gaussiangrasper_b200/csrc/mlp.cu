// The CLIP up-projection of the reference model applied to a whole feature map:
//   fea_up = MLP(32 -> 128 -> 512), Linear - ReLU - Linear  (nerfstudio/models/gaussian_splatting.py:198-213, :294)
//   render.sh applies it to every pixel of every rendered view (nerfstudio/pipelines/base_pipeline.py:408):
//   42.8 GFLOP per 640x480 frame -- the one dense contraction next to the rasterizer path.
//
// One fused kernel on the 5th-generation tensor cores (tcgen05.mma, kind::tf32, accumulators in TMEM):
//   * fp32-equivalent accuracy by the 3xTF32 split: x = hi + lo (both TF32), x.w ~ hi.hi + hi.lo + lo.hi, fp32
//     accumulation in TMEM; the dropped lo.lo term is 2^-22 relative.
//   * per 128-pixel tile: X [128x32] is split and written to shared memory in the UMMA canonical K-major layout,
//     layer 1 is 12 MMAs (M128 N128 K8) into TMEM columns 0..127; the epilogue reads the accumulator rows with
//     tcgen05.ld, adds the bias, applies ReLU, splits again and writes H [128x128] (hi, lo) back to shared memory as
//     the A operand of layer 2 -- the hidden activations never leave the SM;
//   * layer 2 runs in 16 column chunks of 32 outputs: the packed (pre-split, pre-laid-out) W2 chunk arrives by one
//     bulk copy (cp.async.bulk, TMA engine) into a 2-deep ring, 48 MMAs (M128 N32 K8) fill one of two TMEM
//     accumulators while the four warps drain the other one (bias add, 128-byte row stores): tensor cores, TMA and
//     the store stream overlap without warp specialisation.
//   * the output [P, 512] fp32 is written exactly once (629 MB per 640x480 frame: the HBM floor of the op).
// Shared memory: H 128 KB + 64 KB that holds {X hi/lo, W1 hi/lo} during layer 1 and the W2 ring afterwards.
#include "gg_common.cuh"
#include "gg_tma.cuh"
#include "gg_b200.h"

namespace gg {

constexpr int kMlpIn = 32, kMlpHid = 128, kMlpOut = 512;
constexpr int kTileM = 128;                 // pixels per tile = UMMA M
constexpr int kChunkN = 32;                 // layer-2 outputs per chunk = UMMA N of layer 2
constexpr int kChunks = kMlpOut / kChunkN;  // 16
constexpr int kSlabA = kTileM * 16;         // bytes of one K-core-column (4 tf32) of a 128-row operand
constexpr int kSlabW2 = kChunkN * 16;       // ... of a 32-row W2 chunk
constexpr int kHBytes = (kMlpHid / 4) * kSlabA;              // 64 KB per half (hi / lo)
constexpr int kXBytes = (kMlpIn / 4) * kSlabA;               // 16 KB per half
constexpr int kW1Bytes = (kMlpIn / 4) * (kMlpHid * 16);      // 16 KB per half
constexpr int kW2ChunkBytes = 2 * (kMlpHid / 4) * kSlabW2;   // 32 KB (hi + lo)
constexpr int kRegionR = 2 * kXBytes + 2 * kW1Bytes;         // 64 KB == 2 * kW2ChunkBytes
static_assert(kRegionR == 2 * kW2ChunkBytes, "the W2 ring aliases the layer-1 operands exactly");
constexpr int kMlpSmem = 2 * kHBytes + kRegionR + (kMlpHid + kMlpOut) * 4 + 128 + 1024;
constexpr uint32_t kTmemCols = 256;         // D1: 128 columns, D2: 2 x 32 columns
constexpr uint32_t kD2Col = 128;

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

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// a wait that cannot hang the GPU: a barrier that never completes is a bug in this file, reported as a trap
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, unsigned parity) {
    for (unsigned spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 28)) __trap();
}

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

__global__ void __launch_bounds__(128, 1)
mlp_up_kernel(long long n_rows, const float* __restrict__ x, long long x_stride, const float* __restrict__ w1p,
              const float* __restrict__ b1, const float* __restrict__ w2p, const float* __restrict__ b2,
              float* __restrict__ y, int relu_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* h_hi = base;
    unsigned char* h_lo = base + kHBytes;
    unsigned char* region = base + 2 * kHBytes;   // {X hi, X lo, W1 hi, W1 lo} during layer 1, then the W2 ring
    unsigned char* x_hi = region;
    unsigned char* x_lo = region + kXBytes;
    unsigned char* w1_s = region + 2 * kXBytes;   // hi then lo, as packed
    float* b1_s = reinterpret_cast<float*>(region + kRegionR);
    float* b2_s = b1_s + kMlpHid;
    uint64_t* bars = reinterpret_cast<uint64_t*>(b2_s + kMlpOut);   // [0] w1, [1] mma1, [2..3] w2 ring, [4..5] mma2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < kMlpHid; i += 128) b1_s[i] = b1[i];
    for (int i = tid; i < kMlpOut; i += 128) b2_s[i] = b2[i];
    if (tid == 0) {
        for (int i = 0; i < 6; ++i) mbar_init(bars + i, 1);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;   // a warp reads the TMEM lanes of its quarter

    constexpr uint32_t kIdesc1 = umma_idesc_tf32(kTileM, kMlpHid);
    constexpr uint32_t kIdesc2 = umma_idesc_tf32(kTileM, kChunkN);
    unsigned ph_w1 = 0, ph_mma1 = 0, ph_w2[2] = {0, 0}, ph_mma2[2] = {0, 0};
    const long long n_tiles = (n_rows + kTileM - 1) / kTileM;

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long row = tile * kTileM + tid;
        // ---- stage layer 1: W1 by bulk copy, X split into hi / lo in the canonical layout ----------------------
        if (tid == 0) bulk_load(w1_s, w1p, 2 * kW1Bytes, bars + 0);
        {
            float xv[kMlpIn];
            if (row < n_rows) {
                const float* src = x + row * x_stride;
#pragma unroll
                for (int k = 0; k < kMlpIn; ++k) xv[k] = __ldg(src + k);
            } else {
#pragma unroll
                for (int k = 0; k < kMlpIn; ++k) xv[k] = 0.0f;
            }
#pragma unroll
            for (int s = 0; s < kMlpIn / 4; ++s) {
                float4 hi, lo;
                hi.x = to_tf32(xv[4 * s]); lo.x = to_tf32(xv[4 * s] - hi.x);
                hi.y = to_tf32(xv[4 * s + 1]); lo.y = to_tf32(xv[4 * s + 1] - hi.y);
                hi.z = to_tf32(xv[4 * s + 2]); lo.z = to_tf32(xv[4 * s + 2] - hi.z);
                hi.w = to_tf32(xv[4 * s + 3]); lo.w = to_tf32(xv[4 * s + 3] - hi.w);
                *reinterpret_cast<float4*>(x_hi + s * kSlabA + tid * 16) = hi;
                *reinterpret_cast<float4*>(x_lo + s * kSlabA + tid * 16) = lo;
            }
        }
        proxy_fence();
        __syncthreads();
        // ---- layer 1: D1 = Xhi W1hi + Xhi W1lo + Xlo W1hi ------------------------------------------------------
        if (tid == 0) {
            mbar_wait_bounded(bars + 0, ph_w1);
            tc_fence_after();
            const uint32_t a_addr[3] = {smem_u32(x_hi), smem_u32(x_hi), smem_u32(x_lo)};
            const uint32_t b_addr[3] = {smem_u32(w1_s), smem_u32(w1_s) + kW1Bytes, smem_u32(w1_s)};
            uint32_t acc = 0;
#pragma unroll
            for (int t = 0; t < 3; ++t)
#pragma unroll
                for (int s = 0; s < kMlpIn / 8; ++s) {
                    umma_tf32(tmem, umma_desc(a_addr[t] + s * 2 * kSlabA, kSlabA, 128),
                              umma_desc(b_addr[t] + s * 2 * (kMlpHid * 16), kMlpHid * 16, 128), kIdesc1, acc);
                    acc = 1;
                }
            umma_commit(bars + 1);
        }
        ph_w1 ^= 1;
        mbar_wait_bounded(bars + 1, ph_mma1);
        ph_mma1 ^= 1;
        tc_fence_after();
        // the layer-1 operands are dead: start the W2 ring in their place
        if (tid == 0) {
            bulk_load(region, w2p, kW2ChunkBytes, bars + 2);
            bulk_load(region + kW2ChunkBytes, reinterpret_cast<const unsigned char*>(w2p) + kW2ChunkBytes, kW2ChunkBytes, bars + 3);
        }
        // ---- epilogue 1: bias + ReLU, split, H as the A operand of layer 2 --------------------------------------
#pragma unroll 1
        for (int cb = 0; cb < kMlpHid / 32; ++cb) {
            float v[32];
            tmem_ld32(tmem + lane_base + cb * 32, v);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float4 hi, lo;
                float h;
                h = fmaxf(v[4 * q] + b1_s[cb * 32 + 4 * q], 0.0f); hi.x = to_tf32(h); lo.x = to_tf32(h - hi.x);
                h = fmaxf(v[4 * q + 1] + b1_s[cb * 32 + 4 * q + 1], 0.0f); hi.y = to_tf32(h); lo.y = to_tf32(h - hi.y);
                h = fmaxf(v[4 * q + 2] + b1_s[cb * 32 + 4 * q + 2], 0.0f); hi.z = to_tf32(h); lo.z = to_tf32(h - hi.z);
                h = fmaxf(v[4 * q + 3] + b1_s[cb * 32 + 4 * q + 3], 0.0f); hi.w = to_tf32(h); lo.w = to_tf32(h - hi.w);
                const int slab = cb * 8 + q;
                *reinterpret_cast<float4*>(h_hi + slab * kSlabA + tid * 16) = hi;
                *reinterpret_cast<float4*>(h_lo + slab * kSlabA + tid * 16) = lo;
            }
        }
        proxy_fence();
        tc_fence_before();
        __syncthreads();
        // ---- layer 2, 16 chunks of 32 outputs, MMA of chunk c+1 in flight while chunk c is drained ---------------
        auto issue_chunk = [&](int c) {   // thread 0 only
            const int b = c & 1;
            mbar_wait_bounded(bars + 2 + b, ph_w2[b]);
            ph_w2[b] ^= 1;
            tc_fence_after();
            const uint32_t w_hi = smem_u32(region) + b * kW2ChunkBytes, w_lo = w_hi + kW2ChunkBytes / 2;
            const uint32_t a_addr[3] = {smem_u32(h_hi), smem_u32(h_hi), smem_u32(h_lo)};
            const uint32_t b_addr[3] = {w_hi, w_lo, w_hi};
            const uint32_t d = tmem + kD2Col + b * kChunkN;
            uint32_t acc = 0;
#pragma unroll
            for (int t = 0; t < 3; ++t)
#pragma unroll 4
                for (int s = 0; s < kMlpHid / 8; ++s) {
                    umma_tf32(d, umma_desc(a_addr[t] + s * 2 * kSlabA, kSlabA, 128),
                              umma_desc(b_addr[t] + s * 2 * kSlabW2, kSlabW2, 128), kIdesc2, acc);
                    acc = 1;
                }
            umma_commit(bars + 4 + b);
        };
        if (tid == 0) issue_chunk(0);
#pragma unroll 1
        for (int c = 0; c < kChunks; ++c) {
            const int b = c & 1;
            if (tid == 0 && c + 1 < kChunks) issue_chunk(c + 1);
            mbar_wait_bounded(bars + 4 + b, ph_mma2[b]);
            ph_mma2[b] ^= 1;
            tc_fence_after();
            // chunk c's weights have been consumed: refill its ring slot with chunk c + 2
            if (tid == 0 && c + 2 < kChunks)
                bulk_load(region + b * kW2ChunkBytes, reinterpret_cast<const unsigned char*>(w2p) + (size_t)(c + 2) * kW2ChunkBytes,
                          kW2ChunkBytes, bars + 2 + b);
            float v[32];
            tmem_ld32(tmem + lane_base + kD2Col + b * kChunkN, v);
            if (row < n_rows) {
                float4* dst = reinterpret_cast<float4*>(y + row * kMlpOut + c * kChunkN);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float4 o;
                    o.x = v[4 * q] + b2_s[c * kChunkN + 4 * q];
                    o.y = v[4 * q + 1] + b2_s[c * kChunkN + 4 * q + 1];
                    o.z = v[4 * q + 2] + b2_s[c * kChunkN + 4 * q + 2];
                    o.w = v[4 * q + 3] + b2_s[c * kChunkN + 4 * q + 3];
                    if (relu_out) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    __stcs(dst + q, o);   // streaming store: the 629 MB output must not evict the weights from L2
                }
            }
            tc_fence_before();
            __syncthreads();   // everybody has drained accumulator b: chunk c + 2 may overwrite it
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
    mlp_up_kernel<<<blocks, 128, kMlpSmem, (cudaStream_t)stream>>>(n_rows, x, x_stride, packed, b1,
                                                                  packed + 2 * kMlpHid * kMlpIn, b2, y, 0);
    count_launch();
    return check_launch("mlp_up_kernel");
}
