// Library-wide state of the gg_b200 C ABI: last-error string, launch accounting, device probe.
#include "gg_common.cuh"
#include "gg_b200.h"

#include <atomic>
#include <cstdio>
#include <cstring>

namespace gg {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* msg) {
    std::snprintf(g_err, sizeof(g_err), "%s", msg);
}

int cuda_fail(cudaError_t e, const char* where) {
    std::snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where, cudaGetErrorString(e), cudaGetErrorName(e));
    return (int)e > 0 ? (int)e : 1;
}

void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

}  // namespace gg

extern "C" int gg_version(void) { return GG_ABI_VERSION; }

extern "C" const char* gg_last_error_string(void) { return gg::g_err; }

extern "C" unsigned long long gg_launch_count(void) { return gg::g_launches.load(std::memory_order_relaxed); }

// 0 when a CUDA device of compute capability 10.x is current, otherwise an error code; the Python
// side calls this at import on a GPU box so that a missing / wrong device fails loudly.
extern "C" int gg_check_device(void) {
    int dev = -1;
    GG_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    GG_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) {
        char msg[256];
        std::snprintf(msg, sizeof(msg), "gg_b200 is built for sm_100a only; device %d is sm_%d%d (%s)", dev, prop.major,
                      prop.minor, prop.name);
        gg::set_error(msg);
        return GG_ERR_ARG;
    }
    return GG_OK;
}

// ---------------------------------------------------------------------------------------------
// FP32 FMA throughput probe used by bench.py as the practical roof of the blend kernels
// ("measure an FMA micro-benchmark and use it as the practical roof").
// Each thread runs 8 independent FMA chains; flops = grid * block * iters * 8 * 2.
// ---------------------------------------------------------------------------------------------
namespace gg {
__global__ void __launch_bounds__(256) fma_probe_kernel(int iters, float* out) {
    float a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = 1.0f + 0.001f * (float)(threadIdx.x + k);
    const float m = 0.9999f, c = 1e-4f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = fmaf(a[k], m, c);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    if (s == 123.456f) out[0] = s;  // never true; keeps the chains alive
}
}  // namespace gg

extern "C" int gg_bench_fma(int blocks, int iters, float* out, double* flops, void* stream) {
    GG_REQUIRE(blocks > 0 && iters > 0 && out, "gg_bench_fma: bad arguments");
    gg::fma_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, out);
    gg::count_launch();
    if (flops) *flops = (double)blocks * 256.0 * (double)iters * 16.0;
    return gg::check_launch("fma_probe_kernel");
}
