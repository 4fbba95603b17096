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

extern "C" int gg_version(void) { return 100; }

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
