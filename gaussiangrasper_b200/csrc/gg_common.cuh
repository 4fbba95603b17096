// Shared host-side plumbing for the gg_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define GG_TILE 16
// slack added to tau = ln(255*opacity): pairs with sigma > tau are skipped before the exp
#define GG_TAU_MARGIN 1e-3f
#define GG_OK 0
#define GG_ERR_ARG (-1)
#define GG_ERR_CHANNELS (-2)
#define GG_ERR_WORKSPACE (-3)

namespace gg {

// error string + launch accounting live in capi.cu
void set_error(const char* msg);
int cuda_fail(cudaError_t e, const char* where);
void count_launch(int n = 1);

inline int check_launch(const char* where) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, where);
    return GG_OK;
}

static inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace gg

#define GG_REQUIRE(cond, msg)                   \
    do {                                        \
        if (!(cond)) {                          \
            gg::set_error(msg);                 \
            return GG_ERR_ARG;                  \
        }                                       \
    } while (0)

#define GG_CUDA(call)                                          \
    do {                                                       \
        cudaError_t e__ = (call);                              \
        if (e__ != cudaSuccess) return gg::cuda_fail(e__, #call); \
    } while (0)
