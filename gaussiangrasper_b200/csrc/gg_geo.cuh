// tau / rcut2 of the packed geo record, shared by blend.cu (gg_pack_geo) and fused.cu.
#pragma once
#include "gg_common.cuh"

namespace gg {

// tau   : largest sigma at which alpha = o*exp(-sigma) can still reach 1/255 (plus a margin);
//         -1 when the Gaussian can never contribute.
// rcut2 : conservative squared pixel distance beyond which sigma > tau:
//         sigma = d^T Q d / 2 >= lambda_min(Q) |d|^2 / 2  =>  |d|^2 <= 2 tau / lambda_min(Q).
//         +inf-like when Q is not safely positive definite (never cull), -1 when tau < 0.
__device__ __forceinline__ void geo_tau_rcut(float o, float A, float B, float C, float& tau, float& rcut2) {
    if (!(o * 255.0f > 1.0f)) { tau = -1.0f; rcut2 = -1.0f; return; }
    tau = __logf(o * 255.0f) + GG_TAU_MARGIN;
    const float mid = 0.5f * (A + C);
    const float det = A * C - B * B;
    const float lmin = mid - sqrtf(fmaxf(mid * mid - det, 0.0f));
    rcut2 = (lmin > 1e-12f) ? 2.0f * tau / lmin * 1.01f + 0.01f : 3.0e38f;
}

}  // namespace gg
