// Shared device/host helpers for the cmcd_b200 kernels (sm_100a).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "../../include/cmcd_b200.h"
#include "prng.cuh"
#include "targets.cuh"

namespace cmcd {

constexpr int ACT_SOFTPLUS = 1;  // geffner: jax.nn.softplus = logaddexp(x, 0)   (nn.py:45-51)
constexpr int ACT_GELU = 2;      // dds: x*0.5*(1+erf(x/sqrt 2))                  (nn_dds.py:167-176)

template <int ACT>
__device__ __forceinline__ float act_fwd(float x) {
    if constexpr (ACT == ACT_SOFTPLUS) {
        return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
    } else {
        return x * 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    }
}

// activation and its derivative in one go
template <int ACT>
__device__ __forceinline__ void act_fwd_grad(float x, float& a, float& da) {
    if constexpr (ACT == ACT_SOFTPLUS) {
        const float e = expf(-fabsf(x));
        a = fmaxf(x, 0.f) + log1pf(e);
        const float s = 1.0f / (1.0f + e);      // sigmoid(|x|)
        da = x >= 0.f ? s : 1.0f - s;
    } else {
        const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
        a = x * cdf;
        da = cdf + x * 0.3989422804014327f * expf(-0.5f * x * x);
    }
}

// Device-side view of the network (pointers may be shared or global memory).
struct NetView {
    int arch, H, HP, T;
    const float *U1, *U2, *U3, *W2, *W3, *c1, *c2, *c3;
    float out_scale, out_clip;
};

struct BridgeArgs {
    int mode, K;
    long long N;
    float clip_t, clip_q;
    const int32_t* seeds;
    const float *vd_mean, *vd_logdiag, *betas, *eps;
    NetView net;
    TargetDesc tgt;
    float *out_negw, *out_z, *traj;
};

// set the thread-local last-error string (capi.cu)
void set_error(const char* fmt, ...);

#define CMCD_CUDA_OK(expr)                                                                  \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            cmcd::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return 1;                                                                       \
        }                                                                                   \
    } while (0)

}  // namespace cmcd
