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

// erfc(z) for z >= 0: t*exp(-z^2 + P(t)), t = 1/(1+z/2) (Chebyshev fit, fractional error ~1e-7 in exact
// arithmetic, ~2e-6 in fp32).  Used for the exact-erf GELU of the dds network: Phi(x) = 0.5 erfc(-x/sqrt 2).
// In fp32 the reference's own formula x*0.5*(1+erf(x/sqrt 2)) carries a 7e-7 absolute rounding error
// (cancellation in 1+erf for x<0); this form is within 6e-7 absolute of the exact GELU (tools/erf_accuracy.py)
// at a third of the instruction count of erff().  -DCMCD_EXACT_ERF switches back to erff/expf.
__device__ __forceinline__ float erfc_pos(float z) {
    const float t = __fdividef(1.0f, fmaf(0.5f, z, 1.0f));
    float p = 0.17087277f;
    p = fmaf(p, t, -0.82215223f);
    p = fmaf(p, t, 1.48851587f);
    p = fmaf(p, t, -1.13520398f);
    p = fmaf(p, t, 0.27886807f);
    p = fmaf(p, t, -0.18628806f);
    p = fmaf(p, t, 0.09678418f);
    p = fmaf(p, t, 0.37409196f);
    p = fmaf(p, t, 1.00002368f);
    const float arg = fmaf(t, p, fmaf(-z, z, -1.26551223f));
    return t * exp2f(arg * 1.4426950408889634f);
}

// standard normal CDF
__device__ __forceinline__ float norm_cdf(float x) {
#ifdef CMCD_EXACT_ERF
    return 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
#else
    const float e = 0.5f * erfc_pos(fabsf(x) * 0.70710678118654752440f);
    return x < 0.f ? e : 1.0f - e;
#endif
}

template <int ACT>
__device__ __forceinline__ float act_fwd(float x) {
    if constexpr (ACT == ACT_SOFTPLUS) {
        return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
    } else {
        return x * norm_cdf(x);
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
        const float cdf = norm_cdf(x);
        a = x * cdf;
        da = fmaf(x * 0.3989422804014327f, exp2f(-0.72134752044448170f * x * x), cdf);
    }
}

// ---- tensor-core path activations -------------------------------------------------------------------------
// Exact-erf GELU via Abramowitz-Stegun 7.1.26: 0.5 erfc(|x|/sqrt 2) = t (a1 + t (a2 + ... a5 t)) exp(-x^2/2),
// t = 1/(1 + p |x|/sqrt 2); |error| <= 7.5e-8 in exact arithmetic, 4.7e-7 on gelu in fp32 (the reference's own
// fp32 formula x*0.5*(1+erf(x/sqrt 2)) carries 4.5e-7; tools/erf_accuracy.py).  14 instructions, 2 of them MUFU,
// no predicate: gelu(x) = 0.5 x + |x| (0.5 - h).
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float gelu_half_erfc(float x, float& e_out) {
    const float u = fabsf(x) * 0.84932180028801907f;               // u^2 = (x^2/2) log2 e
    const float t = rcp_ftz(fmaf(u, 0.27273748088f, 1.0f));        // p / sqrt(log2 e), p = 0.3275911 (A&S 7.1.26)
    const float e = ex2_ftz(-u * u);                               // exp(-x^2/2); flushes below 2^-126 (|x| > 13.2)
    float q = 0.5307027145f;
    q = fmaf(q, t, -0.7265760135f);
    q = fmaf(q, t, 0.7107068705f);
    q = fmaf(q, t, -0.142248368f);
    q = fmaf(q, t, 0.127414796f);
    e_out = e;
    return (q * t) * e;                                            // h = 0.5 erfc(|x|/sqrt 2) in (0, 0.5]
}
__device__ __forceinline__ float gelu_fast(float x) {
    float e;
    const float h = gelu_half_erfc(x, e);
    return fmaf(fabsf(x), 0.5f - h, 0.5f * x);
}
__device__ __forceinline__ void gelu_fast_grad(float x, float& a, float& da) {
    float e;
    const float h = gelu_half_erfc(x, e);
    const float w = 0.5f - h;                                      // >= 0
    a = fmaf(fabsf(x), w, 0.5f * x);
    const float cdf = 0.5f + copysignf(w, x);
    da = fmaf(x * 0.3989422804014327f, e, cdf);
}
template <int ACT>
__device__ __forceinline__ float act_tc(float x) {
    if constexpr (ACT == ACT_GELU) return gelu_fast(x);
    else return act_fwd<ACT>(x);
}
template <int ACT>
__device__ __forceinline__ void act_tc_grad(float x, float& a, float& da) {
    if constexpr (ACT == ACT_GELU) gelu_fast_grad(x, a, da);
    else act_fwd_grad<ACT>(x, a, da);
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
