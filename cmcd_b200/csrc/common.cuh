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
// ---- packed fp32x2 arithmetic (Blackwell FFMA2: one issue slot, two lanes of the FMA pipe) ------------------------
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pk2(float a, float b) { f32x2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(f32x2_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c) { f32x2_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2_t mul2(f32x2_t a, f32x2_t b) { f32x2_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2_t add2(f32x2_t a, f32x2_t b) { f32x2_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// Two GELUs at once (same formula as gelu_fast, written as gelu(x) = max(x, 0) - |x| h):
// 10 packed FMA-pipe instructions + 4 ALU + 4 MUFU per pair instead of 2 x 14.  E receives exp(-x^2/2) of both lanes.
__device__ __forceinline__ f32x2_t gelu_half_erfc2(float x0, float x1, f32x2_t& NAX, f32x2_t& E) {
    NAX = pk2(__uint_as_float(__float_as_uint(x0) | 0x80000000u), __uint_as_float(__float_as_uint(x1) | 0x80000000u));   // -|x|
    const f32x2_t NU = mul2(NAX, pk2(0.84932180028801907f, 0.84932180028801907f));       // -u
    const f32x2_t DEN = fma2(NU, pk2(-0.27273748088f, -0.27273748088f), pk2(1.0f, 1.0f));  // 1 + p' u
    const f32x2_t U2 = mul2(NU, NU);
    float d0, d1, s0, s1;
    upk2(DEN, d0, d1);
    upk2(U2, s0, s1);
    const f32x2_t T = pk2(rcp_ftz(d0), rcp_ftz(d1));
    E = pk2(ex2_ftz(-s0), ex2_ftz(-s1));
    f32x2_t Q = fma2(pk2(0.5307027145f, 0.5307027145f), T, pk2(-0.7265760135f, -0.7265760135f));
    Q = fma2(Q, T, pk2(0.7107068705f, 0.7107068705f));
    Q = fma2(Q, T, pk2(-0.142248368f, -0.142248368f));
    Q = fma2(Q, T, pk2(0.127414796f, 0.127414796f));
    return mul2(mul2(Q, T), E);   // h
}
__device__ __forceinline__ f32x2_t gelu_fast2(float x0, float x1) {
    f32x2_t NAX, E;
    const f32x2_t H = gelu_half_erfc2(x0, x1, NAX, E);
    return fma2(NAX, H, pk2(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));
}
// value and derivative of two GELUs: A = gelu, DA = Phi(x) + x phi(x)
__device__ __forceinline__ void gelu_fast_grad2(float x0, float x1, f32x2_t& A, f32x2_t& DA) {
#ifdef BT_X_NOGELU
    A = mul2(pk2(x0, x1), pk2(0.5f, 0.5f)); DA = add2(pk2(x0, x1), pk2(0.5f, 0.5f)); return;
#endif
    f32x2_t NAX, E;
    const f32x2_t H = gelu_half_erfc2(x0, x1, NAX, E);
    A = fma2(NAX, H, pk2(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));
    const f32x2_t W = fma2(H, pk2(-1.0f, -1.0f), pk2(0.5f, 0.5f));   // 0.5 - h >= 0
    float w0, w1;
    upk2(W, w0, w1);
    const f32x2_t WS = pk2(__uint_as_float(__float_as_uint(w0) | (__float_as_uint(x0) & 0x80000000u)),
                           __uint_as_float(__float_as_uint(w1) | (__float_as_uint(x1) & 0x80000000u)));   // copysign(w, x)
    const f32x2_t CDF = add2(WS, pk2(0.5f, 0.5f));
    DA = fma2(mul2(pk2(x0, x1), pk2(0.3989422804014327f, 0.3989422804014327f)), E, CDF);
}

// softplus(x) = max(x, 0) + ln2 lg2(1 + e), e = 2^(-|x| log2 e): 6 instructions, 2 of them MUFU, absolute error <= 1.5e-7
// (lg2.approx: 2^-22.6 absolute on [1, 2]) -- the form the block kernels use (blk_net.cuh), a quarter of expf + log1pf.
__device__ __forceinline__ float lg2_ftz(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float softplus_fast(float x) {
    const float e = ex2_ftz(-fabsf(x) * 1.4426950408889634f);
    return fmaf(0.6931471805599453f, lg2_ftz(1.0f + e), fmaxf(x, 0.f));
}
template <int ACT>
__device__ __forceinline__ float act_tc(float x) {
    if constexpr (ACT == ACT_GELU) return gelu_fast(x);
    else return softplus_fast(x);
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
    const float* out_scale_dev;   // optional device scalar overriding out_scale
};
// the output gate of the network: device scalar if the caller provided one (trainable factor_sn), else the by-value field
__device__ __forceinline__ float net_out_scale(const NetView& nv) { return nv.out_scale_dev ? __ldg(nv.out_scale_dev) : nv.out_scale; }

struct BridgeArgs {
    int mode, K;
    long long N;
    float clip_t, clip_q;
    const int32_t* seeds;
    const float *vd_mean, *vd_logdiag, *betas, *eps;
    NetView net;
    TargetDesc tgt;
    float *out_negw, *out_z, *traj;
    // evolve entry (cmcd_bridge_evolve): start from caller-supplied z0[N][d] and per-particle keys[N][2] instead of the integer
    // seed; out_negw then receives +w of the K steps only (no -log q(z0), no log p(z_K)), like mcd_utils.evolve
    const float* z0;
    const uint32_t* keys;
    int tab_tma;   // tensor-core forward kernel: stage the per-step table rows with TMA bulk copies (1) or per-warp cp.async (0)
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
