// Analytic log-density / score / Hessian-vector products of the closed-form targets.
//
// Replaces jax.grad(log_prob_model) (and its reverse-mode transpose) for the targets of
// /root/reference/src/model_handler.py: many_gmm :245-284 (40 diagonal components, the
// "-inf below -1e4" override with zero gradient :279-280), gmm :157-242 (3 full-covariance
// components symmetrised by a coordinate flip = 6 components), funnel :124-154.
// (lgcp :287-409 is d=1600 and lives in the wide path, wide.cu.)
//
// Layout of the packed parameter block `tp` (floats, staged in shared memory by the kernels):
//   mixture targets: ncomp rows of MIX_STRIDE floats
//     many_gmm row: mu0, mu1, -, -, -, -            (shared scale / constants in TargetDesc)
//     gmm row:      m0, m1, p00, p01, p11, logc      (precision matrix P, logc incl. log-weight - log 2)
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

namespace cmcd {

enum TargetKind : int { TGT_GMM = 0, TGT_MANY_GMM = 1, TGT_FUNNEL = 2, TGT_LGCP = 3 };
constexpr int MIX_STRIDE = 6;
constexpr int MIX_MAX = 64;

struct TargetDesc {
    int kind;
    int ncomp;
    float scale;      // many_gmm: component scale (softplus(0.1))
    float inv_var;    // many_gmm: 1/scale^2
    float comp_norm;  // many_gmm: 0.5*log(2 pi) + log(scale)   (per dimension)
    float log_mix;    // many_gmm: -log(ncomp)
    float invalid_below;  // many_gmm: -1e4
    const float* mix;     // device [ncomp][MIX_STRIDE]
};

// ---- 2-D mixtures ---------------------------------------------------------------------------
// Returns log p(z).  g = grad log p (zero where the reference's override kills the gradient).
// If WANT_HVP, also hv = (Hessian of log p) * v.
template <bool WANT_HVP>
__device__ __forceinline__ float mixture2_eval(const TargetDesc& t, const float* __restrict__ tp,
                                               const float (&z)[2], float (&g)[2],
                                               const float (&v)[2], float (&hv)[2]) {
    // Hessian of a log-mixture: H = sum_k r_k (a_k a_k^T + A_k) - g g^T with a_k = grad log N_k, A_k = hess log N_k.
    // Evaluated in the shift-invariant form sum_k r_k (a_k-p)(a_k-p)^T - (g-p)(g-p)^T with the pivot p = a of the
    // dominant component, which avoids the fp32 cancellation of the naive form far away from the modes.
    const int nc = t.ncomp;
    float m = -CUDART_INF_F;
    int kb = 0;
    if (t.kind == TGT_MANY_GMM) {
        float qmin = CUDART_INF_F;
        for (int k = 0; k < nc; ++k) {
            const float d0 = z[0] - tp[k * MIX_STRIDE + 0], d1 = z[1] - tp[k * MIX_STRIDE + 1];
            const float q = fmaf(d0, d0, d1 * d1);       // l_k = -0.5 q / s^2 + const: argmax l = argmin q
            if (q < qmin) { qmin = q; kb = k; }
        }
        const float p0 = -(z[0] - tp[kb * MIX_STRIDE + 0]) * t.inv_var, p1 = -(z[1] - tp[kb * MIX_STRIDE + 1]) * t.inv_var;
        // log N_k - max = -0.5 (q_k - q_min)/s^2 ; the constants (2 comp_norm, log_mix) re-enter in lp below
        const float hl2 = -0.5f * t.inv_var * 1.4426950408889634f;
        m = (-0.5f * qmin * t.inv_var - 2.0f * t.comp_norm) + t.log_mix;
        float S = 0.f, G0 = 0.f, G1 = 0.f, Q0 = 0.f, Q1 = 0.f;
        for (int k = 0; k < nc; ++k) {
            const float d0 = z[0] - tp[k * MIX_STRIDE + 0], d1 = z[1] - tp[k * MIX_STRIDE + 1];
            const float e = exp2f(hl2 * (fmaf(d0, d0, d1 * d1) - qmin));
            const float a0 = -d0 * t.inv_var, a1 = -d1 * t.inv_var;
            S += e; G0 += e * a0; G1 += e * a1;
            if (WANT_HVP) {
                const float c0 = a0 - p0, c1 = a1 - p1;
                const float cv = c0 * v[0] + c1 * v[1];
                Q0 += e * c0 * cv; Q1 += e * c1 * cv;
            }
        }
        const float lp = logf(S) + m;
        const bool valid = lp > t.invalid_below;
        const float inv = 1.0f / S;
        g[0] = valid ? G0 * inv : 0.f;
        g[1] = valid ? G1 * inv : 0.f;
        if (WANT_HVP) {
            const float e0 = G0 * inv - p0, e1 = G1 * inv - p1;
            const float ev = e0 * v[0] + e1 * v[1];
            hv[0] = valid ? (Q0 * inv - v[0] * t.inv_var - e0 * ev) : 0.f;
            hv[1] = valid ? (Q1 * inv - v[1] * t.inv_var - e1 * ev) : 0.f;
        }
        return valid ? lp : -CUDART_INF_F;
    }
    // full-covariance 2-D mixture (gmm)
    for (int k = 0; k < nc; ++k) {
        const float* r = tp + k * MIX_STRIDE;
        const float d0 = z[0] - r[0], d1 = z[1] - r[1];
        const float b0 = r[2] * d0 + r[3] * d1, b1 = r[3] * d0 + r[4] * d1;
        const float l = -0.5f * (d0 * b0 + d1 * b1) + r[5];
        if (l > m) { m = l; kb = k; }
    }
    float p0, p1;
    {
        const float* r = tp + kb * MIX_STRIDE;
        const float d0 = z[0] - r[0], d1 = z[1] - r[1];
        p0 = -(r[2] * d0 + r[3] * d1); p1 = -(r[3] * d0 + r[4] * d1);
    }
    float S = 0.f, G0 = 0.f, G1 = 0.f, Q0 = 0.f, Q1 = 0.f;
    for (int k = 0; k < nc; ++k) {
        const float* r = tp + k * MIX_STRIDE;
        const float d0 = z[0] - r[0], d1 = z[1] - r[1];
        const float b0 = r[2] * d0 + r[3] * d1, b1 = r[3] * d0 + r[4] * d1;
        const float e = expf((-0.5f * (d0 * b0 + d1 * b1) + r[5]) - m);
        S += e; G0 -= e * b0; G1 -= e * b1;
        if (WANT_HVP) {
            const float c0 = -b0 - p0, c1 = -b1 - p1;
            const float cv = c0 * v[0] + c1 * v[1];
            Q0 += e * (c0 * cv - (r[2] * v[0] + r[3] * v[1]));
            Q1 += e * (c1 * cv - (r[3] * v[0] + r[4] * v[1]));
        }
    }
    const float inv = 1.0f / S;
    g[0] = G0 * inv; g[1] = G1 * inv;
    if (WANT_HVP) {
        const float e0 = g[0] - p0, e1 = g[1] - p1;
        const float ev = e0 * v[0] + e1 * v[1];
        hv[0] = Q0 * inv - e0 * ev;
        hv[1] = Q1 * inv - e1 * ev;
    }
    return logf(S) + m;
}

// ---- 40-GMM fast path (many_gmm, d = 2) -----------------------------------------------------------------------
// Same mathematics as the TGT_MANY_GMM branch of mixture2_eval, specialised for the tensor-core kernels: scalars in
// registers (no array references -> no stack traffic, inlined), component means read as float2 from a dense
// shared-memory array, exp2 with flush-to-zero (arguments are <= 0), displacement-based HVP
//   H v = (1/s^4) [ sum_k r_k delta_k (delta_k . v) - ebar (ebar . v) ] - v / s^2,  delta_k = mu_k - mu_piv,
// which is the pivoted form above written in units of the displacement (one scaling by 1/s^4 at the end).
struct ManyGmmConst {
    float inv_var, hl2, norm_const, invalid_below;   // hl2 = -0.5 inv_var log2(e); norm_const = -2 comp_norm + log_mix
    int nc;
};
__device__ __forceinline__ ManyGmmConst many_gmm_const(const TargetDesc& t) {
    ManyGmmConst c;
    c.inv_var = t.inv_var;
    c.hl2 = -0.5f * t.inv_var * 1.4426950408889634f;
    c.norm_const = -2.0f * t.comp_norm + t.log_mix;
    c.invalid_below = t.invalid_below;
    c.nc = t.ncomp;
    return c;
}
template <bool WANT_HVP>
__device__ __forceinline__ float many_gmm_eval(const ManyGmmConst& c, const float2* __restrict__ smu, float z0, float z1,
                                               float& g0, float& g1, float v0, float v1, float& hv0, float& hv1) {
    float qmin = CUDART_INF_F, p0 = 0.f, p1 = 0.f;   // pivot displacement d_piv = z - mu_piv
#pragma unroll 8
    for (int k = 0; k < c.nc; ++k) {
        const float2 m = smu[k];
        const float d0 = z0 - m.x, d1 = z1 - m.y;
        const float q = fmaf(d0, d0, d1 * d1);
        if (q < qmin) { qmin = q; p0 = d0; p1 = d1; }
    }
    const float off = -c.hl2 * qmin;
    float S = 0.f, G0 = 0.f, G1 = 0.f, Q0 = 0.f, Q1 = 0.f;
#pragma unroll 8
    for (int k = 0; k < c.nc; ++k) {
        const float2 m = smu[k];
        const float d0 = z0 - m.x, d1 = z1 - m.y;
        float e;
        {
            const float arg = fmaf(c.hl2, fmaf(d0, d0, d1 * d1), off);
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(arg));
        }
        S += e; G0 = fmaf(e, d0, G0); G1 = fmaf(e, d1, G1);
        if (WANT_HVP) {
            const float e0 = d0 - p0, e1 = d1 - p1;       // = -(delta_k) ; sign cancels in the quadratic form
            const float t = e * fmaf(e0, v0, e1 * v1);
            Q0 = fmaf(t, e0, Q0); Q1 = fmaf(t, e1, Q1);
        }
    }
    const float lp = logf(S) + fmaf(-0.5f * qmin, c.inv_var, c.norm_const);
    const bool valid = lp > c.invalid_below;
    const float inv = 1.0f / S;
    const float m0 = G0 * inv, m1 = G1 * inv;              // responsibility-weighted mean displacement
    g0 = valid ? -m0 * c.inv_var : 0.f;
    g1 = valid ? -m1 * c.inv_var : 0.f;
    if (WANT_HVP) {
        const float b0 = m0 - p0, b1 = m1 - p1;
        const float bv = fmaf(b0, v0, b1 * v1);
        const float iv2 = c.inv_var * c.inv_var;
        hv0 = valid ? fmaf(iv2, fmaf(Q0, inv, -b0 * bv), -v0 * c.inv_var) : 0.f;
        hv1 = valid ? fmaf(iv2, fmaf(Q1, inv, -b1 * bv), -v1 * c.inv_var) : 0.f;
    }
    return valid ? lp : -CUDART_INF_F;
}

// Score and full 2x2 Hessian of the 40-GMM log-density in one sweep over the components:
//   H = (1/s^4) [ M / S - b b^T ] - I / s^2,   M = sum_k e_k delta_k delta_k^T (3 numbers), b = mbar - d_piv,
// so that every later Hessian-vector product at this point costs 4 FMAs (the adjoint needs H(z_k) v for two different
// v per bridge step: once as z of step k, once as z' of step k-1).
__device__ __forceinline__ float many_gmm_eval_hess(const ManyGmmConst& c, const float2* __restrict__ smu, float z0, float z1,
                                                    float& g0, float& g1, float& h00, float& h01, float& h11) {
    float qmin = CUDART_INF_F, p0 = 0.f, p1 = 0.f;
#pragma unroll 8
    for (int k = 0; k < c.nc; ++k) {
        const float2 m = smu[k];
        const float d0 = z0 - m.x, d1 = z1 - m.y;
        const float q = fmaf(d0, d0, d1 * d1);
        if (q < qmin) { qmin = q; p0 = d0; p1 = d1; }
    }
    const float off = -c.hl2 * qmin;
    float S = 0.f, G0 = 0.f, G1 = 0.f, M00 = 0.f, M01 = 0.f, M11 = 0.f;
#pragma unroll 8
    for (int k = 0; k < c.nc; ++k) {
        const float2 m = smu[k];
        const float d0 = z0 - m.x, d1 = z1 - m.y;
        float e;
        {
            const float arg = fmaf(c.hl2, fmaf(d0, d0, d1 * d1), off);
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(arg));
        }
        S += e; G0 = fmaf(e, d0, G0); G1 = fmaf(e, d1, G1);
        const float e0 = d0 - p0, e1 = d1 - p1;
        const float t0 = e * e0;
        M00 = fmaf(t0, e0, M00); M01 = fmaf(t0, e1, M01); M11 = fmaf(e * e1, e1, M11);
    }
    const float lp = logf(S) + fmaf(-0.5f * qmin, c.inv_var, c.norm_const);
    const bool valid = lp > c.invalid_below;
    const float inv = 1.0f / S;
    const float m0 = G0 * inv, m1 = G1 * inv;
    g0 = valid ? -m0 * c.inv_var : 0.f;
    g1 = valid ? -m1 * c.inv_var : 0.f;
    const float b0 = m0 - p0, b1 = m1 - p1;
    const float iv2 = c.inv_var * c.inv_var;
    h00 = valid ? fmaf(iv2, fmaf(M00, inv, -b0 * b0), -c.inv_var) : 0.f;
    h01 = valid ? iv2 * fmaf(M01, inv, -b0 * b1) : 0.f;
    h11 = valid ? fmaf(iv2, fmaf(M11, inv, -b1 * b1), -c.inv_var) : 0.f;
    return valid ? lp : -CUDART_INF_F;
}

// ---- funnel (any D >= 2) --------------------------------------------------------------------
template <int D, bool WANT_HVP>
__device__ __forceinline__ float funnel_eval(const float (&z)[D], float (&g)[D], const float (&v)[D], float (&hv)[D]) {
    const float vv = z[0];
    const float var = expf(vv);
    const float ldiag = sqrtf(var);
    float ss = 0.f;
#pragma unroll
    for (int j = 1; j < D; ++j) { const float y = z[j] / ldiag; ss += y * y; }
    const float n = (float)(D - 1);
    const float lp_v = -0.5f * (vv / 3.0f) * (vv / 3.0f) - 1.0986122886681098f - 0.9189385332046727f;
    const float lp_o = -0.5f * ss - 0.5f * n * 1.8378770664093453f - n * logf(ldiag);
    const float iv = 1.0f / var;
    g[0] = -vv / 9.0f + 0.5f * ss - 0.5f * n;
    float xv = 0.f;
#pragma unroll
    for (int j = 1; j < D; ++j) { g[j] = -iv * z[j]; if (WANT_HVP) xv += z[j] * v[j]; }
    if (WANT_HVP) {
        hv[0] = (-1.0f / 9.0f - 0.5f * ss) * v[0] + iv * xv;
#pragma unroll
        for (int j = 1; j < D; ++j) hv[j] = iv * (z[j] * v[0] - v[j]);
    }
    return lp_v + lp_o;
}

#ifdef CMCD_INLINE_TARGET
#define CMCD_TARGET_INL __forceinline__
#else
#define CMCD_TARGET_INL __noinline__
#endif
template <int D, bool WANT_HVP>
__device__ CMCD_TARGET_INL float target_eval(const TargetDesc& t, const float* __restrict__ tp,
                                             const float (&z)[D], float (&g)[D],
                                             const float (&v)[D], float (&hv)[D]) {
    if constexpr (D == 2) {
        if (t.kind == TGT_FUNNEL) return funnel_eval<D, WANT_HVP>(z, g, v, hv);
        return mixture2_eval<WANT_HVP>(t, tp, z, g, v, hv);
    } else {
        return funnel_eval<D, WANT_HVP>(z, g, v, hv);
    }
}

}  // namespace cmcd
