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
// Component means are staged in shared memory as PAIRS of components, negated: float4 p = (-x_{2p}, -x_{2p+1}, -y_{2p}, -y_{2p+1})
// (same footprint as a float2 per component), so that one LDS.128 feeds two components and the displacement, squared
// distance, exponent argument and all moment accumulations run as packed fp32x2 instructions (FADD2 / FMUL2 / FFMA2: one
// issue slot for two components).  An odd component count is padded with a component at 1e15 (weight exp(-inf) = 0).
__device__ __forceinline__ void many_gmm_stage_means(const TargetDesc& t, float2* smu, int tid, int nthreads) {
    float* f = reinterpret_cast<float*>(smu);
    const int np = (t.ncomp + 1) >> 1;
    for (int i = tid; i < 2 * np; i += nthreads) {
        const bool real = i < t.ncomp;
        const float mx = real ? t.mix[i * MIX_STRIDE] : 1e15f, my = real ? t.mix[i * MIX_STRIDE + 1] : 1e15f;
        f[(i >> 1) * 4 + (i & 1)] = -mx;
        f[(i >> 1) * 4 + 2 + (i & 1)] = -my;
    }
}
__device__ __forceinline__ unsigned long long mg_pk2(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void mg_upk2(unsigned long long v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ unsigned long long mg_fma2(unsigned long long a, unsigned long long b, unsigned long long c) { unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ unsigned long long mg_mul2(unsigned long long a, unsigned long long b) { unsigned long long d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ unsigned long long mg_add2(unsigned long long a, unsigned long long b) { unsigned long long d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// Sweep 1: the nearest component.  Squared distances are >= 0, so their bit patterns order like the values: the component index
// rides in the 6 low mantissa bits of the key and one integer min per component finds the argmin (ties / near-ties within
// 2^-17 relative go to the lower index -- any near-nearest component is an equally good pivot).  Returns the pivot displacement.
__device__ __forceinline__ void many_gmm_pivot(const ManyGmmConst& c, const float2* __restrict__ smu, float z0, float z1, float& p0, float& p1) {
    const float4* __restrict__ m4 = reinterpret_cast<const float4*>(smu);
    const unsigned long long Z0 = mg_pk2(z0, z0), Z1 = mg_pk2(z1, z1);
    const int np = (c.nc + 1) >> 1;
    uint32_t kmin = 0x7F800000u;
#pragma unroll 4
    for (int p = 0; p < np; ++p) {
        const float4 m = m4[p];
        const unsigned long long D0 = mg_add2(Z0, mg_pk2(m.x, m.y)), D1 = mg_add2(Z1, mg_pk2(m.z, m.w));
        float qa, qb;
        mg_upk2(mg_fma2(D0, D0, mg_mul2(D1, D1)), qa, qb);
        kmin = min(kmin, (__float_as_uint(qa) & 0xFFFFFFC0u) | (uint32_t)(2 * p));
        kmin = min(kmin, (__float_as_uint(qb) & 0xFFFFFFC0u) | (uint32_t)(2 * p + 1));
    }
    const int ks = (int)(kmin & 63u);
    const float* f = reinterpret_cast<const float*>(smu) + (ks >> 1) * 4 + (ks & 1);
    p0 = z0 + f[0];
    p1 = z1 + f[2];
}

template <bool WANT_HVP>
__device__ __forceinline__ float many_gmm_eval(const ManyGmmConst& c, const float2* __restrict__ smu, float z0, float z1,
                                               float& g0, float& g1, float v0, float v1, float& hv0, float& hv1) {
    float p0, p1;                                     // pivot displacement d_piv = z - mu_piv
    many_gmm_pivot(c, smu, z0, z1, p0, p1);
    const float qmin = fmaf(p0, p0, p1 * p1);
    const float off = -c.hl2 * qmin;
    const float4* __restrict__ m4 = reinterpret_cast<const float4*>(smu);
    const unsigned long long Z0 = mg_pk2(z0, z0), Z1 = mg_pk2(z1, z1), HL2 = mg_pk2(c.hl2, c.hl2), OFF = mg_pk2(off, off);
    const unsigned long long NP0 = mg_pk2(-p0, -p0), NP1 = mg_pk2(-p1, -p1), V0 = mg_pk2(v0, v0), V1 = mg_pk2(v1, v1);
    unsigned long long S = mg_pk2(0.f, 0.f), G0 = S, G1 = S, Q0 = S, Q1 = S;
    const int np = (c.nc + 1) >> 1;
#pragma unroll 4
    for (int p = 0; p < np; ++p) {
        const float4 m = m4[p];
        const unsigned long long D0 = mg_add2(Z0, mg_pk2(m.x, m.y)), D1 = mg_add2(Z1, mg_pk2(m.z, m.w));
        float aa, ab, ea, eb;
        mg_upk2(mg_fma2(HL2, mg_fma2(D0, D0, mg_mul2(D1, D1)), OFF), aa, ab);
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ea) : "f"(aa));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(eb) : "f"(ab));
        const unsigned long long E = mg_pk2(ea, eb);
        S = mg_add2(S, E); G0 = mg_fma2(E, D0, G0); G1 = mg_fma2(E, D1, G1);
        if (WANT_HVP) {
            const unsigned long long E0 = mg_add2(D0, NP0), E1 = mg_add2(D1, NP1);   // = -(delta_k); the sign cancels in the quadratic form
            const unsigned long long T = mg_mul2(E, mg_fma2(E0, V0, mg_mul2(E1, V1)));
            Q0 = mg_fma2(T, E0, Q0); Q1 = mg_fma2(T, E1, Q1);
        }
    }
    float sa, sb, ga, gb, ha, hb;
    mg_upk2(S, sa, sb); const float Ss = sa + sb;
    mg_upk2(G0, ga, gb); mg_upk2(G1, ha, hb);
    const float lp = logf(Ss) + fmaf(-0.5f * qmin, c.inv_var, c.norm_const);
    const bool valid = lp > c.invalid_below;
    const float inv = 1.0f / Ss;
    const float m0 = (ga + gb) * inv, m1 = (ha + hb) * inv;   // responsibility-weighted mean displacement
    g0 = valid ? -m0 * c.inv_var : 0.f;
    g1 = valid ? -m1 * c.inv_var : 0.f;
    if (WANT_HVP) {
        float qa, qb, ra, rb;
        mg_upk2(Q0, qa, qb); mg_upk2(Q1, ra, rb);
        const float b0 = m0 - p0, b1 = m1 - p1;
        const float bv = fmaf(b0, v0, b1 * v1);
        const float iv2 = c.inv_var * c.inv_var;
        hv0 = valid ? fmaf(iv2, fmaf(qa + qb, inv, -b0 * bv), -v0 * c.inv_var) : 0.f;
        hv1 = valid ? fmaf(iv2, fmaf(ra + rb, inv, -b1 * bv), -v1 * c.inv_var) : 0.f;
    }
    return valid ? lp : -CUDART_INF_F;
}

// Score and full 2x2 Hessian of the 40-GMM log-density in one sweep over the components:
//   H = (1/s^4) [ M / S - b b^T ] - I / s^2,   M = sum_k e_k delta_k delta_k^T (3 numbers), b = mbar - d_piv,
// so that every later Hessian-vector product at this point costs 4 FMAs (the adjoint needs H(z_k) v for two different
// v per bridge step: once as z of step k, once as z' of step k-1).
__device__ __forceinline__ float many_gmm_eval_hess(const ManyGmmConst& c, const float2* __restrict__ smu, float z0, float z1,
                                                    float& g0, float& g1, float& h00, float& h01, float& h11) {
    float p0, p1;
    many_gmm_pivot(c, smu, z0, z1, p0, p1);
    const float qmin = fmaf(p0, p0, p1 * p1);
    const float off = -c.hl2 * qmin;
    const float4* __restrict__ m4 = reinterpret_cast<const float4*>(smu);
    const unsigned long long Z0 = mg_pk2(z0, z0), Z1 = mg_pk2(z1, z1), HL2 = mg_pk2(c.hl2, c.hl2), OFF = mg_pk2(off, off);
    const unsigned long long NP0 = mg_pk2(-p0, -p0), NP1 = mg_pk2(-p1, -p1);
    unsigned long long S = mg_pk2(0.f, 0.f), G0 = S, G1 = S, M00 = S, M01 = S, M11 = S;
    const int np = (c.nc + 1) >> 1;
#pragma unroll 4
    for (int p = 0; p < np; ++p) {
        const float4 m = m4[p];
        const unsigned long long D0 = mg_add2(Z0, mg_pk2(m.x, m.y)), D1 = mg_add2(Z1, mg_pk2(m.z, m.w));
        float aa, ab, ea, eb;
        mg_upk2(mg_fma2(HL2, mg_fma2(D0, D0, mg_mul2(D1, D1)), OFF), aa, ab);
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ea) : "f"(aa));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(eb) : "f"(ab));
        const unsigned long long E = mg_pk2(ea, eb);
        S = mg_add2(S, E); G0 = mg_fma2(E, D0, G0); G1 = mg_fma2(E, D1, G1);
        const unsigned long long E0 = mg_add2(D0, NP0), E1 = mg_add2(D1, NP1);
        const unsigned long long T0 = mg_mul2(E, E0);
        M00 = mg_fma2(T0, E0, M00); M01 = mg_fma2(T0, E1, M01); M11 = mg_fma2(mg_mul2(E, E1), E1, M11);
    }
    float sa, sb, ga, gb, ha, hb, xa, xb, ya, yb, wa, wb;
    mg_upk2(S, sa, sb); const float Ss = sa + sb;
    mg_upk2(G0, ga, gb); mg_upk2(G1, ha, hb);
    mg_upk2(M00, xa, xb); mg_upk2(M01, ya, yb); mg_upk2(M11, wa, wb);
    const float lp = logf(Ss) + fmaf(-0.5f * qmin, c.inv_var, c.norm_const);
    const bool valid = lp > c.invalid_below;
    const float inv = 1.0f / Ss;
    const float m0 = (ga + gb) * inv, m1 = (ha + hb) * inv;
    g0 = valid ? -m0 * c.inv_var : 0.f;
    g1 = valid ? -m1 * c.inv_var : 0.f;
    const float b0 = m0 - p0, b1 = m1 - p1;
    const float iv2 = c.inv_var * c.inv_var;
    h00 = valid ? fmaf(iv2, fmaf(xa + xb, inv, -b0 * b0), -c.inv_var) : 0.f;
    h01 = valid ? iv2 * fmaf(ya + yb, inv, -b0 * b1) : 0.f;
    h11 = valid ? fmaf(iv2, fmaf(wa + wb, inv, -b1 * b1), -c.inv_var) : 0.f;
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
