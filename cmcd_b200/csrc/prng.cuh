// Bit-exact threefry2x32 / split / normal for the CMCD bridge kernels (sm_100a).
//
// Replaces the XLA-generated PRNG chain of the reference:
//   jax.random.PRNGKey/split/normal at src/mcdboundingmachine.py:151-162,
//   src/mcd_cais.py:66,87,94 (same in mcd_cais_var.py / mcd_over_orig.py),
//   src/mcd_utils.py:14-16, src/vardist/diag_gauss.py:49-62.
// Integer stream: exact by construction.  float32 normal: every op is a single IEEE
// round-to-nearest op in a fixed order (__fadd_rn/__fmul_rn/__fdiv_rn/__fsqrt_rn prevent
// FMA contraction), so it matches oracle/prng.py bit for bit.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace cmcd {

struct Key { uint32_t k0, k1; };

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return __funnelshift_l(x, x, r); }

// 20-round Threefry-2x32 (Random123).  Rotation schedule 13,15,26,6 / 17,29,16,24.
// (__noinline__, value in / value out: one shared copy of the 20 rounds keeps the bridge kernels' hot loop
// inside the instruction cache; arguments and results travel in registers.)
static __device__ __noinline__ uint2 threefry2x32_core(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1) {
    const uint32_t ks0 = k0, ks1 = k1, ks2 = k0 ^ k1 ^ 0x1BD11BDAu;
    x0 += ks0; x1 += ks1;
#define CMCD_TF_R(r) { x0 += x1; x1 = rotl32(x1, r); x1 ^= x0; }
    CMCD_TF_R(13) CMCD_TF_R(15) CMCD_TF_R(26) CMCD_TF_R(6)
    x0 += ks1; x1 += ks2 + 1u;
    CMCD_TF_R(17) CMCD_TF_R(29) CMCD_TF_R(16) CMCD_TF_R(24)
    x0 += ks2; x1 += ks0 + 2u;
    CMCD_TF_R(13) CMCD_TF_R(15) CMCD_TF_R(26) CMCD_TF_R(6)
    x0 += ks0; x1 += ks1 + 3u;
    CMCD_TF_R(17) CMCD_TF_R(29) CMCD_TF_R(16) CMCD_TF_R(24)
    x0 += ks1; x1 += ks2 + 4u;
    CMCD_TF_R(13) CMCD_TF_R(15) CMCD_TF_R(26) CMCD_TF_R(6)
    x0 += ks2; x1 += ks0 + 5u;
#undef CMCD_TF_R
    return make_uint2(x0, x1);
}
__device__ __forceinline__ void threefry2x32(Key key, uint32_t& x0, uint32_t& x1) {
    const uint2 r = threefry2x32_core(key.k0, key.k1, x0, x1);
    x0 = r.x; x1 = r.y;
}

// jax.random.split(key) with num=2: threefry over counts iota(4) = blocks (0,2),(1,3).
__device__ __forceinline__ void split(Key key, Key& a, Key& b) {
    uint32_t p0 = 0u, p1 = 2u, q0 = 1u, q1 = 3u;
    threefry2x32(key, p0, p1);
    threefry2x32(key, q0, q1);
    a.k0 = p0; a.k1 = q0;
    b.k0 = p1; b.k1 = q1;
}
__device__ __forceinline__ Key split_first(Key key) { Key a, b; split(key, a, b); return a; }
__device__ __forceinline__ Key split_second(Key key) { Key a, b; split(key, a, b); return b; }

__device__ __forceinline__ Key prng_key(int32_t seed) { Key k; k.k0 = 0u; k.k1 = (uint32_t)seed; return k; }

// ---- deterministic fp32 log / log1p (mirror of oracle/prng.py log_f32 / log1p_f32) ----------
__device__ __forceinline__ float det_logf(float t) {
    // t normal positive.  frexp: m in [0.5,1)
    int e;
    uint32_t b = __float_as_uint(t);
    e = (int)(b >> 23) - 126;
    float m = __uint_as_float((b & 0x007FFFFFu) | 0x3F000000u);
    if (m < 0.70710678118654752440f) { m = __fmul_rn(m, 2.0f); e -= 1; }
    const float f = __fadd_rn(m, -1.0f);
    const float s = __fdiv_rn(f, __fadd_rn(2.0f, f));
    const float z = __fmul_rn(s, s);
    const float w = __fmul_rn(z, z);
    const float t1 = __fmul_rn(w, __fadd_rn(0.40000972152f, __fmul_rn(w, 0.24279078841f)));
    const float t2 = __fmul_rn(z, __fadd_rn(0.66666662693f, __fmul_rn(w, 0.28498786688f)));
    const float r = __fadd_rn(t2, t1);
    const float hfsq = __fmul_rn(__fmul_rn(0.5f, f), f);
    const float dk = (float)e;
    // dk*ln2_hi - ((hfsq - (s*(hfsq+r) + dk*ln2_lo)) - f)
    const float inner = __fadd_rn(__fmul_rn(s, __fadd_rn(hfsq, r)), __fmul_rn(dk, 9.0580006145e-06f));
    return __fadd_rn(__fmul_rn(dk, 6.9313812256e-01f), -__fadd_rn(__fadd_rn(hfsq, -inner), -f));
}

__device__ __forceinline__ float det_log1pf_neg(float u) {  // u in (-1, 0]
    const float t = __fadd_rn(1.0f, u);
    const float c = __fadd_rn(__fadd_rn(t, -1.0f), -u);
    return __fadd_rn(det_logf(t), -__fdiv_rn(c, t));
}

// XLA float32 erf_inv (Giles), op order preserved, no FMA contraction.
__device__ __forceinline__ float erf_inv_f32(float x) {
    float w = -det_log1pf_neg(-__fmul_rn(x, x));
    const bool lt = w < 5.0f;
    w = lt ? __fadd_rn(w, -2.5f) : __fadd_rn(__fsqrt_rn(fmaxf(w, 0.0f)), -3.0f);
    float p = lt ? 2.81022636e-08f : -0.000200214257f;
#define CMCD_EI(a, b) p = __fadd_rn(lt ? (a) : (b), __fmul_rn(p, w));
    CMCD_EI(3.43273939e-07f, 0.000100950558f)
    CMCD_EI(-3.5233877e-06f, 0.00134934322f)
    CMCD_EI(-4.39150654e-06f, -0.00367342844f)
    CMCD_EI(0.00021858087f, 0.00573950773f)
    CMCD_EI(-0.00125372503f, -0.0076224613f)
    CMCD_EI(-0.00417768164f, 0.00943887047f)
    CMCD_EI(0.246640727f, 1.00167406f)
    CMCD_EI(1.50140941f, 2.83297682f)
#undef CMCD_EI
    const float r = __fmul_rn(p, x);
    return (fabsf(x) == 1.0f) ? __fmul_rn(x, __int_as_float(0x7f800000)) : r;
}

// bits -> normal: uniform on [nextafter(-1,0), 1) then sqrt(2)*erf_inv.
__device__ __forceinline__ float bits_to_normal(uint32_t bits) {
    const float lo = -0.99999994f;  // nextafter(-1, 0)
    float f = __fadd_rn(__uint_as_float((bits >> 9) | 0x3F800000u), -1.0f);
    // (maxval - minval) = fl(1 - lo) = 2.0f exactly (ties-to-even)
    f = fmaxf(lo, __fadd_rn(__fmul_rn(f, 2.0f), lo));
    return __fmul_rn(1.41421354f, erf_inv_f32(f));
}

// jax.random.normal(key, (D,)): counts iota(D) padded to even, split in halves.
template <int D>
__device__ __forceinline__ void normal_vec(Key key, float (&out)[D]) {
    constexpr int M = (D + 1) / 2;
#pragma unroll
    for (int j = 0; j < M; ++j) {
        uint32_t x0 = (uint32_t)j;
        uint32_t x1 = (j + M < D) ? (uint32_t)(j + M) : 0u;
        threefry2x32(key, x0, x1);
        out[j] = bits_to_normal(x0);
        if (j + M < D) out[j + M] = bits_to_normal(x1);
    }
}

// N independent Threefry-2x32 blocks advanced in lock step (fully inlined): the 20 rounds of one block are a serial
// dependency chain, interleaving N of them gives the scheduler N-way instruction-level parallelism.
template <int N>
__device__ __forceinline__ void threefry2x32_multi(const Key (&key)[N], uint32_t (&x0)[N], uint32_t (&x1)[N]) {
    uint32_t ks2[N];
#pragma unroll
    for (int n = 0; n < N; ++n) { ks2[n] = key[n].k0 ^ key[n].k1 ^ 0x1BD11BDAu; x0[n] += key[n].k0; x1[n] += key[n].k1; }
#define CMCD_TFM_R(r) { _Pragma("unroll") for (int n = 0; n < N; ++n) { x0[n] += x1[n]; x1[n] = rotl32(x1[n], r); x1[n] ^= x0[n]; } }
#define CMCD_TFM_INJ(a, b, c) { _Pragma("unroll") for (int n = 0; n < N; ++n) { x0[n] += (a); x1[n] += (b) + (c); } }
    CMCD_TFM_R(13) CMCD_TFM_R(15) CMCD_TFM_R(26) CMCD_TFM_R(6)
    CMCD_TFM_INJ(key[n].k1, ks2[n], 1u)
    CMCD_TFM_R(17) CMCD_TFM_R(29) CMCD_TFM_R(16) CMCD_TFM_R(24)
    CMCD_TFM_INJ(ks2[n], key[n].k0, 2u)
    CMCD_TFM_R(13) CMCD_TFM_R(15) CMCD_TFM_R(26) CMCD_TFM_R(6)
    CMCD_TFM_INJ(key[n].k0, key[n].k1, 3u)
    CMCD_TFM_R(17) CMCD_TFM_R(29) CMCD_TFM_R(16) CMCD_TFM_R(24)
    CMCD_TFM_INJ(key[n].k1, ks2[n], 4u)
    CMCD_TFM_R(13) CMCD_TFM_R(15) CMCD_TFM_R(26) CMCD_TFM_R(6)
    CMCD_TFM_INJ(ks2[n], key[n].k0, 5u)
#undef CMCD_TFM_R
#undef CMCD_TFM_INJ
}

// One bridge step's key traffic in two interleaved batches (same values as split / normal_vec / split_second):
//   (ka, kn) = split(k);  xi = normal(ka, (D,));  k <- second half of split(kn)      (mcd_cais.py:66-67,87)
template <int D>
__device__ __forceinline__ void step_keys_and_normal(Key& k, float (&xi)[D]) {
    constexpr int M = (D + 1) / 2;
    Key ka, kn;
    {
        Key kk[2] = {k, k};
        uint32_t x0[2] = {0u, 1u}, x1[2] = {2u, 3u};
        threefry2x32_multi<2>(kk, x0, x1);
        ka.k0 = x0[0]; ka.k1 = x0[1]; kn.k0 = x1[0]; kn.k1 = x1[1];
    }
    Key kk[M + 2];
    uint32_t x0[M + 2], x1[M + 2];
#pragma unroll
    for (int j = 0; j < M; ++j) { kk[j] = ka; x0[j] = (uint32_t)j; x1[j] = (j + M < D) ? (uint32_t)(j + M) : 0u; }
    kk[M] = kn; x0[M] = 0u; x1[M] = 2u;
    kk[M + 1] = kn; x0[M + 1] = 1u; x1[M + 1] = 3u;
    threefry2x32_multi<M + 2>(kk, x0, x1);
#pragma unroll
    for (int j = 0; j < M; ++j) {
        xi[j] = bits_to_normal(x0[j]);
        if (j + M < D) xi[j + M] = bits_to_normal(x1[j]);
    }
    k.k0 = x1[M]; k.k1 = x1[M + 1];
}

// Runtime-d element access (wide path, d=1600): element j of normal(key,(d,)).
__device__ __forceinline__ uint32_t random_bits_at(Key key, int j, int d) {
    const int m = (d + 1) / 2;
    const int blk = (j < m) ? j : j - m;
    uint32_t x0 = (uint32_t)blk;
    uint32_t x1 = (blk + m < d) ? (uint32_t)(blk + m) : 0u;
    threefry2x32(key, x0, x1);
    return (j < m) ? x0 : x1;
}

}  // namespace cmcd
