// Shared pieces of the block-cooperative FP32 mapping (bridge_blk.cu: overdamped operators; bridge_blk_ud.cu: underdamped
// operators): tile constants, MUFU-based activations, the network forward / pull-back on a CTA's 32 particles, dispatch rule.
#pragma once
#include <cstdlib>

#include "net_bwd.cuh"

namespace cmcd {


constexpr int BK_T = 384;      // threads per CTA (12 warps)
constexpr int BK_P = 32;       // particles per CTA (= the first warp)
constexpr int BK_RS = 36;      // row stride of the [HP][32] activation arrays: consecutive rows start 4 banks apart, so the float4
                               // reads of the weight-gradient tiles (one row per lane) are conflict-free (ncu: 34 % of the stall
                               // samples sat on that read at stride 32) and rows stay 16-byte aligned
constexpr int BK_MAXT = 4;     // 4x4 weight-gradient tiles per thread held in registers: (HP/4)^2 <= BK_MAXT * BK_T  =>  HP <= 156
constexpr int BK_HP_MAX = 156;

// ---- activations of the block kernels: MUFU-based forms (ex2 / lg2 / rcp.approx), a quarter of the instructions of the
// expf / log1pf / division forms used by the one-thread kernels -- the activations were 25 % of this path's instructions (ncu).
//   softplus(x) = max(x, 0) + ln2 lg2(1 + e),  e = 2^(-|x| log2 e);  softplus'(x) = 1/(1+e) (x >= 0) | e/(1+e) (x < 0)
// absolute error <= 1.5e-7 on the value (lg2.approx: 2^-22.6 absolute on [1, 2]) and 1.2e-7 on the derivative; the dds GELU uses
// the Abramowitz-Stegun form of the tensor-core kernels (common.cuh, |error| <= 4.7e-7).
template <int ACT>
__device__ __forceinline__ float bk_act(float x) {
    if constexpr (ACT == ACT_SOFTPLUS) {
        return softplus_fast(x);   // common.cuh
    } else {
        return gelu_fast(x);
    }
}
template <int ACT>
__device__ __forceinline__ void bk_act_grad(float x, float& a, float& da) {
    if constexpr (ACT == ACT_SOFTPLUS) {
        const float e = ex2_ftz(-fabsf(x) * 1.4426950408889634f);
        const float t = 1.0f + e;
        a = fmaf(0.6931471805599453f, lg2_ftz(t), fmaxf(x, 0.f));
        const float s = rcp_ftz(t);
        da = x >= 0.f ? s : e * s;
    } else {
        gelu_fast_grad(x, a, da);
    }
}

template <int D>
__device__ __forceinline__ float bk_gauss_logprob(const float (&x)[D], const float (&mean)[D], float scale, float lognorm) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        const float v = (x[j] - mean[j]) / scale;
        s += -0.5f * v * v - lognorm;
    }
    return s;
}

// ---- network forward on the CTA's 32 particles.  In: sX [D][32].  Out: sO [D][32] = RAW output (before clamp / out_scale),
// S1 = a1, S2 = a2, S3 = act'(pre2) if STORE.  Contains two __syncthreads(); the caller synchronises before (sX written) and
// after (sO complete).
// `side()` runs on the particle warp while the other warps are in the layer-2 GEMM (its tiles are handed out from the last thread
// downwards, so warp 0 has none up to hidden_pad 176): per-particle work that does not depend on the network output.
struct BkNoSide { __device__ __forceinline__ void operator()() const {} };
template <int D, int ACT, bool STORE, typename Side = BkNoSide, int DI = D>
__device__ __forceinline__ void bk_net_fwd(const NetView& nv, const NetSmem& s, int HP, int t, float* __restrict__ S1,
                                           float* __restrict__ S2, float* __restrict__ S3, const float* __restrict__ sX,
                                           float* __restrict__ sO, float* __restrict__ sPart, Side side = Side()) {
    constexpr bool has_u2 = (ACT == ACT_SOFTPLUS), has_u3 = (ACT == ACT_SOFTPLUS);
    constexpr float skip = (ACT == ACT_SOFTPLUS) ? 1.f : 0.f;
    const int tid = threadIdx.x;
    const float* __restrict__ c1 = nv.c1 + (size_t)t * HP;
    const float* __restrict__ c2 = nv.c2 + (size_t)t * HP;
    const float* __restrict__ c3 = nv.c3 + (size_t)t * D;
    // layer 1
    for (int idx = tid; idx < HP * BK_P; idx += BK_T) {
        const int j = idx >> 5, p = idx & 31;
        float pre = __ldg(c1 + j);
#pragma unroll
        for (int a = 0; a < DI; ++a) pre = fmaf(sX[a * BK_P + p], s.U1[a * HP + j], pre);
        S1[j * BK_RS + p] = bk_act<ACT>(pre);
    }
    __syncthreads();
    if (tid < BK_P) side();
    // layer 2: thread tile = 8 units x 2 particles
    const int ntile = (HP >> 3) * (BK_P >> 1);
    for (int tile = BK_T - 1 - tid; tile < ntile; tile += BK_T) {   // reversed: the particle warp (warp 0) gets matrix work last
        const int j0 = (tile >> 4) << 3, p0 = (tile & 15) << 1;
        float acc[8][2];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            float b0 = __ldg(c2 + j0 + jj), b1 = b0;
            if (has_u2) {
#pragma unroll
                for (int a = 0; a < DI; ++a) {
                    const float u = s.U2[a * HP + j0 + jj];
                    b0 = fmaf(sX[a * BK_P + p0], u, b0);
                    b1 = fmaf(sX[a * BK_P + p0 + 1], u, b1);
                }
            }
            acc[jj][0] = b0; acc[jj][1] = b1;
        }
        // packed fp32x2: one FFMA2 updates two hidden units of one particle (same per-element rounding as the scalar FMAs)
        f32x2_t acc2[4][2];
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) { acc2[jp][0] = pk2(acc[2 * jp][0], acc[2 * jp + 1][0]); acc2[jp][1] = pk2(acc[2 * jp][1], acc[2 * jp + 1][1]); }
#pragma unroll 8
        for (int i = 0; i < HP; ++i) {
            const float4 w0 = *reinterpret_cast<const float4*>(s.W2 + (size_t)i * HP + j0);
            const float4 w1 = *reinterpret_cast<const float4*>(s.W2 + (size_t)i * HP + j0 + 4);
            const float2 h = *reinterpret_cast<const float2*>(S1 + i * BK_RS + p0);
            const f32x2_t hx = pk2(h.x, h.x), hy = pk2(h.y, h.y);
            const f32x2_t w[4] = {pk2(w0.x, w0.y), pk2(w0.z, w0.w), pk2(w1.x, w1.y), pk2(w1.z, w1.w)};
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) {
                acc2[jp][0] = fma2(hx, w[jp], acc2[jp][0]);
                acc2[jp][1] = fma2(hy, w[jp], acc2[jp][1]);
            }
        }
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) { upk2(acc2[jp][0], acc[2 * jp][0], acc[2 * jp + 1][0]); upk2(acc2[jp][1], acc[2 * jp][1], acc[2 * jp + 1][1]); }
        // activation, and this tile's share of layer 3 (8 of the HP terms of every output of its 2 particles)
        float o3[D][2];
#pragma unroll
        for (int m = 0; m < D; ++m) { o3[m][0] = 0.f; o3[m][1] = 0.f; }
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int o = (j0 + jj) * BK_RS + p0 + q;
                float a2;
                if constexpr (STORE) {
                    float g2;
                    bk_act_grad<ACT>(acc[jj][q], a2, g2);
                    S2[o] = a2; S3[o] = g2;
                } else {
                    a2 = bk_act<ACT>(acc[jj][q]);
                }
                const float hs = a2 + skip * S1[o];
#pragma unroll
                for (int m = 0; m < D; ++m) o3[m][q] = fmaf(hs, s.W3[(j0 + jj) * D + m], o3[m][q]);
            }
        }
        const int jg = tile >> 4;
#pragma unroll
        for (int m = 0; m < D; ++m) {
            sPart[(jg * D + m) * BK_P + p0] = o3[m][0];
            sPart[(jg * D + m) * BK_P + p0 + 1] = o3[m][1];
        }
    }
    __syncthreads();
    // layer 3: fixed-order sum of the HP/8 partials (deterministic), one output element per thread
    for (int idx = tid; idx < D * BK_P; idx += BK_T) {
        const int m = idx >> 5, p = idx & 31;
        float o = __ldg(c3 + m);
        if (has_u3) {
#pragma unroll
            for (int a = 0; a < DI; ++a) o = fmaf(sX[a * BK_P + p], s.U3[a * D + m], o);
        }
        for (int g = 0; g < (HP >> 3); ++g) o += sPart[(g * D + m) * BK_P + p];
        sO[idx] = o;
    }
}

// ---- network pull-back on the CTA's 32 particles (after bk_net_fwd<STORE = true> at the same point).  In: sVo [D][32] = cotangent
// of the raw output, sX [DI][32].  Out: sDx [DI][32] = J_x^T vo; parameter cotangents into `gw` (W2, register tiles) and the CTA's
// partial slice.  Every thread of the CTA calls it; contains five __syncthreads() (the first on entry, the last on exit).
template <int D, int ACT, int DI, typename Side = BkNoSide>
__device__ __forceinline__ void bk_net_bwd(const NetView& nv, const NetSmem& ns, int HP, int t, float* __restrict__ S1,
                                           float* __restrict__ S2, float* __restrict__ S3, const float* __restrict__ sW2T,
                                           const float* __restrict__ sX, const float* __restrict__ sVo, float* __restrict__ sDx,
                                           float* __restrict__ sPart, float (&gw)[BK_MAXT][4][4], float* __restrict__ part,
                                           const BwdLayout& L, Side side = Side()) {
    constexpr bool has_u2 = (ACT == ACT_SOFTPLUS), has_u3 = (ACT == ACT_SOFTPLUS);
    constexpr float skip = (ACT == ACT_SOFTPLUS) ? 1.f : 0.f;
    const int tid = threadIdx.x;
    const int G = HP >> 2;
    __syncthreads();
    // ---- dP2 = (W3 Vo) o act'(pre2) -> S3
    for (int idx = tid; idx < HP * BK_P; idx += BK_T) {
        const int jj = idx >> 5, p = idx & 31;
        float d2 = 0.f;
#pragma unroll
        for (int m = 0; m < D; ++m) d2 = fmaf(ns.W3[jj * D + m], sVo[m * BK_P + p], d2);
        S3[jj * BK_RS + p] = d2 * S3[jj * BK_RS + p];
    }
    __syncthreads();
    // ---- gW2 += S1 dP2^T into the register tiles
#pragma unroll
    for (int r0 = 0; r0 < BK_MAXT; ++r0) {
        const int tl = (BK_T - 1 - tid) + r0 * BK_T;
        if (tl < G * G) {
            const int ti = tl / G, tj = tl % G;
#pragma unroll 4
            for (int p = 0; p < BK_P; p += 4) {
                float4 A[4], B[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    A[r] = *reinterpret_cast<const float4*>(S1 + (ti + G * r) * BK_RS + p);
                    B[r] = *reinterpret_cast<const float4*>(S3 + (tj + G * r) * BK_RS + p);
                }
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float s = gw[r0][r][q];
                        s = fmaf(A[r].x, B[q].x, s); s = fmaf(A[r].y, B[q].y, s);
                        s = fmaf(A[r].z, B[q].z, s); s = fmaf(A[r].w, B[q].w, s);
                        gw[r0][r][q] = s;
                    }
            }
        }
    }
    // ---- skinny cotangents of this node: gc2[t], gU2, gW3 (one hidden unit per thread), gc3[t], gU3
    for (int jj = tid; jj < HP; jj += BK_T) {
        float s2 = 0.f, gu2[DI], gw3[D];
#pragma unroll
        for (int m = 0; m < DI; ++m) gu2[m] = 0.f;
#pragma unroll
        for (int m = 0; m < D; ++m) gw3[m] = 0.f;
        for (int p = 0; p < BK_P; ++p) {
            const float d2 = S3[jj * BK_RS + p], h = S2[jj * BK_RS + p] + skip * S1[jj * BK_RS + p];
            s2 += d2;
            if (has_u2) {
#pragma unroll
                for (int m = 0; m < DI; ++m) gu2[m] = fmaf(d2, sX[m * BK_P + p], gu2[m]);
            }
#pragma unroll
            for (int m = 0; m < D; ++m) gw3[m] = fmaf(h, sVo[m * BK_P + p], gw3[m]);
        }
        atomicAdd(part + L.c2 + (size_t)t * HP + jj, s2);
#pragma unroll
        for (int m = 0; m < D; ++m) atomicAdd(part + L.W3 + jj * D + m, gw3[m]);
        if (has_u2) {
#pragma unroll
            for (int m = 0; m < DI; ++m) atomicAdd(part + L.U2 + m * HP + jj, gu2[m]);
        }
    }
    for (int job = tid; job < D + (has_u3 ? DI * D : 0); job += BK_T) {
        float sacc = 0.f;
        if (job < D) {
            for (int p = 0; p < BK_P; ++p) sacc += sVo[job * BK_P + p];
            atomicAdd(part + L.c3 + (size_t)t * D + job, sacc);
        } else {
            const int aa = (job - D) / D, m = (job - D) % D;
            for (int p = 0; p < BK_P; ++p) sacc = fmaf(sX[aa * BK_P + p], sVo[m * BK_P + p], sacc);
            atomicAdd(part + L.U3 + aa * D + m, sacc);
        }
    }
    __syncthreads();
    // ---- dA1 = W2 dP2 (+ skip W3 Vo); dP1 = dA1 o act'(pre1) -> S2   (thread tile = 8 units x 2 particles, W2^T rows)
    if (tid < BK_P) side();   // per-particle work that needs nothing from this phase: the particle warp has no tile in this GEMM
    {
        const float* __restrict__ c1 = nv.c1 + (size_t)t * HP;
        const int ntile = (HP >> 3) * (BK_P >> 1);
        for (int tl = BK_T - 1 - tid; tl < ntile; tl += BK_T) {
            const int i0 = (tl >> 4) << 3, p0 = (tl & 15) << 1;
            float acc[8][2];
#pragma unroll
            for (int ii = 0; ii < 8; ++ii) { acc[ii][0] = 0.f; acc[ii][1] = 0.f; }
            f32x2_t acc2[4][2];   // packed fp32x2 over pairs of hidden units, as in the layer-2 product
#pragma unroll
            for (int ip = 0; ip < 4; ++ip) { acc2[ip][0] = pk2(0.f, 0.f); acc2[ip][1] = pk2(0.f, 0.f); }
#pragma unroll 8
            for (int jj = 0; jj < HP; ++jj) {
                const float4 w0 = *reinterpret_cast<const float4*>(sW2T + (size_t)jj * HP + i0);
                const float4 w1 = *reinterpret_cast<const float4*>(sW2T + (size_t)jj * HP + i0 + 4);
                const float2 h = *reinterpret_cast<const float2*>(S3 + jj * BK_RS + p0);
                const f32x2_t hx = pk2(h.x, h.x), hy = pk2(h.y, h.y);
                const f32x2_t wv[4] = {pk2(w0.x, w0.y), pk2(w0.z, w0.w), pk2(w1.x, w1.y), pk2(w1.z, w1.w)};
#pragma unroll
                for (int ip = 0; ip < 4; ++ip) {
                    acc2[ip][0] = fma2(hx, wv[ip], acc2[ip][0]);
                    acc2[ip][1] = fma2(hy, wv[ip], acc2[ip][1]);
                }
            }
#pragma unroll
            for (int ip = 0; ip < 4; ++ip) { upk2(acc2[ip][0], acc[2 * ip][0], acc[2 * ip + 1][0]); upk2(acc2[ip][1], acc[2 * ip][1], acc[2 * ip + 1][1]); }
            float dxp[DI][2];
#pragma unroll
            for (int m = 0; m < DI; ++m) { dxp[m][0] = 0.f; dxp[m][1] = 0.f; }
#pragma unroll
            for (int ii = 0; ii < 8; ++ii) {
                const int i = i0 + ii;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int p = p0 + q;
                    float da1 = acc[ii][q];
                    if (skip != 0.f) {
#pragma unroll
                        for (int m = 0; m < D; ++m) da1 = fmaf(ns.W3[i * D + m], sVo[m * BK_P + p], da1);
                    }
                    float pre = __ldg(c1 + i);
#pragma unroll
                    for (int m = 0; m < DI; ++m) pre = fmaf(sX[m * BK_P + p], ns.U1[m * HP + i], pre);
                    float a1, g1;
                    bk_act_grad<ACT>(pre, a1, g1);
                    const float dp1 = da1 * g1;
                    S2[i * BK_RS + p] = dp1;
                    const float dp2 = has_u2 ? S3[i * BK_RS + p] : 0.f;
#pragma unroll
                    for (int m = 0; m < DI; ++m) {   // this tile's share of dX = U1 dP1 + U2 dP2
                        dxp[m][q] = fmaf(ns.U1[m * HP + i], dp1, dxp[m][q]);
                        if (has_u2) dxp[m][q] = fmaf(ns.U2[m * HP + i], dp2, dxp[m][q]);
                    }
                }
            }
            const int ig = tl >> 4;
#pragma unroll
            for (int m = 0; m < DI; ++m) {
                sPart[(ig * DI + m) * BK_P + p0] = dxp[m][0];
                sPart[(ig * DI + m) * BK_P + p0 + 1] = dxp[m][1];
            }
        }
    }
    __syncthreads();
    // ---- gc1[t], gU1 (one hidden unit per thread); dX = U1 dP1 + U2 dP2 + U3 Vo (one element per thread)
    for (int jj = tid; jj < HP; jj += BK_T) {
        float s1 = 0.f, gu1[DI];
#pragma unroll
        for (int m = 0; m < DI; ++m) gu1[m] = 0.f;
        for (int p = 0; p < BK_P; ++p) {
            const float d1 = S2[jj * BK_RS + p];
            s1 += d1;
#pragma unroll
            for (int m = 0; m < DI; ++m) gu1[m] = fmaf(d1, sX[m * BK_P + p], gu1[m]);
        }
        atomicAdd(part + L.c1 + (size_t)t * HP + jj, s1);
#pragma unroll
        for (int m = 0; m < DI; ++m) atomicAdd(part + L.U1 + m * HP + jj, gu1[m]);
    }
    for (int idx = BK_T - 1 - tid; idx < DI * BK_P; idx += BK_T) {   // reversed: the last warps have no hidden unit above
        const int aa = idx >> 5, p = idx & 31;
        float acc = 0.f;
        if (has_u3) {
#pragma unroll
            for (int m = 0; m < D; ++m) acc = fmaf(ns.U3[aa * D + m], sVo[m * BK_P + p], acc);
        }
        for (int g = 0; g < (HP >> 3); ++g) acc += sPart[(g * DI + aa) * BK_P + p];
        sDx[idx] = acc;
    }
    __syncthreads();
}

// When the block mapping wins over one thread per particle (measured, tools/blk_crossover.py): always for wide networks
// (hidden_pad > 64: the one-thread kernels fit one 64-particle CTA per SM there and are bound by their serial chains), and for
// narrower ones while the one-thread kernels would leave SMs idle.  CMCD_BLK_ALWAYS=1 / CMCD_DISABLE_BLK=1 force either side.
static inline bool blk_particle_limit_ok(long long N, int HP, int num_sms) {
    if (std::getenv("CMCD_BLK_ALWAYS")) return true;
    if (HP > 64) return true;
    return N <= (long long)BK_P * num_sms * 2;
}


}  // namespace cmcd
