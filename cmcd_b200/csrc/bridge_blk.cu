// Block-cooperative FP32 path for FEW particles x WIDE networks (the README's geffner emb_dim-130 configs: N = 2000, K = 256,
// hidden_pad 136, README.md:30,34 -- and every small-N run of the overdamped modes that the tensor-core kernels do not take).
//
// Same contract as bridge_fwd_kernel / bridge_bwd_kernel (bridge_fwd.cu, bridge_bwd.cu; the step algebra and its citations
// are documented there): replaces vmap(compute_log_elbo) (src/mcdboundingmachine.py:126-205) over src/mcd_cais.py:46-89 /
// src/mcd_cais_var.py:56-101 / src/mcd_over_orig.py:18-55 and its jax.grad (src/main.py:174-176).
//
// Why a second mapping: with one thread per particle, N = 2000 gives 63 warps for 592 scheduler slots and every warp walks
// 257 nodes of ~70 k serial instructions (53 ms per train iteration at hidden_pad 136).  Here a CTA owns 32 particles and
// ALL of its warps work on the network of those 32 particles as shared-memory matrix operations:
//   S1 = act(U1^T X + c1[t])                    [HP][32]   element-wise over all threads
//   S2 = act(W2^T S1 + U2^T X + c2[t])          [HP][32]   register-tiled GEMM (8 units x 2 particles per thread), operands in smem
//   O  = W3^T (S2 + skip S1) + U3^T X + c3[t]   [d][32]
// and in the reverse pass dP2 = (W3 Vo) o act'(pre2), gW2 += S1 dP2^T (4x4 register tiles that stay in registers over ALL
// nodes and particle tiles of the CTA: one flush at the end instead of 18 k atomics per node), dA1 = W2 dP2 from a
// transposed copy of W2, dP1 = dA1 o act'(pre1), dX = U1 dP1 + U2 dP2 + U3 Vo.  The per-particle step algebra (keys, Gaussians,
// target score / HVP, kernel means, log-weights, cotangent carry) runs on the first warp between the matrix phases.
// Shared memory at hidden_pad 136: W2 + W2^T 148 KB, three [136][36] activation arrays 59 KB, small tables 5 KB.
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
__device__ __forceinline__ float lg2_ftz(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int ACT>
__device__ __forceinline__ float bk_act(float x) {
    if constexpr (ACT == ACT_SOFTPLUS) {
        const float e = ex2_ftz(-fabsf(x) * 1.4426950408889634f);
        return fmaf(0.6931471805599453f, lg2_ftz(1.0f + e), fmaxf(x, 0.f));
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
#pragma unroll 8
        for (int i = 0; i < HP; ++i) {
            const float4 w0 = *reinterpret_cast<const float4*>(s.W2 + (size_t)i * HP + j0);
            const float4 w1 = *reinterpret_cast<const float4*>(s.W2 + (size_t)i * HP + j0 + 4);
            const float2 h = *reinterpret_cast<const float2*>(S1 + i * BK_RS + p0);
            const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                acc[jj][0] = fmaf(h.x, w[jj], acc[jj][0]);
                acc[jj][1] = fmaf(h.y, w[jj], acc[jj][1]);
            }
        }
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
#pragma unroll 8
            for (int jj = 0; jj < HP; ++jj) {
                const float4 w0 = *reinterpret_cast<const float4*>(sW2T + (size_t)jj * HP + i0);
                const float4 w1 = *reinterpret_cast<const float4*>(sW2T + (size_t)jj * HP + i0 + 4);
                const float2 h = *reinterpret_cast<const float2*>(S3 + jj * BK_RS + p0);
                const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int ii = 0; ii < 8; ++ii) {
                    acc[ii][0] = fmaf(h.x, wv[ii], acc[ii][0]);
                    acc[ii][1] = fmaf(h.y, wv[ii], acc[ii][1]);
                }
            }
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

template <int D, int ACT>
__global__ void __launch_bounds__(BK_T, 1) bridge_fwd_blk_kernel(const BridgeArgs a) {
    extern __shared__ float4 smem4[];
    float* sm = reinterpret_cast<float*>(smem4);
    const int tid = threadIdx.x;
    const NetView& nv = a.net;
    const int HP = nv.HP;
    NetSmem ns = net_stage_smem(nv, D, sm);
    float* sTp = sm + net_smem_floats(D, HP);
    const int ntp = (a.tgt.kind == TGT_GMM || a.tgt.kind == TGT_MANY_GMM) ? a.tgt.ncomp * MIX_STRIDE : 0;
    for (int i = tid; i < ntp; i += blockDim.x) sTp[i] = a.tgt.mix[i];
    float2* sMu = reinterpret_cast<float2*>(sTp + ((ntp + 3) & ~3));   // many_gmm: dense component means (fast path, D = 2)
    const bool fast_gmm = (D == 2) && a.tgt.kind == TGT_MANY_GMM;
    if (fast_gmm)
        for (int i = tid; i < a.tgt.ncomp; i += blockDim.x) sMu[i] = make_float2(a.tgt.mix[i * MIX_STRIDE], a.tgt.mix[i * MIX_STRIDE + 1]);
    const ManyGmmConst gc = many_gmm_const(a.tgt);
    float* S1 = reinterpret_cast<float*>(sMu + MIX_MAX);
    float* S2 = S1 + (size_t)HP * BK_RS;
    float* sX = S2 + (size_t)HP * BK_RS;
    float* sO = sX + D * BK_P;
    float* sPart = sO + D * BK_P;   // [HP/8][D][32] layer-3 partials
    __syncthreads();

    const bool cais = (a.mode == CMCD_MODE_CAIS_SN || a.mode == CMCD_MODE_CAIS_VAR_SN);
    const bool nn_b = (a.mode != CMCD_MODE_ULA);
    const bool nn_f = cais;
    const int K = a.K;
    const float out_scale = net_out_scale(nv);
    const bool pt = tid < BK_P;   // particle thread

    float mu[D], sig[D];
#pragma unroll
    for (int j = 0; j < D; ++j) { mu[j] = a.vd_mean[j]; sig[j] = expf(a.vd_logdiag[j]); }

    const long long ntiles = (a.N + BK_P - 1) / BK_P;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long n_raw = tile * BK_P + tid;
        const bool active = pt && n_raw < a.N;
        const long long n = (pt && n_raw < a.N) ? n_raw : a.N - 1;   // idle lanes shadow the last particle, never store
        Key k, ka;
        float z[D], zn[D], xi[D], sp[D], dummy[D], nnv[D], mf[D];
        float w = 0.f, wm = 0.f, lp = 0.f;
#pragma unroll
        for (int j = 0; j < D; ++j) { nnv[j] = 0.f; z[j] = 0.f; sp[j] = 0.f; mf[j] = 0.f; }
        if (pt) {
            k = prng_key(a.seeds[n]);
            split(k, ka, k);
            normal_vec<D>(ka, xi);
            float lq = 0.f;
#pragma unroll
            for (int j = 0; j < D; ++j) {
                z[j] = sig[j] * xi[j] + mu[j];
                const float v = (z[j] - mu[j]) / sig[j];
                lq += -0.5f * v * v - logf(2.5066282746310002f * sig[j]);
            }
            w = -lq;
            if (a.traj && active) {
#pragma unroll
                for (int j = 0; j < D; ++j) a.traj[((size_t)0 * D + j) * a.N + n] = z[j];
            }
            if constexpr (D == 2) {
                if (fast_gmm) lp = many_gmm_eval<false>(gc, sMu, z[0], z[1], sp[0], sp[1], 0.f, 0.f, dummy[0], dummy[1]);
                else lp = target_eval<D, false>(a.tgt, sTp, z, sp, dummy, dummy);
            } else lp = target_eval<D, false>(a.tgt, sTp, z, sp, dummy, dummy);
            ka = split_first(k);    // mcdboundingmachine.py:162
            k = split_second(ka);   // mcd_cais.py:94
#pragma unroll
            for (int j = 0; j < D; ++j) sX[j * BK_P + tid] = z[j];
        }
        if (nn_f) {   // NN(z_0, 0)
            __syncthreads();
            bk_net_fwd<D, ACT, false>(nv, ns, HP, 0, S1, S2, nullptr, sX, sO, sPart);
            __syncthreads();
            if (pt) {
#pragma unroll
                for (int j = 0; j < D; ++j) nnv[j] = out_scale * fminf(fmaxf(sO[j * BK_P + tid], -nv.out_clip), nv.out_clip);
            }
        }
        if (pt && K > 0) step_keys_and_normal<D>(k, xi);   // Gaussians of step 0 (split + normal + discarded split, mcd_cais.py:66-67,87)
        for (int i = 0; i < K; ++i) {
            const float beta = __ldg(a.betas + i), eps = __ldg(a.eps + i);
            const float scale = sqrtf(2.0f * eps);
            // score at z' and the next step's Gaussians: independent of the network output -> in the shadow of the layer-2 GEMM
            auto side = [&]() {
                if constexpr (D == 2) {
                    if (fast_gmm) lp = many_gmm_eval<false>(gc, sMu, zn[0], zn[1], sp[0], sp[1], 0.f, 0.f, dummy[0], dummy[1]);
                    else lp = target_eval<D, false>(a.tgt, sTp, zn, sp, dummy, dummy);
                } else lp = target_eval<D, false>(a.tgt, sTp, zn, sp, dummy, dummy);
                if (i + 1 < K) step_keys_and_normal<D>(k, xi);
            };
            if (pt) {
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    const float sq = -((z[j] - mu[j]) / sig[j]) / sig[j];
                    const float gu = fminf(fmaxf(sp[j], -a.clip_t), a.clip_t);
                    const float gq = fminf(fmaxf(sq, -a.clip_q), a.clip_q);
                    const float uf = -(beta * gu + (1.0f - beta) * gq);
                    mf[j] = z[j] - eps * uf;
                }
                if (nn_f) {
#pragma unroll
                    for (int j = 0; j < D; ++j) mf[j] = mf[j] - eps * nnv[j];
                }
#pragma unroll
                for (int j = 0; j < D; ++j) { zn[j] = mf[j] + scale * xi[j]; sX[j * BK_P + tid] = zn[j]; }
            }
            if (nn_b) {
                __syncthreads();
                bk_net_fwd<D, ACT, false>(nv, ns, HP, cais ? i + 1 : i, S1, S2, nullptr, sX, sO, sPart, side);
                __syncthreads();
            } else if (pt) {
                side();
            }
            if (pt) {
                float mb[D];
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    const float sq = -((zn[j] - mu[j]) / sig[j]) / sig[j];
                    const float gu = fminf(fmaxf(sp[j], -a.clip_t), a.clip_t);
                    const float gq = fminf(fmaxf(sq, -a.clip_q), a.clip_q);
                    const float ub = -(beta * gu + (1.0f - beta) * gq);
                    mb[j] = zn[j] - eps * ub;
                }
                if (nn_b) {
#pragma unroll
                    for (int j = 0; j < D; ++j) {
                        nnv[j] = out_scale * fminf(fmaxf(sO[j * BK_P + tid], -nv.out_clip), nv.out_clip);
                        mb[j] = mb[j] + eps * nnv[j];
                    }
                }
                const float lognorm = logf(2.5066282746310002f * scale);
                const float fk = bk_gauss_logprob<D>(zn, mf, scale, lognorm);
                const float bk = bk_gauss_logprob<D>(z, mb, scale, lognorm);
                wm += bk - fk;
#pragma unroll
                for (int j = 0; j < D; ++j) z[j] = zn[j];
                if (a.traj && active) {
#pragma unroll
                    for (int j = 0; j < D; ++j) a.traj[((size_t)(i + 1) * D + j) * a.N + n] = z[j];
                }
            }
        }
        if (active) {
            w += wm;
            w += lp;
            a.out_negw[n] = -w;
#pragma unroll
            for (int j = 0; j < D; ++j) a.out_z[n * D + j] = z[j];
        }
        __syncthreads();   // sX / sO are rewritten by the next particle tile
    }
}

// =====================================================================================================================
// Reverse pass.  Node form exactly as bridge_bwd_kernel (bridge_bwd.cu): K + 1 nodes z_K .. z_0, one network recompute and one
// pull-back per node with the combined output cotangent, one target score and one combined HVP per node.
template <int D, int ACT>
__global__ void __launch_bounds__(BK_T, 1) bridge_bwd_blk_kernel(const BridgeArgs a, const float* __restrict__ cot_negw,
                                                                 float* __restrict__ partials, const BwdLayout L) {
    extern __shared__ float4 smem4[];
    float* sm = reinterpret_cast<float*>(smem4);
    const int tid = threadIdx.x;
    const NetView& nv = a.net;
    const int HP = nv.HP;
    NetSmem ns = net_stage_smem(nv, D, sm);
    float* sW2T = sm + net_smem_floats(D, HP);
    for (int idx = tid; idx < HP * HP; idx += blockDim.x) { const int i = idx / HP, j = idx % HP; sW2T[(size_t)j * HP + i] = nv.W2[idx]; }
    float* sTp = sW2T + (size_t)HP * HP;
    const int ntp = (a.tgt.kind == TGT_GMM || a.tgt.kind == TGT_MANY_GMM) ? a.tgt.ncomp * MIX_STRIDE : 0;
    for (int i = tid; i < ntp; i += blockDim.x) sTp[i] = a.tgt.mix[i];
    float2* sMu = reinterpret_cast<float2*>(sTp + ((ntp + 3) & ~3));
    const bool fast_gmm = (D == 2) && a.tgt.kind == TGT_MANY_GMM;
    if (fast_gmm)
        for (int i = tid; i < a.tgt.ncomp; i += blockDim.x) sMu[i] = make_float2(a.tgt.mix[i * MIX_STRIDE], a.tgt.mix[i * MIX_STRIDE + 1]);
    const ManyGmmConst gc = many_gmm_const(a.tgt);
    float* S1 = reinterpret_cast<float*>(sMu + MIX_MAX);
    float* S2 = S1 + (size_t)HP * BK_RS;
    float* S3 = S2 + (size_t)HP * BK_RS;
    float* sX = S3 + (size_t)HP * BK_RS;
    float* sO = sX + D * BK_P;
    float* sVo = sO + D * BK_P;
    float* sDx = sVo + D * BK_P;
    float* sPart = sDx + D * BK_P;   // [HP/8][D][32] layer-3 / input-cotangent partials
    __syncthreads();

    float* part = partials + (size_t)blockIdx.x * L.P;
    const bool cais = (a.mode == CMCD_MODE_CAIS_SN || a.mode == CMCD_MODE_CAIS_VAR_SN);
    const bool pathwise = a.mode != CMCD_MODE_CAIS_VAR_SN;
    const bool nn_b = (a.mode != CMCD_MODE_ULA);
    const bool nn_f = cais;
    const int K = a.K;
    const float out_scale = net_out_scale(nv);
    const bool pt = tid < BK_P;

    float mu[D], sig[D], ivar[D];
#pragma unroll
    for (int j = 0; j < D; ++j) { mu[j] = a.vd_mean[j]; sig[j] = expf(a.vd_logdiag[j]); ivar[j] = 1.0f / (sig[j] * sig[j]); }

    // weight-gradient accumulators: 4x4 tiles over interleaved rows (tile (ti, tj): rows ti + G r, columns tj + G q), fixed per thread
    const int G = HP >> 2;
    float gw[BK_MAXT][4][4];
#pragma unroll
    for (int r0 = 0; r0 < BK_MAXT; ++r0)
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q) gw[r0][r][q] = 0.f;

    const long long ntiles = (a.N + BK_P - 1) / BK_P;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long n_raw = tile * BK_P + tid;
        const bool active = pt && n_raw < a.N;
        const long long n = active ? n_raw : a.N - 1;
        const float c = active ? -cot_negw[n] : 0.f;   // dL/dw_n (zero for shadow lanes: all their cotangents vanish)
        float x[D], zup[D], zprev[D], znext[D], carry[D], rS[D], gmu[D], gls[D], zero[D], hv[D], sx[D];
        float hx[3] = {0.f, 0.f, 0.f};   // many_gmm fast path: Hessian of log p at x
#pragma unroll
        for (int j = 0; j < D; ++j) {
            x[j] = pt ? a.traj[((size_t)K * D + j) * a.N + n] : 0.f;
            znext[j] = (pt && K > 0) ? a.traj[((size_t)(K - 1) * D + j) * a.N + n] : 0.f;   // rows are requested one node ahead
            gmu[j] = 0.f; gls[j] = 0.f; zero[j] = 0.f; carry[j] = 0.f; rS[j] = 0.f; zup[j] = 0.f; hv[j] = 0.f; sx[j] = 0.f; zprev[j] = 0.f;
        }
        float cgb = 0.f, cge = 0.f;
        const int t0 = cais ? 0 : -1;
        // target score (and, fast path, Hessian) at the node's point: evaluated one node ahead, in the shadow of the dA1 GEMM
        float sxN[D], hxN[3] = {0.f, 0.f, 0.f}, hvN[D];
#pragma unroll
        for (int d = 0; d < D; ++d) { sxN[d] = 0.f; hvN[d] = 0.f; }
        auto eval_point = [&](const float (&pnt)[D]) {
            if constexpr (D == 2) {
                if (fast_gmm) many_gmm_eval_hess(gc, sMu, pnt[0], pnt[1], sxN[0], sxN[1], hxN[0], hxN[1], hxN[2]);
                else target_eval<D, false>(a.tgt, sTp, pnt, sxN, zero, hvN);
            } else target_eval<D, false>(a.tgt, sTp, pnt, sxN, zero, hvN);
        };
        if (pt) eval_point(x);   // node K
        for (int j = K; j >= 0; --j) {
            const bool hasB = j > 0, hasF = j < K;
            const int t = t0 + j;
            const bool use_nn = K > 0 && (cais || (nn_b && hasB));
            const float bB = hasB ? __ldg(a.betas + j - 1) : 0.f, eB = hasB ? __ldg(a.eps + j - 1) : 0.f;
            const float bF = hasF ? __ldg(a.betas + j) : 0.f, eF = hasF ? __ldg(a.eps + j) : 0.f;
            const float tsB = hasB ? 2.0f * eB : 1.f, tsF = hasF ? 2.0f * eF : 1.f;
            const float ombB = 1.0f - bB, ombF = 1.0f - bF;
            const float cB = hasB ? c : 0.f, cF = hasF ? c : 0.f;
            const float eFn = nn_f ? eF : 0.f;
            float sq[D], mk_t[D], mk_q[D], uB[D], uF[D], dc[D], nn[D], o[D], dx[D];
#pragma unroll
            for (int d = 0; d < D; ++d) { nn[d] = 0.f; dx[d] = 0.f; o[d] = 0.f; sq[d] = 0.f; mk_t[d] = 0.f; mk_q[d] = 0.f; uB[d] = 0.f; uF[d] = 0.f; dc[d] = 0.f; }
            if (pt) {
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    zprev[d] = hasB ? znext[d] : 0.f;
                    if (j > 1) znext[d] = a.traj[((size_t)(j - 2) * D + d) * a.N + n];
                }
                hx[0] = hxN[0]; hx[1] = hxN[1]; hx[2] = hxN[2];
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    sx[d] = sxN[d];
                    sq[d] = -(x[d] - mu[d]) * ivar[d];
                    mk_t[d] = (fabsf(sx[d]) <= a.clip_t) ? 1.f : 0.f;
                    mk_q[d] = (fabsf(sq[d]) <= a.clip_q) ? 1.f : 0.f;
                    const float gu = fminf(fmaxf(sx[d], -a.clip_t), a.clip_t);
                    const float gq = fminf(fmaxf(sq[d], -a.clip_q), a.clip_q);
                    dc[d] = gu - gq;
                    uB[d] = -(bB * gu + ombB * gq);
                    uF[d] = -(bF * gu + ombF * gq);
                    sX[d * BK_P + tid] = x[d];
                }
            }
            if (use_nn) {
                __syncthreads();
                bk_net_fwd<D, ACT, true>(nv, ns, HP, t, S1, S2, S3, sX, sO, sPart);
                __syncthreads();
            }
            float GB[D], GF[D], rB[D], xs[D], vv[D], wq[D];
            float rr = 0.f, xx = 0.f;
#pragma unroll
            for (int d = 0; d < D; ++d) { GB[d] = 0.f; GF[d] = 0.f; rB[d] = 0.f; xs[d] = 0.f; vv[d] = 0.f; wq[d] = 0.f; }
            if (pt) {
                if (use_nn) {
#pragma unroll
                    for (int d = 0; d < D; ++d) { o[d] = sO[d * BK_P + tid]; nn[d] = out_scale * fminf(fmaxf(o[d], -nv.out_clip), nv.out_clip); }
                }
                float gos = 0.f;
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    const float meanB = (x[d] - eB * uB[d]) + eB * nn[d];
                    const float meanF = (x[d] - eF * uF[d]) - eFn * nn[d];
                    rB[d] = (zprev[d] - meanB) / tsB;
                    GB[d] = cB * rB[d];
                    rr = fmaf(rB[d], rB[d], rr);
                    xs[d] = (zup[d] - meanF) / tsF;
                    xx = fmaf(xs[d], xs[d], xx);
                    GF[d] = pathwise ? carry[d] : -cF * xs[d];
                    if (!hasF) GF[d] = 0.f;
                    vv[d] = eB * GB[d] - eFn * GF[d];
                    wq[d] = eB * ombB * GB[d] + eF * ombF * GF[d];
                    if (use_nn) {   // output-layer cotangent on the raw output (clamp mask, out_scale)
                        const float oc = fminf(fmaxf(o[d], -nv.out_clip), nv.out_clip);
                        gos = fmaf(vv[d], oc, gos);
                        sVo[d * BK_P + tid] = (fabsf(o[d]) <= nv.out_clip) ? vv[d] * out_scale : 0.f;
                    }
                }
                if (use_nn) {
                    gos = warp_sum_f(gos);
                    if (tid == 0 && gos != 0.f) atomicAdd(part + L.os, gos);
                }
            }
            if (use_nn) {
                auto side = [&]() { if (hasB) eval_point(zprev); };   // next node's score / Hessian, in the shadow of the dA1 GEMM
                bk_net_bwd<D, ACT, D>(nv, ns, HP, t, S1, S2, S3, sW2T, sX, sVo, sDx, sPart, gw, part, L, side);
            }
            if (pt) {
                if (use_nn) {
#pragma unroll
                    for (int d = 0; d < D; ++d) dx[d] = sDx[d * BK_P + tid];
                }
                if (pathwise) {
                    float vm[D], dummy[D];
#pragma unroll
                    for (int d = 0; d < D; ++d) vm[d] = mk_t[d] * (bB * eB * GB[d] + bF * eF * GF[d]);
                    if constexpr (D == 2) {
                        if (fast_gmm) {
                            hv[0] = fmaf(hx[0], vm[0], hx[1] * vm[1]);
                            hv[1] = fmaf(hx[1], vm[0], hx[2] * vm[1]);
                        } else target_eval<D, true>(a.tgt, sTp, x, dummy, vm, hv);
                    } else target_eval<D, true>(a.tgt, sTp, x, dummy, vm, hv);
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        const float fpart = hasF ? (GF[d] - cF * rS[d]) : c * sx[d];
                        carry[d] = fpart + GB[d] - wq[d] * ivar[d] * mk_q[d] + hv[d] + dx[d];
                    }
                }
                {
                    float gb = cgb, ge = cge;
                    float ngb = 0.f, nge = cB * rr;
                    if (!pathwise) ge -= cF * xx;
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        gb += eF * GF[d] * dc[d];
                        ge += GF[d] * (-uF[d] - (nn_f ? nn[d] : 0.f) + (pathwise ? xs[d] : 0.f));
                        ngb += eB * GB[d] * dc[d];
                        nge += GB[d] * (-uB[d] + nn[d]);
                        gmu[d] += wq[d] * ivar[d] * mk_q[d];
                        gls[d] += wq[d] * mk_q[d] * (-2.0f * sq[d]);
                    }
                    if (hasF) {
                        gb = warp_sum_f(gb); ge = warp_sum_f(ge);
                        if (tid == 0) { atomicAdd(part + L.beta + j, gb); atomicAdd(part + L.eps + j, ge); }
                    }
                    cgb = ngb; cge = nge;
                }
#pragma unroll
                for (int d = 0; d < D; ++d) { rS[d] = rB[d]; zup[d] = x[d]; x[d] = zprev[d]; }
                if (!use_nn && hasB) eval_point(x);   // no GEMM to hide it in at this node (x is z_{j-1} now)
            }
        }
        if (pt) {
#pragma unroll
            for (int j = 0; j < D; ++j) {
                if (pathwise) { gmu[j] += carry[j]; gls[j] += carry[j] * (zup[j] - mu[j]); }
                gls[j] += c;
                const float m1 = warp_sum_f(gmu[j]), m2 = warp_sum_f(gls[j]);
                if (tid == 0) { atomicAdd(part + L.mu + j, m1); atomicAdd(part + L.ls + j, m2); }
            }
        }
        __syncthreads();
    }
    // flush the weight-gradient tiles (each (i, j) has exactly one owner in the CTA)
#pragma unroll
    for (int r0 = 0; r0 < BK_MAXT; ++r0) {
        const int tl = (BK_T - 1 - tid) + r0 * BK_T;
        if (tl < G * G) {
            const int ti = tl / G, tj = tl % G;
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int q = 0; q < 4; ++q) part[L.W2 + (ti + G * r) * HP + (tj + G * q)] = gw[r0][r][q];
        }
    }
}

// =====================================================================================================================
// Underdamped operators (bridge_ud.cu: the step, its coefficient rows and the adjoint algebra are documented there) in the
// block-cooperative mapping.  DI = network input width: D (network on z) or 2 D (network on (z, rho')).
template <int D, int ACT, int DI>
__global__ void __launch_bounds__(BK_T, 1) bridge_ud_fwd_blk_kernel(const BridgeArgs a) {
    extern __shared__ float4 smem4[];
    float* sm = reinterpret_cast<float*>(smem4);
    const int tid = threadIdx.x;
    const NetView& nv = a.net;
    const int HP = nv.HP;
    NetSmem ns = net_stage_smem(nv, D, sm, DI);
    float* sTp = sm + net_smem_floats(D, HP, DI);
    const int ntp = (a.tgt.kind == TGT_GMM || a.tgt.kind == TGT_MANY_GMM) ? a.tgt.ncomp * MIX_STRIDE : 0;
    for (int i = tid; i < ntp; i += blockDim.x) sTp[i] = a.tgt.mix[i];
    float2* sMu = reinterpret_cast<float2*>(sTp + ((ntp + 3) & ~3));
    const bool fast_gmm = (D == 2) && a.tgt.kind == TGT_MANY_GMM;
    if (fast_gmm)
        for (int i = tid; i < a.tgt.ncomp; i += blockDim.x) sMu[i] = make_float2(a.tgt.mix[i * MIX_STRIDE], a.tgt.mix[i * MIX_STRIDE + 1]);
    const ManyGmmConst gc = many_gmm_const(a.tgt);
    float* S1 = reinterpret_cast<float*>(sMu + MIX_MAX);
    float* S2 = S1 + (size_t)HP * BK_RS;
    float* sX = S2 + (size_t)HP * BK_RS;
    float* sO = sX + DI * BK_P;
    float* sPart = sO + D * BK_P;
    __syncthreads();

    const bool nn_fm = a.mode == CMCD_MODE_UD_CAIS;   // network in the forward-kernel mean too
    const int K = a.K;
    const size_t TS = (size_t)3 * D;
    const float out_scale = net_out_scale(nv);
    const bool pt = tid < BK_P;
    float mu[D], sig[D];
#pragma unroll
    for (int j = 0; j < D; ++j) { mu[j] = a.vd_mean[j]; sig[j] = expf(a.vd_logdiag[j]); }

    const long long ntiles = (a.N + BK_P - 1) / BK_P;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long n_raw = tile * BK_P + tid;
        const bool active = pt && n_raw < a.N;
        const long long n = active ? n_raw : a.N - 1;
        Key k, ka;
        float z[D], zn[D], xi[D], rho[D], sp[D], spn[D], dummy[D], zeros[D], mf[D], mb[D], rp[D], rpp[D], rn[D], nnf[D];
        float w = 0.f, wm = 0.f, lp = 0.f;
#pragma unroll
        for (int j = 0; j < D; ++j) { z[j] = 0.f; rho[j] = 0.f; sp[j] = 0.f; spn[j] = 0.f; zeros[j] = 0.f; nnf[j] = 0.f; zn[j] = 0.f; rp[j] = 0.f; rpp[j] = 0.f; mf[j] = 0.f; xi[j] = 0.f; }
        const float ln1 = logf(2.5066282746310002f);
        auto score_at = [&](const float (&pnt)[D], float (&out)[D]) -> float {
            if constexpr (D == 2) {
                if (fast_gmm) return many_gmm_eval<false>(gc, sMu, pnt[0], pnt[1], out[0], out[1], 0.f, 0.f, dummy[0], dummy[1]);
            }
            return target_eval<D, false>(a.tgt, sTp, pnt, out, dummy, dummy);
        };
        if (pt) {
            k = prng_key(a.seeds[n]);
            split(k, ka, k);
            normal_vec<D>(ka, xi);
            float lq = 0.f;
#pragma unroll
            for (int j = 0; j < D; ++j) {
                z[j] = sig[j] * xi[j] + mu[j];
                const float v = (z[j] - mu[j]) / sig[j];
                lq += -0.5f * v * v - logf(2.5066282746310002f * sig[j]);
            }
            w = -lq;
            lp = score_at(z, sp);
            Key g = split_first(k);
            split(g, ka, g);
            normal_vec<D>(ka, rho);
            wm = wm - bk_gauss_logprob<D>(rho, zeros, 1.0f, ln1);
            k = split_second(g);
            step_keys_and_normal<D>(k, xi);   // Gaussians of step 0
        }
        for (int i = 0; i < K; ++i) {
            const float beta = __ldg(a.betas + i), eps = __ldg(a.eps + i);
            const float af = __ldg(a.eps + K + i), sf = __ldg(a.eps + 2 * K + i), ab = __ldg(a.eps + 3 * K + i);
            const float cn = __ldg(a.eps + 4 * K + i), sb = __ldg(a.eps + 5 * K + i), cf = __ldg(a.eps + 6 * K + i);
            if constexpr (DI > D) {
                if (nn_fm) {   // NN((z, rho), i) of the forward-kernel mean
                    if (pt) {
#pragma unroll
                        for (int j = 0; j < D; ++j) { sX[j * BK_P + tid] = z[j]; sX[(D + j) * BK_P + tid] = rho[j]; }
                    }
                    __syncthreads();
                    bk_net_fwd<D, ACT, false, BkNoSide, DI>(nv, ns, HP, i, S1, S2, nullptr, sX, sO, sPart);
                    __syncthreads();
                    if (pt) {
#pragma unroll
                        for (int j = 0; j < D; ++j) nnf[j] = out_scale * fminf(fmaxf(sO[j * BK_P + tid], -nv.out_clip), nv.out_clip);
                    }
                    __syncthreads();   // sO / sX are rewritten below
                }
            }
            if (pt) {
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    mf[j] = rho[j] * af + cf * nnf[j];
                    rp[j] = mf[j] + sf * xi[j];
                    const float sq = -((z[j] - mu[j]) / sig[j]) / sig[j];
                    const float g0 = -(beta * fminf(fmaxf(sp[j], -a.clip_t), a.clip_t) + (1.0f - beta) * sq);
                    rpp[j] = rp[j] - eps * g0 / 2.0f;
                    zn[j] = z[j] + eps * rpp[j];
                    sX[j * BK_P + tid] = z[j];
                    if constexpr (DI > D) sX[(D + j) * BK_P + tid] = rp[j];
                }
                if (a.traj && active) {
#pragma unroll
                    for (int j = 0; j < D; ++j) {
                        a.traj[((size_t)i * TS + j) * a.N + n] = z[j];
                        a.traj[((size_t)i * TS + D + j) * a.N + n] = rho[j];
                        a.traj[((size_t)i * TS + 2 * D + j) * a.N + n] = rp[j];
                    }
                }
            }
            auto side = [&]() {   // score at z' and the next step's Gaussians: in the shadow of the layer-2 GEMM
                lp = score_at(zn, spn);
                if (i + 1 < K) step_keys_and_normal<D>(k, xi);
            };
            __syncthreads();
            bk_net_fwd<D, ACT, false, decltype(side), DI>(nv, ns, HP, i, S1, S2, nullptr, sX, sO, sPart, side);
            __syncthreads();
            if (pt) {
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    const float nnv = out_scale * fminf(fmaxf(sO[j * BK_P + tid], -nv.out_clip), nv.out_clip);
                    mb[j] = rp[j] * ab;
                    mb[j] = mb[j] + cn * nnv;
                    const float sq = -((zn[j] - mu[j]) / sig[j]) / sig[j];
                    const float g1 = -(beta * fminf(fmaxf(spn[j], -a.clip_t), a.clip_t) + (1.0f - beta) * sq);
                    rn[j] = rpp[j] - eps * g1 / 2.0f;
                }
                const float fk = bk_gauss_logprob<D>(rp, mf, sf, logf(2.5066282746310002f * sf));
                const float bk = bk_gauss_logprob<D>(rho, mb, sb, logf(2.5066282746310002f * sb));
                wm += bk - fk;
#pragma unroll
                for (int j = 0; j < D; ++j) { z[j] = zn[j]; rho[j] = rn[j]; sp[j] = spn[j]; }
            }
            __syncthreads();   // sX / sO are rewritten by the next step
        }
        if (active) {
            wm = wm + bk_gauss_logprob<D>(rho, zeros, 1.0f, ln1);
            w += wm;
            if (a.traj) {
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    a.traj[((size_t)K * TS + j) * a.N + n] = z[j];
                    a.traj[((size_t)K * TS + D + j) * a.N + n] = rho[j];
                    a.traj[((size_t)K * TS + 2 * D + j) * a.N + n] = 0.f;
                }
            }
            w += lp;
            a.out_negw[n] = -w;
#pragma unroll
            for (int j = 0; j < D; ++j) a.out_z[n * D + j] = z[j];
        }
    }
}

template <int D, int ACT, int DI>
__global__ void __launch_bounds__(BK_T, 1) bridge_ud_bwd_blk_kernel(const BridgeArgs a, const float* __restrict__ cot_negw,
                                                                    float* __restrict__ partials, const BwdLayout L) {
    extern __shared__ float4 smem4[];
    float* sm = reinterpret_cast<float*>(smem4);
    const int tid = threadIdx.x;
    const NetView& nv = a.net;
    const int HP = nv.HP;
    NetSmem ns = net_stage_smem(nv, D, sm, DI);
    float* sW2T = sm + net_smem_floats(D, HP, DI);
    for (int idx = tid; idx < HP * HP; idx += blockDim.x) { const int i = idx / HP, j = idx % HP; sW2T[(size_t)j * HP + i] = nv.W2[idx]; }
    float* sTp = sW2T + (size_t)HP * HP;
    const int ntp = (a.tgt.kind == TGT_GMM || a.tgt.kind == TGT_MANY_GMM) ? a.tgt.ncomp * MIX_STRIDE : 0;
    for (int i = tid; i < ntp; i += blockDim.x) sTp[i] = a.tgt.mix[i];
    float2* sMu = reinterpret_cast<float2*>(sTp + ((ntp + 3) & ~3));
    const bool fast_gmm = (D == 2) && a.tgt.kind == TGT_MANY_GMM;
    if (fast_gmm)
        for (int i = tid; i < a.tgt.ncomp; i += blockDim.x) sMu[i] = make_float2(a.tgt.mix[i * MIX_STRIDE], a.tgt.mix[i * MIX_STRIDE + 1]);
    const ManyGmmConst gc = many_gmm_const(a.tgt);
    float* S1 = reinterpret_cast<float*>(sMu + MIX_MAX);
    float* S2 = S1 + (size_t)HP * BK_RS;
    float* S3 = S2 + (size_t)HP * BK_RS;
    float* sX = S3 + (size_t)HP * BK_RS;
    float* sO = sX + DI * BK_P;
    float* sVo = sO + D * BK_P;
    float* sDx = sVo + D * BK_P;
    float* sPart = sDx + DI * BK_P;
    __syncthreads();

    float* part = partials + (size_t)blockIdx.x * L.P;
    const bool nn_fm = a.mode == CMCD_MODE_UD_CAIS;
    const bool clipped = a.clip_t < 3.0e38f;
    const int K = a.K;
    const size_t TS = (size_t)3 * D;
    const float out_scale = net_out_scale(nv);
    const bool pt = tid < BK_P;
    float mu[D], sig[D], ivar[D];
#pragma unroll
    for (int j = 0; j < D; ++j) { mu[j] = a.vd_mean[j]; sig[j] = expf(a.vd_logdiag[j]); ivar[j] = 1.0f / (sig[j] * sig[j]); }
    float gw[BK_MAXT][4][4];
#pragma unroll
    for (int r0 = 0; r0 < BK_MAXT; ++r0)
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q) gw[r0][r][q] = 0.f;

    const long long ntiles = (a.N + BK_P - 1) / BK_P;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long n_raw = tile * BK_P + tid;
        const bool active = pt && n_raw < a.N;
        const long long n = active ? n_raw : a.N - 1;
        const float c = active ? -cot_negw[n] : 0.f;
        float zn[D], zb[D], rb[D], gmu[D], gls[D], sp1[D], h1[3] = {0.f, 0.f, 0.f}, hv[D], zero[D];
#pragma unroll
        for (int j = 0; j < D; ++j) { zn[j] = 0.f; zb[j] = 0.f; rb[j] = 0.f; gmu[j] = 0.f; gls[j] = 0.f; sp1[j] = 0.f; hv[j] = 0.f; zero[j] = 0.f; }
        // score (and, fast path, Hessian) of the target at a point
        auto eval_point = [&](const float (&pnt)[D], float (&sc)[D], float (&hh)[3]) {
            if constexpr (D == 2) {
                if (fast_gmm) { many_gmm_eval_hess(gc, sMu, pnt[0], pnt[1], sc[0], sc[1], hh[0], hh[1], hh[2]); return; }
            }
            float hvd[D];
            target_eval<D, false>(a.tgt, sTp, pnt, sc, zero, hvd);
        };
        // H_p(pnt) v : from the stored Hessian (fast path) or a fresh evaluation
        auto hvp_at = [&](const float (&pnt)[D], const float (&hh)[3], const float (&v)[D], float (&out)[D]) {
            if constexpr (D == 2) {
                if (fast_gmm) { out[0] = fmaf(hh[0], v[0], hh[1] * v[1]); out[1] = fmaf(hh[1], v[0], hh[2] * v[1]); return; }
            }
            float scd[D];
            target_eval<D, true>(a.tgt, sTp, pnt, scd, v, out);
        };
        if (pt) {
            float rhoK[D];
#pragma unroll
            for (int j = 0; j < D; ++j) {
                zn[j] = a.traj[((size_t)K * TS + j) * a.N + n];
                rhoK[j] = a.traj[((size_t)K * TS + D + j) * a.N + n];
            }
            eval_point(zn, sp1, h1);
#pragma unroll
            for (int j = 0; j < D; ++j) { zb[j] = c * sp1[j]; rb[j] = (K >= 1) ? -c * rhoK[j] : 0.f; }
        }
        float sp0N[D], h0N[3] = {0.f, 0.f, 0.f};   // score / Hessian at z_i, evaluated one step ahead (in a GEMM shadow)
#pragma unroll
        for (int j = 0; j < D; ++j) sp0N[j] = 0.f;
        float zN[D];
#pragma unroll
        for (int j = 0; j < D; ++j) zN[j] = (pt && K > 0) ? a.traj[((size_t)(K - 1) * TS + j) * a.N + n] : 0.f;
        if (pt && K > 0) eval_point(zN, sp0N, h0N);

        for (int i = K - 1; i >= 0; --i) {
            const float beta = __ldg(a.betas + i), eps = __ldg(a.eps + i);
            const float af = __ldg(a.eps + K + i), sf = __ldg(a.eps + 2 * K + i), ab = __ldg(a.eps + 3 * K + i);
            const float cn = __ldg(a.eps + 4 * K + i), sb = __ldg(a.eps + 5 * K + i), cf = __ldg(a.eps + 6 * K + i);
            const float omb = 1.0f - beta, s2 = sb * sb, he = 0.5f * eps;
            float z[D], rho[D], rp[D], sp0[D], h0[3], zbc[D], rbp[D], G[D], nn[D], o[D], nnf[D], rext[D];
            float gbeta = 0.f, geps = 0.f, gaf = 0.f, gsf = 0.f, gab = 0.f, gcn = 0.f, gsb = 0.f, gcf = 0.f;
#pragma unroll
            for (int j = 0; j < D; ++j) { z[j] = 0.f; rho[j] = 0.f; rp[j] = 0.f; sp0[j] = 0.f; zbc[j] = 0.f; rbp[j] = 0.f; G[j] = 0.f; nn[j] = 0.f; o[j] = 0.f; nnf[j] = 0.f; rext[j] = 0.f; }
            h0[0] = h0N[0]; h0[1] = h0N[1]; h0[2] = h0N[2];
            if (pt) {
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    z[j] = zN[j];
                    sp0[j] = sp0N[j];
                    rho[j] = a.traj[((size_t)i * TS + D + j) * a.N + n];
                    rp[j] = a.traj[((size_t)i * TS + 2 * D + j) * a.N + n];
                    if (i > 0) zN[j] = a.traj[((size_t)(i - 1) * TS + j) * a.N + n];   // next step's point, requested early
                }
                // ---- second half kick: rho_new = rho'' - (eps/2) gradU(z')
                float g1b[D], zbn[D], rbpp[D], g0b[D], hvin[D];
#pragma unroll
                for (int j = 0; j < D; ++j) { g1b[j] = -he * rb[j]; hvin[j] = (fabsf(sp1[j]) <= a.clip_t) ? g1b[j] : 0.f; }
                hvp_at(zn, h1, hvin, hv);
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    const float sq1 = -(zn[j] - mu[j]) * ivar[j];
                    const float c1 = fminf(fmaxf(sp1[j], -a.clip_t), a.clip_t);
                    const float g1 = -(beta * c1 + omb * sq1);
                    zbn[j] = zb[j] - beta * hv[j] + omb * ivar[j] * g1b[j];
                    geps = fmaf(-0.5f * rb[j], g1, geps);
                    gbeta = fmaf(g1b[j], -(c1 - sq1), gbeta);
                    gmu[j] = fmaf(-omb * ivar[j], g1b[j], gmu[j]);
                    gls[j] = fmaf(2.0f * omb * sq1, g1b[j], gls[j]);
                    rbpp[j] = fmaf(eps, zbn[j], rb[j]);      // drift: z' = z + eps rho''
                    g0b[j] = -he * rbpp[j];
                    hvin[j] = (!clipped || fabsf(sp0[j]) <= a.clip_t) ? g0b[j] : 0.f;
                }
                // ---- first half kick: rho'' = rho' - (eps/2) gradU(z)
                hvp_at(z, h0, hvin, hv);
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    const float sq0 = -(z[j] - mu[j]) * ivar[j];
                    const float c0 = fminf(fmaxf(sp0[j], -a.clip_t), a.clip_t);
                    const float g0 = -(beta * c0 + omb * sq0);
                    const float rpp = rp[j] - eps * g0 / 2.0f;
                    geps = fmaf(zbn[j], rpp, geps);
                    geps = fmaf(-0.5f * rbpp[j], g0, geps);
                    gbeta = fmaf(g0b[j], -(c0 - sq0), gbeta);
                    gmu[j] = fmaf(-omb * ivar[j], g0b[j], gmu[j]);
                    gls[j] = fmaf(2.0f * omb * sq0, g0b[j], gls[j]);
                    zbc[j] = zbn[j] - beta * hv[j] + omb * ivar[j] * g0b[j];
                    rbp[j] = rbpp[j];
                    sX[j * BK_P + tid] = z[j];
                    if constexpr (DI > D) sX[(D + j) * BK_P + tid] = rp[j];
                }
            }
            // ---- backward-kernel mean: recompute NN(x, i), pull back v = c_n c r
            __syncthreads();
            bk_net_fwd<D, ACT, true, BkNoSide, DI>(nv, ns, HP, i, S1, S2, S3, sX, sO, sPart);
            __syncthreads();
            if (pt) {
                float rr = 0.f;
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    o[j] = sO[j * BK_P + tid];
                    nn[j] = out_scale * fminf(fmaxf(o[j], -nv.out_clip), nv.out_clip);
                    const float mbv = rp[j] * ab + cn * nn[j];
                    const float r = (rho[j] - mbv) / s2;
                    G[j] = c * r;
                    rr = fmaf(r, r, rr);
                    gab = fmaf(G[j], rp[j], gab);
                    gcn = fmaf(G[j], nn[j], gcn);
                    rbp[j] = fmaf(ab, G[j], rbp[j]);
                }
                gsb = c * (sb * rr - (float)D / sb);
                gsf = c * (float)D / sf;
                float gos = 0.f;
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    const float v = cn * G[j];
                    const float oc = fminf(fmaxf(o[j], -nv.out_clip), nv.out_clip);
                    gos = fmaf(v, oc, gos);
                    sVo[j * BK_P + tid] = (fabsf(o[j]) <= nv.out_clip) ? v * out_scale : 0.f;
                }
                gos = warp_sum_f(gos);
                if (tid == 0 && gos != 0.f) atomicAdd(part + L.os, gos);
            }
            {
                auto side = [&]() { if (i > 0 && !nn_fm) eval_point(zN, sp0N, h0N); };   // next step's score / Hessian in the dA1 shadow
                bk_net_bwd<D, ACT, DI>(nv, ns, HP, i, S1, S2, S3, sW2T, sX, sVo, sDx, sPart, gw, part, L, side);
            }
            if (pt) {
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    zbc[j] += sDx[j * BK_P + tid];
                    if constexpr (DI > D) rbp[j] += sDx[(D + j) * BK_P + tid];
                }
            }
            // ---- forward-kernel mean network (CMCD_MODE_UD_CAIS): m_f = a_f rho + c_f NN((z, rho), i), cotangent of m_f = rbp
            if constexpr (DI > D) {
                if (nn_fm) {
                    __syncthreads();   // everyone is done with sDx / sX of the first pull-back
                    if (pt) {
#pragma unroll
                        for (int j = 0; j < D; ++j) { sX[j * BK_P + tid] = z[j]; sX[(D + j) * BK_P + tid] = rho[j]; }
                    }
                    __syncthreads();
                    bk_net_fwd<D, ACT, true, BkNoSide, DI>(nv, ns, HP, i, S1, S2, S3, sX, sO, sPart);
                    __syncthreads();
                    if (pt) {
                        float gos = 0.f;
#pragma unroll
                        for (int j = 0; j < D; ++j) {
                            const float of = sO[j * BK_P + tid];
                            const float oc = fminf(fmaxf(of, -nv.out_clip), nv.out_clip);
                            nnf[j] = out_scale * oc;
                            const float v = cf * rbp[j];
                            gcf = fmaf(rbp[j], nnf[j], gcf);
                            gos = fmaf(v, oc, gos);
                            sVo[j * BK_P + tid] = (fabsf(of) <= nv.out_clip) ? v * out_scale : 0.f;
                        }
                        gos = warp_sum_f(gos);
                        if (tid == 0 && gos != 0.f) atomicAdd(part + L.os, gos);
                    }
                    auto side = [&]() { if (i > 0) eval_point(zN, sp0N, h0N); };
                    bk_net_bwd<D, ACT, DI>(nv, ns, HP, i, S1, S2, S3, sW2T, sX, sVo, sDx, sPart, gw, part, L, side);
                    if (pt) {
#pragma unroll
                        for (int j = 0; j < D; ++j) { zbc[j] += sDx[j * BK_P + tid]; rext[j] = sDx[(D + j) * BK_P + tid]; }
                    }
                }
            }
            if (pt) {
                // ---- momentum refresh: rho' = m_f + s_f xi,  m_f = a_f rho (+ c_f NN_f)
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    const float mfv = rho[j] * af + cf * nnf[j];
                    gsf = fmaf(rbp[j], (rp[j] - mfv) / sf, gsf);
                    gaf = fmaf(rho[j], rbp[j], gaf);
                    rb[j] = fmaf(af, rbp[j], -G[j]) + rext[j];
                    zb[j] = zbc[j];
                    zn[j] = z[j];
                    sp1[j] = sp0[j];
                }
                h1[0] = h0[0]; h1[1] = h0[1]; h1[2] = h0[2];
                gbeta = warp_sum_f(gbeta); geps = warp_sum_f(geps);
                gaf = warp_sum_f(gaf); gsf = warp_sum_f(gsf); gab = warp_sum_f(gab); gcn = warp_sum_f(gcn); gsb = warp_sum_f(gsb);
                gcf = warp_sum_f(gcf);
                if (tid == 0) {
                    atomicAdd(part + L.beta + i, gbeta);
                    atomicAdd(part + L.eps + i, geps);
                    atomicAdd(part + L.eps + K + i, gaf);
                    atomicAdd(part + L.eps + 2 * K + i, gsf);
                    atomicAdd(part + L.eps + 3 * K + i, gab);
                    atomicAdd(part + L.eps + 4 * K + i, gcn);
                    atomicAdd(part + L.eps + 5 * K + i, gsb);
                    if (nn_fm) atomicAdd(part + L.eps + 6 * K + i, gcf);
                }
            }
            __syncthreads();   // sX / sVo / sDx are rewritten by the next step
        }
        if (pt) {   // z_0 = mu + sigma xi0, w_0 = -log q(z_0); rho_0 is pure noise   (zn = z_0 here)
#pragma unroll
            for (int j = 0; j < D; ++j) {
                gmu[j] += zb[j];
                gls[j] += zb[j] * (zn[j] - mu[j]) + c;
                const float m1 = warp_sum_f(gmu[j]), m2 = warp_sum_f(gls[j]);
                if (tid == 0) { atomicAdd(part + L.mu + j, m1); atomicAdd(part + L.ls + j, m2); }
            }
        }
        __syncthreads();
    }
    const int G4 = HP >> 2;
#pragma unroll
    for (int r0 = 0; r0 < BK_MAXT; ++r0) {
        const int tl = (BK_T - 1 - tid) + r0 * BK_T;
        if (tl < G4 * G4) {
            const int ti = tl / G4, tj = tl % G4;
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int q = 0; q < 4; ++q) part[L.W2 + (ti + G4 * r) * HP + (tj + G4 * q)] = gw[r0][r][q];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------- launchers
static size_t blk_fwd_smem(int D, int HP) {
    return (net_smem_floats(D, HP) + MIX_MAX * MIX_STRIDE + 2 * MIX_MAX + 2 * (size_t)HP * BK_RS + 2 * (size_t)D * BK_P + (size_t)(HP / 8) * D * BK_P + 8) * sizeof(float);
}
static size_t blk_bwd_smem(int D, int HP) {
    return (net_smem_floats(D, HP) + (size_t)HP * HP + MIX_MAX * MIX_STRIDE + 2 * MIX_MAX + 3 * (size_t)HP * BK_RS + 4 * (size_t)D * BK_P + (size_t)(HP / 8) * D * BK_P + 8) * sizeof(float);
}

// When the block mapping wins over one thread per particle (measured, tools/blk_crossover.py): always for wide networks
// (hidden_pad > 64: the one-thread kernels fit one 64-particle CTA per SM there and are bound by their serial chains), and for
// narrower ones while the one-thread kernels would leave SMs idle.  CMCD_BLK_ALWAYS=1 / CMCD_DISABLE_BLK=1 force either side.
static bool blk_particle_limit_ok(long long N, int HP, int num_sms) {
    if (std::getenv("CMCD_BLK_ALWAYS")) return true;
    if (HP > 64) return true;
    return N <= (long long)BK_P * num_sms * 2;
}

// Few particles (the one-thread-per-particle kernels would leave most SMs idle), a network, widths the register tiles cover.
bool blk_supported(const BridgeArgs& a, int D, int num_sms) {
    if (a.net.arch == CMCD_ARCH_NONE || a.K < 1 || a.mode > CMCD_MODE_CAIS_VAR_SN || a.mode == CMCD_MODE_ULA) return false;
    if (D != 2 && D != 10) return false;
    const int HP = a.net.HP;
    if (HP > BK_HP_MAX || (HP & 7)) return false;
    if (blk_bwd_smem(D, HP) > 227 * 1024) return false;
    return blk_particle_limit_ok(a.N, HP, num_sms);
}

template <int D, int ACT>
static int launch_fwd_blk_t(const BridgeArgs& a, cudaStream_t st, int num_sms) {
    const size_t smem = blk_fwd_smem(D, a.net.HP);
    auto kern = bridge_fwd_blk_kernel<D, ACT>;
    CMCD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long ntiles = (a.N + BK_P - 1) / BK_P;
    const int grid = (int)(ntiles < num_sms ? ntiles : num_sms);
    kern<<<grid < 1 ? 1 : grid, BK_T, smem, st>>>(a);
    CMCD_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_bridge_fwd_blk(const BridgeArgs& a, int D, cudaStream_t st, int num_sms) {
    const bool gelu = a.net.arch == CMCD_ARCH_DDS;
    if (D == 2) return gelu ? launch_fwd_blk_t<2, ACT_GELU>(a, st, num_sms) : launch_fwd_blk_t<2, ACT_SOFTPLUS>(a, st, num_sms);
    if (D == 10) return gelu ? launch_fwd_blk_t<10, ACT_GELU>(a, st, num_sms) : launch_fwd_blk_t<10, ACT_SOFTPLUS>(a, st, num_sms);
    set_error("bridge_fwd_blk: dim=%d has no instantiation", D);
    return 2;
}

size_t bridge_bwd_blk_workspace_bytes(int D, int K, int HP, int arch, int num_sms) {
    return (size_t)num_sms * make_layout(D, K, HP, arch).P * sizeof(float);
}

template <int D, int ACT>
static int launch_bwd_blk_t(const BridgeArgs& a, cudaStream_t st, int num_sms, const float* cot, const BwdOut& out, void* ws, size_t ws_bytes) {
    const int HP = a.net.HP;
    const size_t smem = blk_bwd_smem(D, HP);
    auto kern = bridge_bwd_blk_kernel<D, ACT>;
    CMCD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const BwdLayout L = make_layout(D, a.K, HP, a.net.arch);
    const long long ntiles = (a.N + BK_P - 1) / BK_P;
    int grid = (int)(ntiles < num_sms ? ntiles : num_sms);
    if (grid < 1) grid = 1;
    const size_t need = (size_t)grid * L.P * sizeof(float);
    if (ws_bytes < need || !ws) { set_error("bridge_bwd_blk: workspace too small (%zu < %zu)", ws_bytes, need); return 2; }
    CMCD_CUDA_OK(cudaMemsetAsync(ws, 0, need, st));
    kern<<<grid, BK_T, smem, st>>>(a, cot, (float*)ws, L);
    CMCD_CUDA_OK(cudaGetLastError());
    return launch_bwd_reduce((const float*)ws, grid, L, out, HP, D, a.K, st);
}

int launch_bridge_bwd_blk(const BridgeArgs& a, int D, cudaStream_t st, int num_sms, const float* cot_negw,
                          float* g_vd_mean, float* g_vd_logdiag, float* g_betas, float* g_eps,
                          const cmcd_net_grad* g, void* ws, size_t ws_bytes) {
    BwdOut o{};
    if (g) {
        o.W2 = g->W2; o.U1 = g->U1; o.U2 = g->U2; o.U3 = g->U3; o.W3 = g->W3;
        o.c1 = g->c1; o.c2 = g->c2; o.c3 = g->c3; o.os = g->out_scale;
    }
    o.beta = g_betas; o.eps = g_eps; o.mu = g_vd_mean; o.ls = g_vd_logdiag;
    const bool gelu = a.net.arch == CMCD_ARCH_DDS;
    if (D == 2) return gelu ? launch_bwd_blk_t<2, ACT_GELU>(a, st, num_sms, cot_negw, o, ws, ws_bytes)
                            : launch_bwd_blk_t<2, ACT_SOFTPLUS>(a, st, num_sms, cot_negw, o, ws, ws_bytes);
    if (D == 10) return gelu ? launch_bwd_blk_t<10, ACT_GELU>(a, st, num_sms, cot_negw, o, ws, ws_bytes)
                             : launch_bwd_blk_t<10, ACT_SOFTPLUS>(a, st, num_sms, cot_negw, o, ws, ws_bytes);
    set_error("bridge_bwd_blk: dim=%d has no instantiation", D);
    return 2;
}


// ---- underdamped operators through the block path
static size_t blk_ud_fwd_smem(int D, int DI, int HP) {
    return (net_smem_floats(D, HP, DI) + MIX_MAX * MIX_STRIDE + 2 * MIX_MAX + 2 * (size_t)HP * BK_RS + (size_t)(DI + D) * BK_P +
            (size_t)(HP / 8) * DI * BK_P + 8) * sizeof(float);
}
static size_t blk_ud_bwd_smem(int D, int DI, int HP) {
    return (net_smem_floats(D, HP, DI) + (size_t)HP * HP + MIX_MAX * MIX_STRIDE + 2 * MIX_MAX + 3 * (size_t)HP * BK_RS +
            (size_t)(2 * DI + 2 * D) * BK_P + (size_t)(HP / 8) * DI * BK_P + 8) * sizeof(float);
}

bool blk_ud_supported(const BridgeArgs& a, int D, int num_sms) {
    if (a.net.arch == CMCD_ARCH_NONE || a.K < 1) return false;
    const int din = ud_net_in(a.mode, D);
    if (din == 0 || (D != 2 && D != 10)) return false;
    const int HP = a.net.HP;
    if (HP > BK_HP_MAX || (HP & 7)) return false;
    if (blk_ud_bwd_smem(D, din, HP) > 227 * 1024) return false;
    return blk_particle_limit_ok(a.N, HP, num_sms);
}

template <int D, int ACT, int DI>
static int launch_ud_fwd_blk_t(const BridgeArgs& a, cudaStream_t st, int num_sms) {
    const size_t smem = blk_ud_fwd_smem(D, DI, a.net.HP);
    auto kern = bridge_ud_fwd_blk_kernel<D, ACT, DI>;
    CMCD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long ntiles = (a.N + BK_P - 1) / BK_P;
    const int grid = (int)(ntiles < num_sms ? ntiles : num_sms);
    kern<<<grid < 1 ? 1 : grid, BK_T, smem, st>>>(a);
    CMCD_CUDA_OK(cudaGetLastError());
    return 0;
}
template <int D, int DI>
static int launch_ud_fwd_blk_a(const BridgeArgs& a, cudaStream_t st, int num_sms) {
    return a.net.arch == CMCD_ARCH_DDS ? launch_ud_fwd_blk_t<D, ACT_GELU, DI>(a, st, num_sms)
                                       : launch_ud_fwd_blk_t<D, ACT_SOFTPLUS, DI>(a, st, num_sms);
}
int launch_bridge_ud_fwd_blk(const BridgeArgs& a, int D, cudaStream_t st, int num_sms) {
    const int din = ud_net_in(a.mode, D);
    if (D == 2) return din == 2 ? launch_ud_fwd_blk_a<2, 2>(a, st, num_sms) : launch_ud_fwd_blk_a<2, 4>(a, st, num_sms);
    if (D == 10) return din == 10 ? launch_ud_fwd_blk_a<10, 10>(a, st, num_sms) : launch_ud_fwd_blk_a<10, 20>(a, st, num_sms);
    set_error("bridge_ud_fwd_blk: dim=%d has no instantiation", D);
    return 2;
}

template <int D, int ACT, int DI>
static int launch_ud_bwd_blk_t(const BridgeArgs& a, cudaStream_t st, int num_sms, const float* cot, const BwdOut& out, void* ws, size_t ws_bytes) {
    const int HP = a.net.HP;
    const size_t smem = blk_ud_bwd_smem(D, DI, HP);
    auto kern = bridge_ud_bwd_blk_kernel<D, ACT, DI>;
    CMCD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const BwdLayout L = ud_layout(D, a.K, HP, a.net.arch, DI);
    const long long ntiles = (a.N + BK_P - 1) / BK_P;
    int grid = (int)(ntiles < num_sms ? ntiles : num_sms);
    if (grid < 1) grid = 1;
    const size_t need = (size_t)grid * L.P * sizeof(float);
    if (ws_bytes < need || !ws) { set_error("bridge_ud_bwd_blk: workspace too small (%zu < %zu)", ws_bytes, need); return 2; }
    CMCD_CUDA_OK(cudaMemsetAsync(ws, 0, need, st));
    kern<<<grid, BK_T, smem, st>>>(a, cot, (float*)ws, L);
    CMCD_CUDA_OK(cudaGetLastError());
    return launch_bwd_reduce((const float*)ws, grid, L, out, HP, D, a.K, st);
}
template <int D, int DI>
static int launch_ud_bwd_blk_a(const BridgeArgs& a, cudaStream_t st, int num_sms, const float* cot, const BwdOut& out, void* ws, size_t ws_bytes) {
    return a.net.arch == CMCD_ARCH_DDS ? launch_ud_bwd_blk_t<D, ACT_GELU, DI>(a, st, num_sms, cot, out, ws, ws_bytes)
                                       : launch_ud_bwd_blk_t<D, ACT_SOFTPLUS, DI>(a, st, num_sms, cot, out, ws, ws_bytes);
}
int launch_bridge_ud_bwd_blk(const BridgeArgs& a, int D, cudaStream_t st, int num_sms, const float* cot_negw,
                             float* g_vd_mean, float* g_vd_logdiag, float* g_betas, float* g_eps,
                             const cmcd_net_grad* g, void* ws, size_t ws_bytes) {
    BwdOut o{};
    if (g) {
        o.W2 = g->W2; o.U1 = g->U1; o.U2 = g->U2; o.U3 = g->U3; o.W3 = g->W3;
        o.c1 = g->c1; o.c2 = g->c2; o.c3 = g->c3; o.os = g->out_scale;
    }
    o.beta = g_betas; o.eps = g_eps; o.mu = g_vd_mean; o.ls = g_vd_logdiag;
    const int din = ud_net_in(a.mode, D);
    if (D == 2) return din == 2 ? launch_ud_bwd_blk_a<2, 2>(a, st, num_sms, cot_negw, o, ws, ws_bytes)
                                : launch_ud_bwd_blk_a<2, 4>(a, st, num_sms, cot_negw, o, ws, ws_bytes);
    if (D == 10) return din == 10 ? launch_ud_bwd_blk_a<10, 10>(a, st, num_sms, cot_negw, o, ws, ws_bytes)
                                  : launch_ud_bwd_blk_a<10, 20>(a, st, num_sms, cot_negw, o, ws, ws_bytes);
    set_error("bridge_ud_bwd_blk: dim=%d has no instantiation", D);
    return 2;
}

}  // namespace cmcd
