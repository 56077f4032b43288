// Reverse-mode (adjoint) bridge kernel: recompute from the stored z_k trajectory, analytic target
// scores / Hessian-vector products, block-cooperative weight-gradient tiles.
//
// Replaces the XLA transposed-scan that jax.grad builds for compute_bound / compute_bound_var
// (src/main.py:174-176 over src/mcdboundingmachine.py:126-231 and the step bodies
// src/mcd_cais.py:46-89, src/mcd_cais_var.py:56-101, src/mcd_over_orig.py:18-55).
//
// Per step k (z = z_k, z' = z_{k+1}, eps, beta, s^2 = 2 eps, c = dL/dw_n, a' = dL/dz'):
//   m_f = z  - eps uf(z)  - eps NN(z,k)            uf = -(beta clip(sp) + (1-beta) clip(sq))
//   m_b = z' - eps ub(z') + eps NN(z',tb)          r  = (z - m_b)/s^2,   G_mb = c r
//   KL / pathwise (z' = m_f + s xi):               abar = a' + J_mb(z')^T G_mb,  G_mf = abar,
//                                                  a = -c r + J_mf(z)^T abar
//   log-variance (stop_gradient on z, z'):         G_mf = -c (z'-m_f)/s^2, a == 0
// The forward-kernel log-prob contributes no pathwise gradient (z'-m_f == s xi), see SURVEY 8a.
// Parameter cotangents are reduced per block into a block-private slice of the workspace
// (no inter-block atomics) and summed by bwd_reduce_kernel.
#include "net.cuh"

// A/B on B200 (tools/ab.sh): inlining the two network functions into the adjoint kernel is 21% faster than
// calling them (the 255-register kernel pays for the ABI spills); the forward kernel prefers the call.
#ifdef CMCD_NOINLINE_NET_BWD
#define CMCD_NETB_INL __noinline__
#else
#define CMCD_NETB_INL __forceinline__
#endif

namespace cmcd {

struct BwdLayout {  // offsets (floats) into one block's partial-gradient slice
    int W2, U1, U2, U3, W3, c1, c2, c3, os, beta, eps, mu, ls, P;
};

static BwdLayout make_layout(int D, int K, int HP, int arch) {
    BwdLayout l;
    int o = 0;
    const int T = K + 1;
    const bool net = arch != CMCD_ARCH_NONE;
    l.W2 = o; o += net ? HP * HP : 0;
    l.U1 = o; o += net ? D * HP : 0;
    l.U2 = o; o += net ? D * HP : 0;
    l.U3 = o; o += net ? D * D : 0;
    l.W3 = o; o += net ? HP * D : 0;
    l.c1 = o; o += net ? T * HP : 0;
    l.c2 = o; o += net ? T * HP : 0;
    l.c3 = o; o += net ? T * D : 0;
    l.os = o; o += 1;
    l.beta = o; o += K > 0 ? K : 1;
    l.eps = o; o += K > 0 ? K : 1;
    l.mu = o; o += D;
    l.ls = o; o += D;
    l.P = (o + 3) & ~3;
    return l;
}

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- network forward with stored activations ------------------------------------------------
// S1 <- a1, S2 <- a2, S3 <- act'(pre2).  Returns raw o (before clamp / out_scale).
template <int D, int ACT, int HPT, int JC, int RS>
__device__ CMCD_NETB_INL void net_fwd_store(const NetView& nv, const NetSmem& s, int t, const float* __restrict__ xin,
                                           float* __restrict__ oout, float* __restrict__ S1c, float* __restrict__ S2c,
                                           float* __restrict__ S3c) {
    const int HP = HPT ? HPT : nv.HP;
    float x[D], o[D];
#pragma unroll
    for (int a = 0; a < D; ++a) x[a] = xin[a];
    const float* __restrict__ c1 = nv.c1 + (size_t)t * HP;
    const float* __restrict__ c2 = nv.c2 + (size_t)t * HP;
    const float* __restrict__ c3 = nv.c3 + (size_t)t * D;
    constexpr bool has_u2 = (ACT == ACT_SOFTPLUS), has_u3 = (ACT == ACT_SOFTPLUS);
    constexpr float skip = (ACT == ACT_SOFTPLUS) ? 1.f : 0.f;
#pragma unroll 4
    for (int j = 0; j < HP; ++j) {
        float p = __ldg(c1 + j);
#pragma unroll
        for (int a = 0; a < D; ++a) p = fmaf(x[a], s.U1[a * HP + j], p);
        S1c[j * RS] = act_fwd<ACT>(p);
    }
#pragma unroll
    for (int m = 0; m < D; ++m) {
        float p = __ldg(c3 + m);
        if (has_u3) {
#pragma unroll
            for (int a = 0; a < D; ++a) p = fmaf(x[a], s.U3[a * D + m], p);
        }
        o[m] = p;
    }
#pragma unroll 1
    for (int j0 = 0; j0 < HP; j0 += JC) {
        float acc[JC];
#pragma unroll
        for (int jj = 0; jj < JC; ++jj) {
            float p = __ldg(c2 + j0 + jj);
            if (has_u2) {
#pragma unroll
                for (int a = 0; a < D; ++a) p = fmaf(x[a], s.U2[a * HP + j0 + jj], p);
            }
            acc[jj] = p;
        }
#pragma unroll (inner_unroll(JC))
        for (int i = 0; i < HP; ++i) {
            const float h = S1c[i * RS];
            const float4* __restrict__ w = reinterpret_cast<const float4*>(s.W2 + (size_t)i * HP + j0);
#pragma unroll
            for (int q = 0; q < JC / 4; ++q) {
                const float4 ww = w[q];
                acc[4 * q + 0] = fmaf(h, ww.x, acc[4 * q + 0]);
                acc[4 * q + 1] = fmaf(h, ww.y, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(h, ww.z, acc[4 * q + 2]);
                acc[4 * q + 3] = fmaf(h, ww.w, acc[4 * q + 3]);
            }
        }
        // park the pre-activations, then a rolled activation / layer-3 loop (compact code)
#pragma unroll
        for (int jj = 0; jj < JC; ++jj) S2c[(j0 + jj) * RS] = acc[jj];
#pragma unroll 4
        for (int jj = 0; jj < JC; ++jj) {
            float a2, da2;
            act_fwd_grad<ACT>(S2c[(j0 + jj) * RS], a2, da2);
            S2c[(j0 + jj) * RS] = a2;
            S3c[(j0 + jj) * RS] = da2;
            const float hs = a2 + skip * S1c[(j0 + jj) * RS];
#pragma unroll
            for (int m = 0; m < D; ++m) o[m] = fmaf(hs, s.W3[(j0 + jj) * D + m], o[m]);
        }
    }
#pragma unroll
    for (int m = 0; m < D; ++m) oout[m] = o[m];
}

// ---- network backward (block-cooperative) ----------------------------------------------------
// v: cotangent on the network output (per particle).  Returns dx = J_x^T v and accumulates the
// parameter cotangents into this block's partial slice.  Contains __syncthreads(): every thread of
// the block must call it (inactive particles pass v = 0).
template <int D, int ACT, int HPT, int JC, int BPB>
__device__ CMCD_NETB_INL void net_bwd(const NetView& nv, const NetSmem& s, int t, const float* __restrict__ xin,
                                     const float* __restrict__ oin, const float* __restrict__ vin, float* __restrict__ dxout,
                                     float* __restrict__ S1, float* __restrict__ S2, float* __restrict__ S3,
                                     float* __restrict__ sX, float* __restrict__ sVo,
                                     float* __restrict__ part, const BwdLayout& L) {
    constexpr int RS = BPB + 4;
    float x[D], o[D], v[D], dx[D];
#pragma unroll
    for (int a = 0; a < D; ++a) { x[a] = xin[a]; o[a] = oin[a]; v[a] = vin[a]; }
    const int HP = HPT ? HPT : nv.HP;
    const int tid = threadIdx.x;
    constexpr float skip = (ACT == ACT_SOFTPLUS) ? 1.f : 0.f;
    constexpr bool has_u2 = (ACT == ACT_SOFTPLUS), has_u3 = (ACT == ACT_SOFTPLUS);
    float* S2c = S2 + tid; float* S3c = S3 + tid;

    // (1) private: output layer cotangent, dp2 = W3 vo * act'(pre2) -> S3
    float vo[D];
    float gos = 0.f;
#pragma unroll
    for (int m = 0; m < D; ++m) {
        const float oc = fminf(fmaxf(o[m], -nv.out_clip), nv.out_clip);
        gos = fmaf(v[m], oc, gos);
        vo[m] = (fabsf(o[m]) <= nv.out_clip) ? v[m] * net_out_scale(nv) : 0.f;
        sVo[m * RS + tid] = vo[m];
        sX[m * RS + tid] = x[m];
    }
#pragma unroll 4
    for (int j = 0; j < HP; ++j) {
        float d2 = 0.f;
#pragma unroll
        for (int m = 0; m < D; ++m) d2 = fmaf(s.W3[j * D + m], vo[m], d2);
        S3c[j * RS] = d2 * S3c[j * RS];
    }
    gos = warp_sum_f(gos);
    if ((tid & 31) == 0 && gos != 0.f) atomicAdd(part + L.os, gos);
    __syncthreads();

    // (2) cooperative: gW2 += a1^T dp2 (4x4 register tiles over interleaved rows), then the
    //     skinny products gc2, gU2, gW3, gc3, gU3.
    {
        const int G = HP / 4;  // tile grid edge; rows of tile (ti,.) are ti + G*r
        for (int tile = tid; tile < G * G; tile += BPB) {
            const int ti = tile / G, tj = tile % G;
            float acc[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[r][q] = 0.f;
            for (int p = 0; p < BPB; p += 4) {
                float4 A[4], B[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    A[r] = *reinterpret_cast<const float4*>(S1 + (ti + G * r) * RS + p);
                    B[r] = *reinterpret_cast<const float4*>(S3 + (tj + G * r) * RS + p);
                }
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        acc[r][q] = fmaf(A[r].x, B[q].x, acc[r][q]);
                        acc[r][q] = fmaf(A[r].y, B[q].y, acc[r][q]);
                        acc[r][q] = fmaf(A[r].z, B[q].z, acc[r][q]);
                        acc[r][q] = fmaf(A[r].w, B[q].w, acc[r][q]);
                    }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    atomicAdd(part + L.W2 + (ti + G * r) * HP + (tj + G * q), acc[r][q]);
        }
        // skinny: one (row j, particle-slice) per thread
        int nparts = 1;                       // power-of-two particle slices per row (slices stay float4 aligned)
        while (nparts * 2 * HP <= BPB) nparts *= 2;
        const int plen = BPB / nparts;
        for (int job = tid; job < HP * nparts; job += BPB) {
            const int j = job % HP, pp = job / HP;
            float s2 = 0.f, gu2[D], gw3[D];
#pragma unroll
            for (int a = 0; a < D; ++a) { gu2[a] = 0.f; gw3[a] = 0.f; }
            for (int p = pp * plen; p < (pp + 1) * plen; p += 4) {
                const float4 d2 = *reinterpret_cast<const float4*>(S3 + j * RS + p);
                const float4 a2 = *reinterpret_cast<const float4*>(S2 + j * RS + p);
                const float4 a1 = *reinterpret_cast<const float4*>(S1 + j * RS + p);
                s2 += (d2.x + d2.y) + (d2.z + d2.w);
                const float h0 = a2.x + skip * a1.x, h1 = a2.y + skip * a1.y, h2 = a2.z + skip * a1.z, h3 = a2.w + skip * a1.w;
#pragma unroll
                for (int a = 0; a < D; ++a) {
                    const float4 xx = *reinterpret_cast<const float4*>(sX + a * RS + p);
                    const float4 vv = *reinterpret_cast<const float4*>(sVo + a * RS + p);
                    gu2[a] += d2.x * xx.x + d2.y * xx.y + d2.z * xx.z + d2.w * xx.w;
                    gw3[a] += h0 * vv.x + h1 * vv.y + h2 * vv.z + h3 * vv.w;
                }
            }
            atomicAdd(part + L.c2 + (size_t)t * HP + j, s2);
#pragma unroll
            for (int a = 0; a < D; ++a) {
                if (has_u2) atomicAdd(part + L.U2 + a * HP + j, gu2[a]);
                atomicAdd(part + L.W3 + j * D + a, gw3[a]);
            }
        }
        // gc3[t][m] = sum_p vo[m][p];  gU3[a][m] = sum_p x[a][p] vo[m][p]
        for (int job = tid; job < D + (has_u3 ? D * D : 0); job += BPB) {
            float sacc = 0.f;
            if (job < D) {
                for (int p = 0; p < BPB; ++p) sacc += sVo[job * RS + p];
                atomicAdd(part + L.c3 + (size_t)t * D + job, sacc);
            } else {
                const int a = (job - D) / D, m = (job - D) % D;
                for (int p = 0; p < BPB; ++p) sacc = fmaf(sX[a * RS + p], sVo[m * RS + p], sacc);
                atomicAdd(part + L.U3 + a * D + m, sacc);
            }
        }
    }
    __syncthreads();

    // (3) private: da1 = W2 dp2 + skip * W3 vo ; dp1 = da1 * act'(pre1) -> S2
    {
        const float* __restrict__ c1 = nv.c1 + (size_t)t * HP;
        for (int j0 = 0; j0 < HP; j0 += JC) {
            float dreg[JC];
#pragma unroll
            for (int jj = 0; jj < JC; ++jj) dreg[jj] = S3c[(j0 + jj) * RS];
#pragma unroll (inner_unroll(JC))
            for (int i = 0; i < HP; ++i) {
                const float4* __restrict__ w = reinterpret_cast<const float4*>(s.W2 + (size_t)i * HP + j0);
                float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
#pragma unroll
                for (int q = 0; q < JC / 4; ++q) {
                    const float4 ww = w[q];
                    p0 = fmaf(ww.x, dreg[4 * q + 0], p0);
                    p1 = fmaf(ww.y, dreg[4 * q + 1], p1);
                    p2 = fmaf(ww.z, dreg[4 * q + 2], p2);
                    p3 = fmaf(ww.w, dreg[4 * q + 3], p3);
                }
                const float part_sum = (p0 + p1) + (p2 + p3);
                if (j0 == 0) S2c[i * RS] = part_sum; else S2c[i * RS] += part_sum;
            }
        }
#pragma unroll 2
        for (int i = 0; i < HP; ++i) {
            float da1 = S2c[i * RS];
            if (skip != 0.f) {
#pragma unroll
                for (int m = 0; m < D; ++m) da1 = fmaf(s.W3[i * D + m], vo[m], da1);
            }
            float p = __ldg(c1 + i);
#pragma unroll
            for (int a = 0; a < D; ++a) p = fmaf(x[a], s.U1[a * HP + i], p);
            float a1, g1;
            act_fwd_grad<ACT>(p, a1, g1);
            S2c[i * RS] = da1 * g1;
        }
    }
    __syncthreads();

    // (4) cooperative: gc1[t] += colsum(dp1), gU1 += x^T dp1 ; private: dx
    {
        int nparts = 1;                       // power-of-two particle slices per row (slices stay float4 aligned)
        while (nparts * 2 * HP <= BPB) nparts *= 2;
        const int plen = BPB / nparts;
        for (int job = tid; job < HP * nparts; job += BPB) {
            const int j = job % HP, pp = job / HP;
            float s1 = 0.f, gu1[D];
#pragma unroll
            for (int a = 0; a < D; ++a) gu1[a] = 0.f;
            for (int p = pp * plen; p < (pp + 1) * plen; p += 4) {
                const float4 d1 = *reinterpret_cast<const float4*>(S2 + j * RS + p);
                s1 += (d1.x + d1.y) + (d1.z + d1.w);
#pragma unroll
                for (int a = 0; a < D; ++a) {
                    const float4 xx = *reinterpret_cast<const float4*>(sX + a * RS + p);
                    gu1[a] += d1.x * xx.x + d1.y * xx.y + d1.z * xx.z + d1.w * xx.w;
                }
            }
            atomicAdd(part + L.c1 + (size_t)t * HP + j, s1);
#pragma unroll
            for (int a = 0; a < D; ++a) atomicAdd(part + L.U1 + a * HP + j, gu1[a]);
        }
#pragma unroll
        for (int a = 0; a < D; ++a) {
            float acc = 0.f;
            if (has_u3) {
#pragma unroll
                for (int m = 0; m < D; ++m) acc = fmaf(s.U3[a * D + m], vo[m], acc);
            }
            dx[a] = acc;
        }
#pragma unroll 4
        for (int j = 0; j < HP; ++j) {
            const float d1 = S2c[j * RS], d2 = S3c[j * RS];
#pragma unroll
            for (int a = 0; a < D; ++a) {
                dx[a] = fmaf(s.U1[a * HP + j], d1, dx[a]);
                if (has_u2) dx[a] = fmaf(s.U2[a * HP + j], d2, dx[a]);
            }
        }
    }
#pragma unroll
    for (int a = 0; a < D; ++a) dxout[a] = dx[a];
    __syncthreads();
}

template <int D, int ACT, int HPT, int JC, int BPB>
__global__ void __launch_bounds__(BPB, 1) bridge_bwd_kernel(const BridgeArgs a, const float* __restrict__ cot_negw,
                                                            float* __restrict__ partials, const BwdLayout L) {
    constexpr int RS = BPB + 4;
    extern __shared__ float4 smem4[];
    float* sm = reinterpret_cast<float*>(smem4);
    const int tid = threadIdx.x;
    const NetView& nv = a.net;
    const int HP = HPT ? HPT : nv.HP;
    const bool has_net = nv.arch != CMCD_ARCH_NONE;
    const float out_scale = has_net ? net_out_scale(nv) : 1.0f;
    NetSmem ns = net_stage_smem(nv, D, sm);
    float* sTp = sm + (has_net ? net_smem_floats(D, HP) : 0);
    const int ntp = (a.tgt.kind == TGT_GMM || a.tgt.kind == TGT_MANY_GMM) ? a.tgt.ncomp * MIX_STRIDE : 0;
    for (int i = tid; i < ntp; i += blockDim.x) sTp[i] = a.tgt.mix[i];
    float* S1 = sTp + ((ntp + 3) & ~3);
    float* S2 = S1 + (has_net ? (size_t)HP * RS : 0);
    float* S3 = S2 + (has_net ? (size_t)HP * RS : 0);
    float* sX = S3 + (has_net ? (size_t)HP * RS : 0);
    float* sVo = sX + D * RS;
    __syncthreads();

    float* part = partials + (size_t)blockIdx.x * L.P;
    const bool cais = (a.mode == CMCD_MODE_CAIS_SN || a.mode == CMCD_MODE_CAIS_VAR_SN);
    const bool pathwise = a.mode != CMCD_MODE_CAIS_VAR_SN;
    const bool nn_b = (a.mode != CMCD_MODE_ULA) && has_net;
    const bool nn_f = cais && has_net;
    const int K = a.K;

    float mu[D], sig[D], ivar[D];
#pragma unroll
    for (int j = 0; j < D; ++j) { mu[j] = a.vd_mean[j]; sig[j] = expf(a.vd_logdiag[j]); ivar[j] = 1.0f / (sig[j] * sig[j]); }

    const long long ntiles = (a.N + BPB - 1) / BPB;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long n = tile * BPB + tid;
        const bool active = n < a.N;
        const float c = active ? -cot_negw[n] : 0.f;   // dL/dw_n
        // Node form: K + 1 nodes z_K .. z_0, one target score, one network recompute and ONE network pull-back per node.
        // In the CAIS modes NN(z_j, j) serves the backward-kernel mean of step j-1 and the forward-kernel mean of step j
        // (mcd_cais.py:78 / :60: same point, same time index) and the VJP is linear in the output cotangent, so both uses
        // share  v = eps_{j-1} G_mb - eps_j G_mf ; the two Hessian-vector products at z_j merge the same way.
        //   carry_{j-1} = [-c r_j + G_mf (1 - eps_F (1-beta_F) mk_q / sigma^2)] + [G_mb (1 - eps_B (1-beta_B) mk_q / sigma^2)]
        //                 + H_p(z_j) mk_t (beta_F eps_F G_mf + beta_B eps_B G_mb) + J_x^T v     (node K: first bracket = c grad log p)
        float x[D], zup[D], zprev[D], carry[D], rS[D], gmu[D], gls[D], zero[D], hv[D], sx[D];
#pragma unroll
        for (int j = 0; j < D; ++j) {
            x[j] = active ? a.traj[((size_t)K * D + j) * a.N + n] : 0.f;
            gmu[j] = 0.f; gls[j] = 0.f; zero[j] = 0.f; carry[j] = 0.f; rS[j] = 0.f; zup[j] = 0.f; hv[j] = 0.f;
        }
        float cgb = 0.f, cge = 0.f;     // beta_i / eps_i cotangent of step j: backward-kernel part, carried from node j+1 to node j
        const int t0 = cais ? 0 : -1;   // table row of node j: t0 + j  (MCD_ULA_sn: NN(z_j, j-1), mcd_over_orig.py:45)
        for (int j = K; j >= 0; --j) {
            const bool hasB = j > 0, hasF = j < K;
            const int t = t0 + j;
            const bool use_nn = has_net && K > 0 && (cais || (nn_b && hasB));
            // step constants of both uses; an absent use gets eps = 0, c = 0 so that all of its terms vanish
            const float bB = hasB ? __ldg(a.betas + j - 1) : 0.f, eB = hasB ? __ldg(a.eps + j - 1) : 0.f;
            const float bF = hasF ? __ldg(a.betas + j) : 0.f, eF = hasF ? __ldg(a.eps + j) : 0.f;
            const float tsB = hasB ? 2.0f * eB : 1.f, tsF = hasF ? 2.0f * eF : 1.f;
            const float ombB = 1.0f - bB, ombF = 1.0f - bF;
            const float cB = hasB ? c : 0.f, cF = hasF ? c : 0.f;
            const float eFn = nn_f ? eF : 0.f;
#pragma unroll
            for (int d = 0; d < D; ++d) zprev[d] = (active && hasB) ? a.traj[((size_t)(j - 1) * D + d) * a.N + n] : 0.f;
            target_eval<D, false>(a.tgt, sTp, x, sx, zero, hv);
            float sq[D], mk_t[D], mk_q[D], uB[D], uF[D], dc[D], nn[D], o[D], dx[D];
#pragma unroll
            for (int d = 0; d < D; ++d) {
                sq[d] = -(x[d] - mu[d]) * ivar[d];
                mk_t[d] = (fabsf(sx[d]) <= a.clip_t) ? 1.f : 0.f;
                mk_q[d] = (fabsf(sq[d]) <= a.clip_q) ? 1.f : 0.f;
                const float gu = fminf(fmaxf(sx[d], -a.clip_t), a.clip_t);
                const float gq = fminf(fmaxf(sq[d], -a.clip_q), a.clip_q);
                dc[d] = gu - gq;                       // d(-u)/dbeta
                uB[d] = -(bB * gu + ombB * gq);
                uF[d] = -(bF * gu + ombF * gq);
                nn[d] = 0.f; dx[d] = 0.f; o[d] = 0.f;
            }
            if (use_nn) {
                net_fwd_store<D, ACT, HPT, JC, RS>(nv, ns, t, x, o, S1 + tid, S2 + tid, S3 + tid);
#pragma unroll
                for (int d = 0; d < D; ++d) nn[d] = out_scale * fminf(fmaxf(o[d], -nv.out_clip), nv.out_clip);
            }
            float GB[D], GF[D], rB[D], xs[D], vv[D], wq[D];
            float rr = 0.f, xx = 0.f;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const float meanB = (x[d] - eB * uB[d]) + eB * nn[d];
                const float meanF = (x[d] - eF * uF[d]) - eFn * nn[d];
                rB[d] = (zprev[d] - meanB) / tsB;
                GB[d] = cB * rB[d];
                rr = fmaf(rB[d], rB[d], rr);
                xs[d] = (zup[d] - meanF) / tsF;      // = xi / s
                xx = fmaf(xs[d], xs[d], xx);
                GF[d] = pathwise ? carry[d] : -cF * xs[d];
                if (!hasF) GF[d] = 0.f;
                vv[d] = eB * GB[d] - eFn * GF[d];                      // cotangent on the network output
                wq[d] = eB * ombB * GB[d] + eF * ombF * GF[d];         // weight of the q-score terms
            }
            if (use_nn) net_bwd<D, ACT, HPT, JC, BPB>(nv, ns, t, x, o, vv, dx, S1, S2, S3, sX, sVo, part, L);
            if (pathwise) {
                float vm[D], dummy[D];
#pragma unroll
                for (int d = 0; d < D; ++d) vm[d] = mk_t[d] * (bB * eB * GB[d] + bF * eF * GF[d]);
                target_eval<D, true>(a.tgt, sTp, x, dummy, vm, hv);
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    const float fpart = hasF ? (GF[d] - cF * rS[d]) : c * sx[d];   // node K: terminal w += log p(z_K) (mcdboundingmachine.py:178)
                    carry[d] = fpart + GB[d] - wq[d] * ivar[d] * mk_q[d] + hv[d] + dx[d];
                }
            }
            // ---------------- per-step scalar cotangents ----------------
            {
                float gb = cgb, ge = cge;         // step j: backward-kernel part from node j+1, forward-kernel part here
                float ngb = 0.f, nge = cB * rr;   // step j-1: backward-kernel part, completed at node j-1
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
                    if ((tid & 31) == 0) { atomicAdd(part + L.beta + j, gb); atomicAdd(part + L.eps + j, ge); }
                }
                cgb = ngb; cge = nge;
            }
#pragma unroll
            for (int d = 0; d < D; ++d) { rS[d] = rB[d]; zup[d] = x[d]; x[d] = zprev[d]; }
        }
        // initial: z0 = mu + sigma xi0, w0 = -log q(z0) = 0.5|xi0|^2 + sum log(sqrt(2pi) sigma)   (zup = z_0 after the last shift)
#pragma unroll
        for (int j = 0; j < D; ++j) {
            if (pathwise) { gmu[j] += carry[j]; gls[j] += carry[j] * (zup[j] - mu[j]); }
            gls[j] += c;
            const float m1 = warp_sum_f(gmu[j]), m2 = warp_sum_f(gls[j]);
            if ((tid & 31) == 0) { atomicAdd(part + L.mu + j, m1); atomicAdd(part + L.ls + j, m2); }
        }
    }
}

// out[k] = sum_b partials[b][k], scattered into the caller's cotangent buffers
struct BwdOut {
    float *W2, *U1, *U2, *U3, *W3, *c1, *c2, *c3, *os, *beta, *eps, *mu, *ls;
};

__global__ void bwd_reduce_kernel(const float* __restrict__ partials, int nblocks, BwdLayout L, BwdOut o, int HP, int D, int K) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= L.P) return;
    float s = 0.f;
    for (int b = 0; b < nblocks; ++b) s += partials[(size_t)b * L.P + k];
    const int T = K + 1;
    auto put = [&](float* dst, int off, int len) { if (dst && k >= off && k < off + len) dst[k - off] = s; };
    put(o.W2, L.W2, L.U1 - L.W2); put(o.U1, L.U1, L.U2 - L.U1); put(o.U2, L.U2, L.U3 - L.U2);
    put(o.U3, L.U3, L.W3 - L.U3); put(o.W3, L.W3, L.c1 - L.W3); put(o.c1, L.c1, L.c2 - L.c1);
    put(o.c2, L.c2, L.c3 - L.c2); put(o.c3, L.c3, L.os - L.c3); put(o.os, L.os, 1);
    put(o.beta, L.beta, K); put(o.eps, L.eps, K); put(o.mu, L.mu, D); put(o.ls, L.ls, D);
    (void)T; (void)HP;
}

template <int D, int ACT, int HPT, int JC, int BPB>
static size_t bwd_smem_bytes(int HP, bool has_net) {
    constexpr int RS = BPB + 4;
    size_t fl = (has_net ? net_smem_floats(D, HP) + 3 * (size_t)HP * RS : 0) + MIX_MAX * MIX_STRIDE + 2 * D * RS + 8;
    return fl * sizeof(float);
}

static int bwd_grid(long long N, int BPB, int num_sms) {
    const long long ntiles = (N + BPB - 1) / BPB;
    long long g = num_sms;
    if (g > ntiles) g = ntiles;
    return (int)(g < 1 ? 1 : g);
}

template <int D, int ACT, int HPT, int JC, int BPB>
static int launch_bwd_t(const BridgeArgs& a, cudaStream_t st, int num_sms, const float* cot, const BwdOut& out,
                        void* ws, size_t ws_bytes) {
    const int HP = a.net.HP;
    const bool has_net = a.net.arch != CMCD_ARCH_NONE;
    const size_t smem = bwd_smem_bytes<D, ACT, HPT, JC, BPB>(HP, has_net);
    if (smem > 227 * 1024) { set_error("bridge_bwd: hidden_pad=%d needs %zu B shared memory (> 227 KB)", HP, smem); return 2; }
    auto kern = bridge_bwd_kernel<D, ACT, HPT, JC, BPB>;
    CMCD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const BwdLayout L = make_layout(D, a.K, HP, a.net.arch);
    const int grid = bwd_grid(a.N, BPB, num_sms);
    const size_t need = (size_t)grid * L.P * sizeof(float);
    if (ws_bytes < need || !ws) { set_error("bridge_bwd: workspace too small (%zu < %zu)", ws_bytes, need); return 2; }
    CMCD_CUDA_OK(cudaMemsetAsync(ws, 0, need, st));
    kern<<<grid, BPB, smem, st>>>(a, cot, (float*)ws, L);
    CMCD_CUDA_OK(cudaGetLastError());
    bwd_reduce_kernel<<<(L.P + 255) / 256, 256, 0, st>>>((const float*)ws, grid, L, out, HP, D, a.K);
    CMCD_CUDA_OK(cudaGetLastError());
    return 0;
}

template <int D>
static int launch_bwd_d(const BridgeArgs& a, cudaStream_t st, int num_sms, const float* cot, const BwdOut& out,
                        void* ws, size_t ws_bytes) {
    // particles per block: bounded by the three [HP][BPB] activation arrays + [2D][BPB] staging in 227 KB
    constexpr int BIG = (D <= 4) ? 256 : 128;
    const int arch = a.net.arch, HP = a.net.HP;
    if (arch == CMCD_ARCH_NONE) return launch_bwd_t<D, ACT_GELU, 0, 8, 128>(a, st, num_sms, cot, out, ws, ws_bytes);
    if (arch == CMCD_ARCH_DDS) {
        if (HP == 64) return launch_bwd_t<D, ACT_GELU, 64, 64, BIG>(a, st, num_sms, cot, out, ws, ws_bytes);
        return launch_bwd_t<D, ACT_GELU, 0, 8, 64>(a, st, num_sms, cot, out, ws, ws_bytes);
    }
    if (HP == 64) return launch_bwd_t<D, ACT_SOFTPLUS, 64, 64, BIG>(a, st, num_sms, cot, out, ws, ws_bytes);
    // README.md:30,34 (geffner, emb_dim 130, d = 2 -> hidden_pad 136): two register chunks of 68 output units per layer-2 pass
    // (17 broadcast LDS.128 per 68 FMAs instead of 2 per 8 with the generic JC = 8 chunks)
    if (HP == 136) return launch_bwd_t<D, ACT_SOFTPLUS, 136, 68, 64>(a, st, num_sms, cot, out, ws, ws_bytes);
    if (HP < 64) return launch_bwd_t<D, ACT_SOFTPLUS, 0, 8, BIG>(a, st, num_sms, cot, out, ws, ws_bytes);
    return launch_bwd_t<D, ACT_SOFTPLUS, 0, 8, 64>(a, st, num_sms, cot, out, ws, ws_bytes);
}

size_t bridge_bwd_workspace_bytes(int D, int K, int HP, int arch, int num_sms) {
    const BwdLayout L = make_layout(D, K, HP, arch);
    return (size_t)num_sms * L.P * sizeof(float);
}

int launch_bridge_bwd(const BridgeArgs& a, int D, cudaStream_t st, int num_sms, const float* cot_negw,
                      float* g_vd_mean, float* g_vd_logdiag, float* g_betas, float* g_eps,
                      const cmcd_net_grad* g, void* ws, size_t ws_bytes) {
    BwdOut o{};
    if (g && a.net.arch != CMCD_ARCH_NONE) {
        o.W2 = g->W2; o.U1 = g->U1; o.U2 = g->U2; o.U3 = g->U3; o.W3 = g->W3;
        o.c1 = g->c1; o.c2 = g->c2; o.c3 = g->c3; o.os = g->out_scale;
    }
    o.beta = g_betas; o.eps = g_eps; o.mu = g_vd_mean; o.ls = g_vd_logdiag;
    switch (D) {
        case 2: return launch_bwd_d<2>(a, st, num_sms, cot_negw, o, ws, ws_bytes);
        case 10: return launch_bwd_d<10>(a, st, num_sms, cot_negw, o, ws, ws_bytes);
        default:
            set_error("bridge_bwd: dim=%d has no small-d instantiation (supported: 2, 10)", D);
            return 2;
    }
}

}  // namespace cmcd
