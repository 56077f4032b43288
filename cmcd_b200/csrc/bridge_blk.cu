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
#include "blk_net.cuh"

namespace cmcd {

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
        many_gmm_stage_means(a.tgt, sMu, tid, blockDim.x);
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
        many_gmm_stage_means(a.tgt, sMu, tid, blockDim.x);
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

// ---------------------------------------------------------------------------------------------------------------- launchers
static size_t blk_fwd_smem(int D, int HP) {
    return (net_smem_floats(D, HP) + MIX_MAX * MIX_STRIDE + 2 * MIX_MAX + 2 * (size_t)HP * BK_RS + 2 * (size_t)D * BK_P + (size_t)(HP / 8) * D * BK_P + 8) * sizeof(float);
}
static size_t blk_bwd_smem(int D, int HP) {
    return (net_smem_floats(D, HP) + (size_t)HP * HP + MIX_MAX * MIX_STRIDE + 2 * MIX_MAX + 3 * (size_t)HP * BK_RS + 4 * (size_t)D * BK_P + (size_t)(HP / 8) * D * BK_P + 8) * sizeof(float);
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

}  // namespace cmcd
