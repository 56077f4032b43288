// Underdamped operators in the block-cooperative mapping (see bridge_blk.cu for the mapping, bridge_ud.cu for the step, its
// coefficient rows and the adjoint algebra; replaces the same reference code as bridge_ud.cu: src/mcd_under_lp_a.py,
// src/mcd_under_lp_e.py, src/mcd_under_lp_ea.py, src/mcd_under_lp_a_cais.py under src/mcdboundingmachine.py:126-205 and
// jax.grad, src/main.py:174-176).  A translation unit of its own so that the two halves compile in parallel.
#include "blk_net.cuh"

namespace cmcd {

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
        many_gmm_stage_means(a.tgt, sMu, tid, blockDim.x);
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
        many_gmm_stage_means(a.tgt, sMu, tid, blockDim.x);
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
