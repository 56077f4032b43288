// Underdamped bridge kernels (the momentum-augmented operators of mcd_utils.evolve): forward and reverse mode, one thread
// per particle, FP32-FMA path.
//
// Replaces the XLA program of vmap(compute_log_elbo) (src/mcdboundingmachine.py:126-205) and its jax.grad
// (src/main.py:174-176) when mcd_utils.evolve (src/mcd_utils.py:59-133) dispatches to
//   evolve_underdamped_lp_a   MCD_U_a-lp | MCD_U_a-lp-sna | MCD_U_a-lp-sn ("LDVI")      src/mcd_under_lp_a.py:6-87
//   evolve_underdamped_lp_e   MCD_U_e-lp | MCD_U_e-lp-sna                                src/mcd_under_lp_e.py:6-74
//   evolve_underdamped_lp_ea  MCD_U_ea-lp-sn                                             src/mcd_under_lp_ea.py:6-104
// All three scan bodies are one step with different coefficients (state (z, rho); x = network input):
//   rho0 ~ N(0, I); w = -log N(rho0; 0, 1)
//   step i:  m_f = a_f rho;  rho' = m_f + s_f xi;  rho'' = rho' - eps gradU(z)/2;  z' = z + eps rho'';
//            rho_new = rho'' - eps gradU(z')/2;  m_b = a_b rho' + c_n NN(x, i),  x = (z, rho') | z | -
//            w += log N(rho; m_b, s_b) - log N(rho'; m_f, s_f)
//   w += log N(rho_K; 0, 1)        (+ log p(z_K) - log q(z_0) in compute_log_elbo)
//   lp_a : a_f = a_b = 1 - g eps, s_f = s_b = sqrt(2 g eps), c_n = 2 g eps                (g = gamma)
//   lp_e : a_f = a_b = eta, s_f = s_b = sqrt(1 - eta^2), c_n = 2 (1 - eta)
//   lp_ea: a_f = exp(-g eps), s_f = sqrt(1 - a_f^2), a_b = 1 - g eps, c_n = 2 g eps, s_b = sqrt(2 g eps)
//   evolve_underdamped_lp_a_cais  MCD_CAIS_UHA_sn ("2nd order CMCD", README.md:16)     src/mcd_under_lp_a_cais.py:6-115
//     (CMCD_MODE_UD_CAIS; a TypeError at the reference HEAD -- mcd_utils.py:176-188 passes keywords the function does not
//      take -- built to the body as written): lp_a with eps_i on the cosine schedule, the target score clipped at 1e2, and
//      the network ALSO in the forward-kernel mean, m_f = a_f rho + c_f NN((z, rho), i), c_f = -2 g eps_i (row 7; 0 elsewhere).
// with gradU(z) = -(beta_i grad log p(z) + (1 - beta_i) grad log q(z)), never clipped (these operators take no
// grad_clipping).  The kernel mode only says what the network sees (CMCD_MODE_UD_NONE / _NET_Z / _NET_ZRHO).
//
// ABI conventions for these modes (include/cmcd_b200.h): eps = [7][K] = rows (eps, a_f, s_f, a_b, c_n, s_b, c_f) -- the host
// forms the coefficient rows from (eps, gamma, eta) with differentiable ops, so the cotangents g_eps = [7][K] chain into
// those scalars; traj = [K+1][3d][N] = (z_j, rho_j, rho'_j) per node (rho'_j: the refreshed momentum of step j, stored
// so that the adjoint does not have to walk the key chain backwards).
//
// Adjoint of step i (c = dL/dw; zb', rb' = cotangents of z', rho_new; pathwise in xi):
//   g1b = -(eps/2) rb'                       zb'' = zb' - beta H_p(z') g1b + (1-beta) g1b / sigma^2
//   rb'' = rb' + eps zb''                    g0b = -(eps/2) rb''
//   zb   = zb'' - beta H_p(z) g0b + (1-beta) g0b / sigma^2 + J_z^T v
//   r = (rho - m_b) / s_b^2;  G = c r;  v = c_n G (cotangent of the network output)
//   rbp  = rb'' + a_b G + J_rho'^T v         (cotangent of rho')
//   rb   = -G + a_f rbp
//   d a_f = rbp.rho;  d s_f = rbp.xi + c d / s_f;  d a_b = G.rho';  d c_n = G.NN;  d s_b = c (s_b |r|^2 - d / s_b)
//   CMCD_MODE_UD_CAIS: second pull-back v_f = c_f rbp at (z, rho) -> zb, rb;  d c_f = rbp.NN_f;  clipped scores: the
//   Hessian-vector products take mk o gb (mk = 1 where |s_p| <= clip), beta sees clip(s_p) - s_q.
//   d eps = -rb'.gradU(z')/2 + zb''.rho'' - rb''.gradU(z)/2
//   d beta = -g1b.(s_p(z') - s_q(z')) - g0b.(s_p(z) - s_q(z));  vd: through s_q at both points and z_0 = mu + sigma xi0.
// (log N(rho'; m_f, s_f) = -|xi|^2/2 - d log s_f - const with xi fixed.)
#include "net_bwd.cuh"

namespace cmcd {

constexpr int UD_FWD_PB = 128;

template <int D>
__device__ __forceinline__ float ud_gauss_logprob(const float (&x)[D], const float (&mean)[D], float scale, float lognorm) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        const float v = (x[j] - mean[j]) / scale;
        s += -0.5f * v * v - lognorm;
    }
    return s;
}

// DI: 0 = no network, D = network on z, 2D = network on (z, rho')
template <int D, int ACT, int HPT, int JC, int DI, int PB = UD_FWD_PB>
__global__ void __launch_bounds__(PB, (HPT > 64 ? 1 : 2)) bridge_ud_fwd_kernel(const BridgeArgs a) {
    constexpr int DIN = DI ? DI : D;
    extern __shared__ float4 smem4[];
    float* sm = reinterpret_cast<float*>(smem4);
    const int tid = threadIdx.x;
    const NetView& nv = a.net;
    const int HP = HPT ? HPT : nv.HP;
    const bool has_net = DI != 0 && nv.arch != CMCD_ARCH_NONE;
    const bool nn_fm = has_net && a.mode == CMCD_MODE_UD_CAIS;   // network in the forward-kernel mean too
    NetSmem ns = net_stage_smem(nv, D, sm, DIN);
    float* sTp = sm + (has_net ? net_smem_floats(D, HP, DIN) : 0);
    const int ntp = (a.tgt.kind == TGT_GMM || a.tgt.kind == TGT_MANY_GMM) ? a.tgt.ncomp * MIX_STRIDE : 0;
    for (int i = tid; i < ntp; i += blockDim.x) sTp[i] = a.tgt.mix[i];
    float* a1col = sTp + ((ntp + 3) & ~3) + tid;
    __syncthreads();
    const int K = a.K;
    const size_t TS = (size_t)3 * D;   // trajectory rows per node

    float mu[D], sig[D];
#pragma unroll
    for (int j = 0; j < D; ++j) { mu[j] = a.vd_mean[j]; sig[j] = expf(a.vd_logdiag[j]); }

    const long long ntiles = (a.N + PB - 1) / PB;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long n = tile * PB + tid;
        if (n >= a.N) continue;
        Key k = prng_key(a.seeds[n]);
        Key ka;
        split(k, ka, k);                     // mcdboundingmachine.py:153
        float z[D], zn[D], xi[D], rho[D];
        normal_vec<D>(ka, xi);
        float w = 0.f;
        {   // z0 = sigma*xi + mu ; w = -log q(z0)   (vardist/diag_gauss.py:26-33,44-62)
            float lq = 0.f;
#pragma unroll
            for (int j = 0; j < D; ++j) {
                z[j] = sig[j] * xi[j] + mu[j];
                const float v = (z[j] - mu[j]) / sig[j];
                lq += -0.5f * v * v - logf(2.5066282746310002f * sig[j]);
            }
            w = -lq;
        }
        float sp[D], dummy[D];
        float lp = target_eval<D, false>(a.tgt, sTp, z, sp, dummy, dummy);
        if (K >= 1) {
            Key g = split_first(k);          // mcdboundingmachine.py:162: rng_key handed to evolve as rng_key_gen
            split(g, ka, g);                 // mcd_under_lp_a.py:62
            normal_vec<D>(ka, rho);          // :63
            float zeros[D];
#pragma unroll
            for (int j = 0; j < D; ++j) zeros[j] = 0.f;
            const float ln1 = logf(2.5066282746310002f);   // log(sqrt(2 pi) * 1)
            float wm = 0.f;
            wm = wm - ud_gauss_logprob<D>(rho, zeros, 1.0f, ln1);   // :66-67
            g = split_second(g);             // :70
            for (int i = 0; i < K; ++i) {
                const float beta = __ldg(a.betas + i), eps = __ldg(a.eps + i);
                const float af = __ldg(a.eps + K + i), sf = __ldg(a.eps + 2 * K + i), ab = __ldg(a.eps + 3 * K + i);
                const float cn = __ldg(a.eps + 4 * K + i), sb = __ldg(a.eps + 5 * K + i), cf = __ldg(a.eps + 6 * K + i);
                float mf[D], mb[D], rp[D], rpp[D], rn[D], nnf[D];
#pragma unroll
                for (int j = 0; j < D; ++j) nnf[j] = 0.f;
                if constexpr (DIN > D) {
                    if (nn_fm) {   // mcd_under_lp_a_cais.py:52-56: NN((z, rho), i) in the forward-kernel mean
                        float x[DIN];
#pragma unroll
                        for (int j = 0; j < D; ++j) { x[j] = z[j]; x[D + j] = rho[j]; }
                        net_fwd<D, ACT, HPT, JC, PB, DIN>(nv, ns, i, x, nnf, a1col);
                    }
                }
                step_keys_and_normal<D>(g, xi);   // :31-32 and :59
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    mf[j] = rho[j] * af + cf * nnf[j];
                    rp[j] = mf[j] + sf * xi[j];
                    const float sq = -((z[j] - mu[j]) / sig[j]) / sig[j];
                    const float g0 = -(beta * fminf(fmaxf(sp[j], -a.clip_t), a.clip_t) + (1.0f - beta) * sq);
                    rpp[j] = rp[j] - eps * g0 / 2.0f;
                    zn[j] = z[j] + eps * rpp[j];
                }
                if (a.traj) {
#pragma unroll
                    for (int j = 0; j < D; ++j) {
                        a.traj[((size_t)i * TS + j) * a.N + n] = z[j];
                        a.traj[((size_t)i * TS + D + j) * a.N + n] = rho[j];
                        a.traj[((size_t)i * TS + 2 * D + j) * a.N + n] = rp[j];
                    }
                }
                // backward-kernel mean: the network sees the OLD position and the refreshed momentum (:47-51)
#pragma unroll
                for (int j = 0; j < D; ++j) mb[j] = rp[j] * ab;
                if constexpr (DI != 0) {
                    if (has_net) {
                        float x[DIN], nnv[D];
#pragma unroll
                        for (int j = 0; j < D; ++j) {
                            x[j] = z[j];
                            if constexpr (DIN > D) x[D + j] = rp[j];
                        }
                        net_fwd<D, ACT, HPT, JC, PB, DIN>(nv, ns, i, x, nnv, a1col);
#pragma unroll
                        for (int j = 0; j < D; ++j) mb[j] = mb[j] + cn * nnv[j];
                    }
                }
                lp = target_eval<D, false>(a.tgt, sTp, zn, sp, dummy, dummy);
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    const float sq = -((zn[j] - mu[j]) / sig[j]) / sig[j];
                    const float g1 = -(beta * fminf(fmaxf(sp[j], -a.clip_t), a.clip_t) + (1.0f - beta) * sq);
                    rn[j] = rpp[j] - eps * g1 / 2.0f;
                }
                const float fk = ud_gauss_logprob<D>(rp, mf, sf, logf(2.5066282746310002f * sf));
                const float bk = ud_gauss_logprob<D>(rho, mb, sb, logf(2.5066282746310002f * sb));
                wm += bk - fk;
#pragma unroll
                for (int j = 0; j < D; ++j) { z[j] = zn[j]; rho[j] = rn[j]; }
            }
            wm = wm + ud_gauss_logprob<D>(rho, zeros, 1.0f, ln1);   // :83-84
            w += wm;
            if (a.traj) {
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    a.traj[((size_t)K * TS + j) * a.N + n] = z[j];
                    a.traj[((size_t)K * TS + D + j) * a.N + n] = rho[j];
                    a.traj[((size_t)K * TS + 2 * D + j) * a.N + n] = 0.f;
                }
            }
        } else if (a.traj) {
#pragma unroll
            for (int j = 0; j < D; ++j) {
                a.traj[(size_t)j * a.N + n] = z[j];
                a.traj[((size_t)D + j) * a.N + n] = 0.f;
                a.traj[((size_t)2 * D + j) * a.N + n] = 0.f;
            }
        }
        w += lp;
        a.out_negw[n] = -w;
#pragma unroll
        for (int j = 0; j < D; ++j) a.out_z[n * D + j] = z[j];
    }
}

template <int D, int ACT, int HPT, int JC, int BPB, int DI>
__global__ void __launch_bounds__(BPB, 1) bridge_ud_bwd_kernel(const BridgeArgs a, const float* __restrict__ cot_negw,
                                                               float* __restrict__ partials, const BwdLayout L) {
    constexpr int DIN = DI ? DI : D;
    constexpr int RS = BPB + 4;
    extern __shared__ float4 smem4[];
    float* sm = reinterpret_cast<float*>(smem4);
    const int tid = threadIdx.x;
    const NetView& nv = a.net;
    const int HP = HPT ? HPT : nv.HP;
    const bool has_net = DI != 0 && nv.arch != CMCD_ARCH_NONE;
    const bool nn_fm = has_net && a.mode == CMCD_MODE_UD_CAIS;
    const bool clipped = a.clip_t < 3.0e38f;
    const float out_scale = has_net ? net_out_scale(nv) : 1.0f;
    NetSmem ns = net_stage_smem(nv, D, sm, DIN);
    float* sTp = sm + (has_net ? net_smem_floats(D, HP, DIN) : 0);
    const int ntp = (a.tgt.kind == TGT_GMM || a.tgt.kind == TGT_MANY_GMM) ? a.tgt.ncomp * MIX_STRIDE : 0;
    for (int i = tid; i < ntp; i += blockDim.x) sTp[i] = a.tgt.mix[i];
    float* S1 = sTp + ((ntp + 3) & ~3);
    float* S2 = S1 + (has_net ? (size_t)HP * RS : 0);
    float* S3 = S2 + (has_net ? (size_t)HP * RS : 0);
    float* sX = S3 + (has_net ? (size_t)HP * RS : 0);
    float* sVo = sX + DIN * RS;
    __syncthreads();

    float* part = partials + (size_t)blockIdx.x * L.P;
    const int K = a.K;
    const size_t TS = (size_t)3 * D;

    float mu[D], sig[D], ivar[D];
#pragma unroll
    for (int j = 0; j < D; ++j) { mu[j] = a.vd_mean[j]; sig[j] = expf(a.vd_logdiag[j]); ivar[j] = 1.0f / (sig[j] * sig[j]); }

    const long long ntiles = (a.N + BPB - 1) / BPB;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long n_raw = tile * BPB + tid;
        const bool active = n_raw < a.N;
        const long long n = active ? n_raw : a.N - 1;   // tail lanes shadow the last particle with zero cotangent
        const float c = active ? -cot_negw[n] : 0.f;    // dL/dw_n
        float zn[D], rhoK[D], zb[D], rb[D], gmu[D], gls[D], sp1[D], hv[D], zero[D];
#pragma unroll
        for (int j = 0; j < D; ++j) {
            zn[j] = a.traj[((size_t)K * TS + j) * a.N + n];
            rhoK[j] = a.traj[((size_t)K * TS + D + j) * a.N + n];
            gmu[j] = 0.f; gls[j] = 0.f; zero[j] = 0.f;
        }
        // terminal terms: w += log p(z_K) (mcdboundingmachine.py:178) and, for K >= 1, w += log N(rho_K; 0, 1) (mcd_under_lp_a.py:84)
        target_eval<D, false>(a.tgt, sTp, zn, sp1, zero, hv);
#pragma unroll
        for (int j = 0; j < D; ++j) { zb[j] = c * sp1[j]; rb[j] = (K >= 1) ? -c * rhoK[j] : 0.f; }

        for (int i = K - 1; i >= 0; --i) {
            const float beta = __ldg(a.betas + i), eps = __ldg(a.eps + i);
            const float af = __ldg(a.eps + K + i), sf = __ldg(a.eps + 2 * K + i), ab = __ldg(a.eps + 3 * K + i);
            const float cn = __ldg(a.eps + 4 * K + i), sb = __ldg(a.eps + 5 * K + i), cf = __ldg(a.eps + 6 * K + i);
            const float omb = 1.0f - beta, s2 = sb * sb, he = 0.5f * eps;
            float z[D], rho[D], rp[D];
#pragma unroll
            for (int j = 0; j < D; ++j) {
                z[j] = a.traj[((size_t)i * TS + j) * a.N + n];
                rho[j] = a.traj[((size_t)i * TS + D + j) * a.N + n];
                rp[j] = a.traj[((size_t)i * TS + 2 * D + j) * a.N + n];
            }
            float gbeta = 0.f, geps = 0.f, gaf = 0.f, gsf = 0.f, gab = 0.f, gcn = 0.f, gsb = 0.f, gcf = 0.f;
            // ---- second half kick: rho_new = rho'' - (eps/2) gradU(z')
            float g1b[D], zbn[D], rbpp[D], g0b[D], sp0[D], hvin[D];
#pragma unroll
            for (int j = 0; j < D; ++j) {   // sp1 = score at z' (from the previous iteration / the terminal evaluation): clip mask
                g1b[j] = -he * rb[j];
                hvin[j] = (fabsf(sp1[j]) <= a.clip_t) ? g1b[j] : 0.f;
            }
            target_eval<D, true>(a.tgt, sTp, zn, sp1, hvin, hv);
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
                // ---- drift: z' = z + eps rho''
                rbpp[j] = fmaf(eps, zbn[j], rb[j]);
                g0b[j] = -he * rbpp[j];
            }
            // ---- first half kick: rho'' = rho' - (eps/2) gradU(z)
            if (clipped) {   // the mask needs the score at z before the Hessian-vector product
                target_eval<D, false>(a.tgt, sTp, z, sp0, zero, hv);
#pragma unroll
                for (int j = 0; j < D; ++j) hvin[j] = (fabsf(sp0[j]) <= a.clip_t) ? g0b[j] : 0.f;
            } else {
#pragma unroll
                for (int j = 0; j < D; ++j) hvin[j] = g0b[j];
            }
            target_eval<D, true>(a.tgt, sTp, z, sp0, hvin, hv);
            float zbc[D], rbp[D];
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
            }
            // ---- backward-kernel log-density: recompute NN(x, i), pull back v = c_n c r
            float x[DIN], o[D], nn[D], vv[D], dx[DIN], G[D];
#pragma unroll
            for (int j = 0; j < D; ++j) {
                x[j] = z[j];
                if constexpr (DIN > D) x[D + j] = rp[j];
                o[j] = 0.f; nn[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < DIN; ++j) dx[j] = 0.f;
            if constexpr (DI != 0) {
                if (has_net) {
                    net_fwd_store<D, ACT, HPT, JC, RS, DIN>(nv, ns, i, x, o, S1 + tid, S2 + tid, S3 + tid);
#pragma unroll
                    for (int j = 0; j < D; ++j) nn[j] = out_scale * fminf(fmaxf(o[j], -nv.out_clip), nv.out_clip);
                }
            }
            float rr = 0.f;
#pragma unroll
            for (int j = 0; j < D; ++j) {
                const float mb = rp[j] * ab + cn * nn[j];
                const float r = (rho[j] - mb) / s2;
                G[j] = c * r;
                rr = fmaf(r, r, rr);
                vv[j] = cn * G[j];
                gab = fmaf(G[j], rp[j], gab);
                gcn = fmaf(G[j], nn[j], gcn);
                rbp[j] = fmaf(ab, G[j], rbp[j]);
            }
            gsb = c * (sb * rr - (float)D / sb);
            gsf = c * (float)D / sf;
            if constexpr (DI != 0) {
                if (has_net) {
                    net_bwd<D, ACT, HPT, JC, BPB, DIN>(nv, ns, i, x, o, vv, dx, S1, S2, S3, sX, sVo, part, L);
#pragma unroll
                    for (int j = 0; j < D; ++j) {
                        zbc[j] += dx[j];
                        if constexpr (DIN > D) rbp[j] += dx[D + j];
                    }
                }
            }
            // ---- forward-kernel mean network (CMCD_MODE_UD_CAIS): m_f = a_f rho + c_f NN((z, rho), i), cotangent of m_f = rbp
            float nnf[D], rext[D];
#pragma unroll
            for (int j = 0; j < D; ++j) { nnf[j] = 0.f; rext[j] = 0.f; }
            if constexpr (DIN > D) {
                if (nn_fm) {
                    float of[D], vf[D];
#pragma unroll
                    for (int j = 0; j < D; ++j) { x[j] = z[j]; x[D + j] = rho[j]; }
                    net_fwd_store<D, ACT, HPT, JC, RS, DIN>(nv, ns, i, x, of, S1 + tid, S2 + tid, S3 + tid);
#pragma unroll
                    for (int j = 0; j < D; ++j) {
                        nnf[j] = out_scale * fminf(fmaxf(of[j], -nv.out_clip), nv.out_clip);
                        vf[j] = cf * rbp[j];
                        gcf = fmaf(rbp[j], nnf[j], gcf);
                    }
                    net_bwd<D, ACT, HPT, JC, BPB, DIN>(nv, ns, i, x, of, vf, dx, S1, S2, S3, sX, sVo, part, L);
#pragma unroll
                    for (int j = 0; j < D; ++j) { zbc[j] += dx[j]; rext[j] = dx[D + j]; }
                }
            }
            // ---- momentum refresh: rho' = m_f + s_f xi,  m_f = a_f rho (+ c_f NN_f)
#pragma unroll
            for (int j = 0; j < D; ++j) {
                const float mf = rho[j] * af + cf * nnf[j];
                gsf = fmaf(rbp[j], (rp[j] - mf) / sf, gsf);
                gaf = fmaf(rho[j], rbp[j], gaf);
                rb[j] = fmaf(af, rbp[j], -G[j]) + rext[j];
                zb[j] = zbc[j];
                zn[j] = z[j];
                sp1[j] = sp0[j];
            }
            gbeta = warp_sum_f(gbeta); geps = warp_sum_f(geps);
            gaf = warp_sum_f(gaf); gsf = warp_sum_f(gsf); gab = warp_sum_f(gab); gcn = warp_sum_f(gcn); gsb = warp_sum_f(gsb);
            if (nn_fm) gcf = warp_sum_f(gcf);
            if ((tid & 31) == 0) {
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
        // initial: z0 = mu + sigma xi0, w0 = -log q(z0) = 0.5|xi0|^2 + sum log(sqrt(2pi) sigma); rho0 is pure noise   (zn = z_0 here)
#pragma unroll
        for (int j = 0; j < D; ++j) {
            gmu[j] += zb[j];
            gls[j] += zb[j] * (zn[j] - mu[j]) + c;
            const float m1 = warp_sum_f(gmu[j]), m2 = warp_sum_f(gls[j]);
            if ((tid & 31) == 0) { atomicAdd(part + L.mu + j, m1); atomicAdd(part + L.ls + j, m2); }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------- launchers

template <int D, int ACT, int HPT, int JC, int DI, int PB = UD_FWD_PB>
static int launch_ud_fwd_t(const BridgeArgs& a, cudaStream_t st, int num_sms) {
    constexpr int DIN = DI ? DI : D;
    const int HP = a.net.HP;
    const bool has_net = DI != 0;
    size_t fl = (has_net ? net_smem_floats(D, HP, DIN) + (size_t)HP * PB : 0) + MIX_MAX * MIX_STRIDE + 8;
    const size_t smem = fl * sizeof(float);
    if constexpr (PB > 32 && HPT == 0 && ACT == ACT_SOFTPLUS) {   // wide (z, rho) nets (funnel, emb_dim ~140: hidden_pad 168): 32 particles per block
        if (smem > 227 * 1024) return launch_ud_fwd_t<D, ACT, HPT, JC, DI, 32>(a, st, num_sms);
    }
    auto kern = bridge_ud_fwd_kernel<D, ACT, HPT, JC, DI, PB>;
    if (smem > 227 * 1024) { set_error("bridge_ud_fwd: hidden_pad=%d needs %zu B shared memory (> 227 KB)", HP, smem); return 2; }
    CMCD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CMCD_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, PB, smem));
    if (occ < 1) { set_error("bridge_ud_fwd: kernel does not fit on an SM"); return 2; }
    const long long ntiles = (a.N + PB - 1) / PB;
    long long grid = (long long)num_sms * occ;
    if (grid > ntiles) grid = ntiles;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, PB, smem, st>>>(a);
    CMCD_CUDA_OK(cudaGetLastError());
    return 0;
}

template <int D, int DI>
static int launch_ud_fwd_n(const BridgeArgs& a, cudaStream_t st, int num_sms) {
    if constexpr (DI == 0) {
        return launch_ud_fwd_t<D, ACT_GELU, 0, 8, 0>(a, st, num_sms);
    } else {
        if (a.net.arch == CMCD_ARCH_DDS) {
            if (a.net.HP == 64) return launch_ud_fwd_t<D, ACT_GELU, 64, 64, DI>(a, st, num_sms);
            return launch_ud_fwd_t<D, ACT_GELU, 0, 8, DI>(a, st, num_sms);
        }
        return launch_ud_fwd_t<D, ACT_SOFTPLUS, 0, 8, DI>(a, st, num_sms);
    }
}

template <int D>
static int launch_ud_fwd_d(const BridgeArgs& a, cudaStream_t st, int num_sms) {
    const int din = (a.net.arch == CMCD_ARCH_NONE || a.K < 1) ? 0 : ud_net_in(a.mode, D);
    if (din == 0) return launch_ud_fwd_n<D, 0>(a, st, num_sms);
    if (din == D) return launch_ud_fwd_n<D, D>(a, st, num_sms);
    return launch_ud_fwd_n<D, 2 * D>(a, st, num_sms);
}

int launch_bridge_ud_fwd(const BridgeArgs& a, int D, cudaStream_t st, int num_sms) {
    switch (D) {
        case 2: return launch_ud_fwd_d<2>(a, st, num_sms);
        case 10: return launch_ud_fwd_d<10>(a, st, num_sms);
        default:
            set_error("bridge_ud_fwd: dim=%d has no instantiation (supported: 2, 10)", D);
            return 2;
    }
}


template <int D, int ACT, int HPT, int JC, int BPB, int DI>
static int launch_ud_bwd_t(const BridgeArgs& a, cudaStream_t st, int num_sms, const float* cot, const BwdOut& out,
                           void* ws, size_t ws_bytes) {
    constexpr int DIN = DI ? DI : D;
    constexpr int RS = BPB + 4;
    const int HP = a.net.HP;
    const bool has_net = DI != 0;
    const size_t fl = (has_net ? net_smem_floats(D, HP, DIN) + 3 * (size_t)HP * RS : 0) + MIX_MAX * MIX_STRIDE + (size_t)(DIN + D) * RS + 8;
    const size_t smem = fl * sizeof(float);
    if (smem > 227 * 1024) { set_error("bridge_ud_bwd: hidden_pad=%d needs %zu B shared memory (> 227 KB)", HP, smem); return 2; }
    auto kern = bridge_ud_bwd_kernel<D, ACT, HPT, JC, BPB, DI>;
    CMCD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const BwdLayout L = ud_layout(D, a.K, HP, has_net ? a.net.arch : CMCD_ARCH_NONE, DI);
    const long long ntiles = (a.N + BPB - 1) / BPB;
    int grid = (int)(ntiles < num_sms ? ntiles : num_sms);
    if (grid < 1) grid = 1;
    const size_t need = (size_t)grid * L.P * sizeof(float);
    if (ws_bytes < need || !ws) { set_error("bridge_ud_bwd: workspace too small (%zu < %zu)", ws_bytes, need); return 2; }
    CMCD_CUDA_OK(cudaMemsetAsync(ws, 0, need, st));
    kern<<<grid, BPB, smem, st>>>(a, cot, (float*)ws, L);
    CMCD_CUDA_OK(cudaGetLastError());
    return launch_bwd_reduce((const float*)ws, grid, L, out, HP, D, a.K, st);
}

template <int D, int DI>
static int launch_ud_bwd_n(const BridgeArgs& a, cudaStream_t st, int num_sms, const float* cot, const BwdOut& out,
                           void* ws, size_t ws_bytes) {
    if constexpr (DI == 0) {
        return launch_ud_bwd_t<D, ACT_GELU, 0, 8, 128, 0>(a, st, num_sms, cot, out, ws, ws_bytes);
    } else {
        if (a.net.arch == CMCD_ARCH_DDS) {
            if (a.net.HP == 64) return launch_ud_bwd_t<D, ACT_GELU, 64, 64, 128, DI>(a, st, num_sms, cot, out, ws, ws_bytes);
            return launch_ud_bwd_t<D, ACT_GELU, 0, 8, 64, DI>(a, st, num_sms, cot, out, ws, ws_bytes);
        }
        // wide nets (funnel with emb_dim ~130: the (z, rho) network has hidden_pad 152): W2 plus three [HP][64 + 4] activation arrays pass
        // 227 KB; 32 particles per block fit
        constexpr int DIN = DI ? DI : D;
        const size_t fl64 = net_smem_floats(D, a.net.HP, DIN) + 3 * (size_t)a.net.HP * 68 + MIX_MAX * MIX_STRIDE + (size_t)(DIN + D) * 68 + 8;
        if (fl64 * sizeof(float) > 227 * 1024)
            return launch_ud_bwd_t<D, ACT_SOFTPLUS, 0, 8, 32, DI>(a, st, num_sms, cot, out, ws, ws_bytes);
        return launch_ud_bwd_t<D, ACT_SOFTPLUS, 0, 8, 64, DI>(a, st, num_sms, cot, out, ws, ws_bytes);
    }
}

template <int D>
static int launch_ud_bwd_d(const BridgeArgs& a, cudaStream_t st, int num_sms, const float* cot, const BwdOut& out,
                           void* ws, size_t ws_bytes) {
    const int din = (a.net.arch == CMCD_ARCH_NONE || a.K < 1) ? 0 : ud_net_in(a.mode, D);
    if (din == 0) return launch_ud_bwd_n<D, 0>(a, st, num_sms, cot, out, ws, ws_bytes);
    if (din == D) return launch_ud_bwd_n<D, D>(a, st, num_sms, cot, out, ws, ws_bytes);
    return launch_ud_bwd_n<D, 2 * D>(a, st, num_sms, cot, out, ws, ws_bytes);
}

size_t bridge_ud_bwd_workspace_bytes(int mode, int D, int K, int HP, int arch, int num_sms) {
    const int din = arch == CMCD_ARCH_NONE ? 0 : ud_net_in(mode, D);
    const BwdLayout L = ud_layout(D, K, HP, din ? arch : CMCD_ARCH_NONE, din);
    return (size_t)num_sms * L.P * sizeof(float);
}

int launch_bridge_ud_bwd(const BridgeArgs& a, int D, cudaStream_t st, int num_sms, const float* cot_negw,
                         float* g_vd_mean, float* g_vd_logdiag, float* g_betas, float* g_eps,
                         const cmcd_net_grad* g, void* ws, size_t ws_bytes) {
    BwdOut o{};
    if (g && a.net.arch != CMCD_ARCH_NONE) {
        o.W2 = g->W2; o.U1 = g->U1; o.U2 = g->U2; o.U3 = g->U3; o.W3 = g->W3;
        o.c1 = g->c1; o.c2 = g->c2; o.c3 = g->c3; o.os = g->out_scale;
    }
    o.beta = g_betas; o.eps = g_eps; o.mu = g_vd_mean; o.ls = g_vd_logdiag;
    switch (D) {
        case 2: return launch_ud_bwd_d<2>(a, st, num_sms, cot_negw, o, ws, ws_bytes);
        case 10: return launch_ud_bwd_d<10>(a, st, num_sms, cot_negw, o, ws, ws_bytes);
        default:
            set_error("bridge_ud_bwd: dim=%d has no instantiation (supported: 2, 10)", D);
            return 2;
    }
}

}  // namespace cmcd
