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
#include "net_bwd.cuh"

namespace cmcd {

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
    put(o.beta, L.beta, K); put(o.eps, L.eps, K > 0 ? L.mu - L.eps : 0);   // underdamped modes: (eps_i, eta_i) = 2K entries
    put(o.mu, L.mu, D); put(o.ls, L.ls, D);
    (void)T; (void)HP;
}

int launch_bwd_reduce(const float* partials, int nblocks, const BwdLayout& L, const BwdOut& out, int HP, int D, int K, cudaStream_t st) {
    bwd_reduce_kernel<<<(L.P + 255) / 256, 256, 0, st>>>(partials, nblocks, L, out, HP, D, K);
    CMCD_CUDA_OK(cudaGetLastError());
    return 0;
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
    // wide nets at d = 10 (e.g. funnel with emb_dim ~140: hidden_pad 152): W2 plus three [HP][64 + 4] activation arrays pass 227 KB;
    // 32 particles per block fit up to hidden_pad ~190
    if (bwd_smem_bytes<D, ACT_SOFTPLUS, 0, 8, 64>(HP, true) > 227 * 1024)
        return launch_bwd_t<D, ACT_SOFTPLUS, 0, 8, 32>(a, st, num_sms, cot, out, ws, ws_bytes);
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
