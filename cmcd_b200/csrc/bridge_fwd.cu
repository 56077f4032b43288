// Forward bridge kernel: one thread per particle, persistent over particle tiles.
//
// Replaces the XLA program of vmap(compute_log_elbo) (src/mcdboundingmachine.py:126-205):
// key chain + z0 ~ q (:151-157), the lax.scan over nbridges steps of
// src/mcd_cais.py:46-89 / src/mcd_cais_var.py:56-101 / src/mcd_over_orig.py:18-55, and
// w + log p(z_K) (:178).  Particle state (z, w, key) stays in registers for all K steps; the
// only HBM traffic is seeds in, (-w, z_K) out and, for training, the z_k trajectory.
#include "net.cuh"

namespace cmcd {

constexpr int FWD_PB = 128;  // particles (= threads) per block

// numpyro Normal.log_prob summed over dims (src/mcd_utils.py:19-21)
template <int D>
__device__ __forceinline__ float gauss_logprob(const float (&x)[D], const float (&mean)[D], float scale, float lognorm) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        const float v = (x[j] - mean[j]) / scale;
        s += -0.5f * v * v - lognorm;
    }
    return s;
}

template <int D, int ACT, int HPT, int JC>
__global__ void __launch_bounds__(FWD_PB, (HPT > 64 ? 1 : (D <= 4 ? 4 : 2))) bridge_fwd_kernel(const BridgeArgs a) {   // wide nets: shared memory allows one CTA per SM anyway
    extern __shared__ float4 smem4[];
    float* sm = reinterpret_cast<float*>(smem4);
    const int tid = threadIdx.x;
    const NetView& nv = a.net;
    const int HP = HPT ? HPT : nv.HP;
    NetSmem ns = net_stage_smem(nv, D, sm);
    float* sTp = sm + (nv.arch != CMCD_ARCH_NONE ? net_smem_floats(D, HP) : 0);
    const int ntp = (a.tgt.kind == TGT_GMM || a.tgt.kind == TGT_MANY_GMM) ? a.tgt.ncomp * MIX_STRIDE : 0;
    for (int i = tid; i < ntp; i += blockDim.x) sTp[i] = a.tgt.mix[i];
    float* a1col = sTp + ((ntp + 3) & ~3) + tid;
    __syncthreads();

    const bool cais = (a.mode == CMCD_MODE_CAIS_SN || a.mode == CMCD_MODE_CAIS_VAR_SN);
    const bool nn_b = (a.mode != CMCD_MODE_ULA) && nv.arch != CMCD_ARCH_NONE;
    const bool nn_f = cais && nv.arch != CMCD_ARCH_NONE;
    const int K = a.K;

    float mu[D], sig[D];
#pragma unroll
    for (int j = 0; j < D; ++j) { mu[j] = a.vd_mean[j]; sig[j] = expf(a.vd_logdiag[j]); }

    const long long ntiles = (a.N + FWD_PB - 1) / FWD_PB;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long n = tile * FWD_PB + tid;
        if (n >= a.N) continue;
        const bool ev = a.z0 != nullptr;   // mcd_utils.evolve entry: (z, rng_key_gen) given by the caller (mcd_utils.py:24-33)
        Key k, ka;
        float z[D], zn[D], xi[D];
        float w = 0.f;
        if (ev) {
#pragma unroll
            for (int j = 0; j < D; ++j) z[j] = a.z0[n * D + j];
            ka.k0 = a.keys[2 * n]; ka.k1 = a.keys[2 * n + 1];
        } else {
            k = prng_key(a.seeds[n]);
            split(k, ka, k);
            normal_vec<D>(ka, xi);
        }
        if (!ev) {   // z0 = sigma*xi + mu ; w = -log q(z0)   (vardist/diag_gauss.py:26-33,44-62)
            float lq = 0.f;
#pragma unroll
            for (int j = 0; j < D; ++j) {
                z[j] = sig[j] * xi[j] + mu[j];
                const float v = (z[j] - mu[j]) / sig[j];
                lq += -0.5f * v * v - logf(2.5066282746310002f * sig[j]);
            }
            w = -lq;
        }
        if (a.traj) {
#pragma unroll
            for (int j = 0; j < D; ++j) a.traj[((size_t)0 * D + j) * a.N + n] = z[j];
        }
        float sp[D], dummy[D];
        float lp = target_eval<D, false>(a.tgt, sTp, z, sp, dummy, dummy);
        if (K >= 1) {
            if (!ev) ka = split_first(k);    // mcdboundingmachine.py:162 (evolve entry: ka is the caller's rng_key_gen)
            k = split_second(ka);            // mcd_cais.py:94
            float wm = 0.f;
            // One network evaluation per trajectory point: in the CAIS modes NN(z', i + 1) of step i's backward-kernel mean
            // (mcd_cais.py:78) is the same evaluation as NN(z, i + 1) of step i + 1's forward-kernel mean (mcd_cais.py:60),
            // so it is carried over in nnv: K + 1 evaluations instead of the reference's 2K.
            float nnv[D];
#pragma unroll
            for (int j = 0; j < D; ++j) nnv[j] = 0.f;
            if (nn_f) net_fwd<D, ACT, HPT, JC, FWD_PB>(nv, ns, 0, z, nnv, a1col);
            for (int i = 0; i < K; ++i) {
                const float beta = __ldg(a.betas + i), eps = __ldg(a.eps + i);
                float mf[D], mb[D];
                // forward kernel mean
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    const float sq = -((z[j] - mu[j]) / sig[j]) / sig[j];
                    const float gu = fminf(fmaxf(sp[j], -a.clip_t), a.clip_t);
                    const float gq = fminf(fmaxf(sq, -a.clip_q), a.clip_q);
                    const float uf = -(beta * gu + (1.0f - beta) * gq);
                    mf[j] = z[j] - eps * uf;
                }
                if (nn_f) {   // nnv = NN(z_i, i), evaluated at the end of the previous step (or before the loop for i = 0)
#pragma unroll
                    for (int j = 0; j < D; ++j) mf[j] = mf[j] - eps * nnv[j];
                }
                const float scale = sqrtf(2.0f * eps);
                split(k, ka, k);
                normal_vec<D>(ka, xi);
#pragma unroll
                for (int j = 0; j < D; ++j) zn[j] = mf[j] + scale * xi[j];
                // backward kernel mean
                lp = target_eval<D, false>(a.tgt, sTp, zn, sp, dummy, dummy);
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    const float sq = -((zn[j] - mu[j]) / sig[j]) / sig[j];
                    const float gu = fminf(fmaxf(sp[j], -a.clip_t), a.clip_t);
                    const float gq = fminf(fmaxf(sq, -a.clip_q), a.clip_q);
                    const float ub = -(beta * gu + (1.0f - beta) * gq);
                    mb[j] = zn[j] - eps * ub;
                }
                if (nn_b) {
                    net_fwd<D, ACT, HPT, JC, FWD_PB>(nv, ns, cais ? i + 1 : i, zn, nnv, a1col);
#pragma unroll
                    for (int j = 0; j < D; ++j) mb[j] = mb[j] + eps * nnv[j];
                }
                const float lognorm = logf(2.5066282746310002f * scale);
                const float fk = gauss_logprob<D>(zn, mf, scale, lognorm);
                const float bk = gauss_logprob<D>(z, mb, scale, lognorm);
                wm += bk - fk;
                k = split_second(k);
#pragma unroll
                for (int j = 0; j < D; ++j) z[j] = zn[j];
                if (a.traj) {
#pragma unroll
                    for (int j = 0; j < D; ++j) a.traj[((size_t)(i + 1) * D + j) * a.N + n] = z[j];
                }
            }
            w += wm;
        }
        // evolve returns the steps' log-ratio sum w (mcd_cais.py:98-99); compute_log_elbo adds log p(z_K) (mcdboundingmachine.py:178)
        a.out_negw[n] = ev ? w : -(w + lp);
#pragma unroll
        for (int j = 0; j < D; ++j) a.out_z[n * D + j] = z[j];
    }
}

template <int D, int ACT, int HPT, int JC>
static int launch_fwd_t(const BridgeArgs& a, cudaStream_t st, int num_sms) {
    const int HP = a.net.HP;
    size_t fl = (a.net.arch != CMCD_ARCH_NONE ? net_smem_floats(D, HP) + (size_t)HP * FWD_PB : 0) + MIX_MAX * MIX_STRIDE + 8;
    const size_t smem = fl * sizeof(float);
    auto kern = bridge_fwd_kernel<D, ACT, HPT, JC>;
    if (smem > 227 * 1024) { set_error("bridge_fwd: hidden_pad=%d needs %zu B shared memory (> 227 KB)", HP, smem); return 2; }
    CMCD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CMCD_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, FWD_PB, smem));
    if (occ < 1) { set_error("bridge_fwd: kernel does not fit on an SM"); return 2; }
    const long long ntiles = (a.N + FWD_PB - 1) / FWD_PB;
    long long grid = (long long)num_sms * occ;
    if (grid > ntiles) grid = ntiles;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, FWD_PB, smem, st>>>(a);
    CMCD_CUDA_OK(cudaGetLastError());
    return 0;
}

template <int D>
static int launch_fwd_d(const BridgeArgs& a, cudaStream_t st, int num_sms) {
    if (a.net.arch == CMCD_ARCH_NONE) return launch_fwd_t<D, ACT_GELU, 0, 8>(a, st, num_sms);
    if (a.net.arch == CMCD_ARCH_DDS) {
        if (a.net.HP == 64) return launch_fwd_t<D, ACT_GELU, 64, 64>(a, st, num_sms);
        return launch_fwd_t<D, ACT_GELU, 0, 8>(a, st, num_sms);
    }
    if (a.net.HP == 64) return launch_fwd_t<D, ACT_SOFTPLUS, 64, 64>(a, st, num_sms);
    if (a.net.HP == 136) return launch_fwd_t<D, ACT_SOFTPLUS, 136, 68>(a, st, num_sms);   // README.md:30,34 (emb_dim 130, d = 2)
    return launch_fwd_t<D, ACT_SOFTPLUS, 0, 8>(a, st, num_sms);
}

int launch_bridge_fwd(const BridgeArgs& a, int D, cudaStream_t st, int num_sms) {
    switch (D) {
        case 2: return launch_fwd_d<2>(a, st, num_sms);
        case 10: return launch_fwd_d<10>(a, st, num_sms);
        default:
            set_error("bridge_fwd: dim=%d has no small-d instantiation (supported: 2, 10; 1600 via the lgcp wide path)", D);
            return 2;
    }
}

}  // namespace cmcd
