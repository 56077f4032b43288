// UHA (Uncorrected Hamiltonian Annealing) bridge kernels: forward and reverse mode, one thread per particle.
//
// Replaces the XLA program of vmap(boundingmachine.compute_log_elbo) (src/boundingmachine.py:73-111) with nbridges >= 1,
// i.e. ais_utils.evolve (src/ais_utils.py:7-69) with the diagonal momentum distribution of src/momdist.py, and its
// jax.grad (src/main.py:129-131) -- boundmode "UHA" (src/main.py:115-133).  No score network.
// Per particle, sigma_m = exp(md) (momentum scales), a = eta, s = sqrt(1 - eta^2):
//   rho = sigma_m xi_r                                              (ais_utils.py:61-62 -> momdist.py:17-19)
//   bridge i:  rho_r = a rho + s sigma_m xi                         (ais_utils.py:15-16 -> momdist.py:21)
//              leapfrog, lfsteps = L (ais_utils.py:27-56):  p = rho_r - eps gradU(z)/2;  z += eps p / sigma_m^2;
//                        (L-1) x [p -= eps gradU(z);  z += eps p / sigma_m^2];  rho_new = p - eps gradU(z)/2
//              w += log N(rho_new; 0, sigma_m) - log N(rho_r; 0, sigma_m)                         (ais_utils.py:21)
//   (+ log p(z_K) - log q(z_0) in compute_log_elbo); gradU(z) = -(beta_i grad log p(z) + (1 - beta_i) grad log q(z)).
// The key chain is the one of the underdamped MCD operators (one split for the initial momentum, one discarded, two per
// bridge).  delta_H (ais_utils.py:54, a diagnostic that compute_bound drops, boundingmachine.py:107-111) is not produced.
//
// ABI conventions for CMCD_MODE_UHA (include/cmcd_b200.h): eps / g_eps = [3][K] rows (eps, a, s); vd_logdiag /
// g_vd_logdiag = [2][d] = (q log-scales, md); desc.lfsteps = L (1..UHA_LMAX); traj = [K+1][3d][N] = (z_j, rho_j, rho_r_j).
//
// Adjoint of bridge i (c = dL/dw; zb, rb = cotangents of z_{i+1}, rho_{i+1}):  rb += -c rho_new / sigma_m^2;
//   rrb = c rho_r / sigma_m^2;  d md += c ((rho_new/sigma_m)^2 - (rho_r/sigma_m)^2);  then the leapfrog backwards --
//   every kick p' = p - k gradU(z) gives  gb = -k pb,  zb += -beta H_p(z) gb + (1-beta) gb / sigma^2  and the beta / eps / vd
//   terms; every drift z' = z + eps p / sigma_m^2 gives  pb += eps zb / sigma_m^2,  d eps += zb.p / sigma_m^2,
//   d md += -2 eps zb p / sigma_m^2 -- and the refresh:  rb = a rrb,  d a = rrb.rho,  d s = rrb.(rho_r - a rho)/s,
//   d md += rrb (rho_r - a rho).
#include "common.cuh"

namespace cmcd {

constexpr int UHA_PB = 128;
constexpr int UHA_LMAX = 8;

__device__ __forceinline__ float uha_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// log N(x; 0, sigma) summed over dims (momdist.py:25-29 -> numpyro Normal.log_prob)
template <int D>
__device__ __forceinline__ float uha_md_logprob(const float (&x)[D], const float (&sm)[D]) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        const float v = (x[j] - 0.0f) / sm[j];
        s += -0.5f * v * v - logf(2.5066282746310002f * sm[j]);
    }
    return s;
}

template <int D>
__global__ void __launch_bounds__(UHA_PB, 4) bridge_uha_fwd_kernel(const BridgeArgs a, int L) {
    extern __shared__ float4 smem4[];
    float* sTp = reinterpret_cast<float*>(smem4);
    const int tid = threadIdx.x;
    const int ntp = (a.tgt.kind == TGT_GMM || a.tgt.kind == TGT_MANY_GMM) ? a.tgt.ncomp * MIX_STRIDE : 0;
    for (int i = tid; i < ntp; i += blockDim.x) sTp[i] = a.tgt.mix[i];
    __syncthreads();
    const int K = a.K;
    const size_t TS = (size_t)3 * D;

    float mu[D], sig[D], sm[D];
#pragma unroll
    for (int j = 0; j < D; ++j) { mu[j] = a.vd_mean[j]; sig[j] = expf(a.vd_logdiag[j]); sm[j] = expf(a.vd_logdiag[D + j]); }

    const long long ntiles = (a.N + UHA_PB - 1) / UHA_PB;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long n = tile * UHA_PB + tid;
        if (n >= a.N) continue;
        Key k = prng_key(a.seeds[n]);
        Key ka;
        split(k, ka, k);                     // boundingmachine.py:87
        float z[D], xi[D], rho[D], rr[D], p[D];
        normal_vec<D>(ka, xi);
        float w = 0.f;
        {
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
            Key g = split_first(k);          // boundingmachine.py:94
            split(g, ka, g);                 // ais_utils.py:61
            normal_vec<D>(ka, xi);
#pragma unroll
            for (int j = 0; j < D; ++j) rho[j] = sm[j] * xi[j];   // momdist.py:17-19
            g = split_second(g);             // ais_utils.py:64
            float wm = 0.f;
            for (int i = 0; i < K; ++i) {
                const float beta = __ldg(a.betas + i), eps = __ldg(a.eps + i);
                const float ca = __ldg(a.eps + K + i), cs = __ldg(a.eps + 2 * K + i);
                step_keys_and_normal<D>(g, xi);   // ais_utils.py:15 and :22
#pragma unroll
                for (int j = 0; j < D; ++j) rr[j] = ca * rho[j] + cs * (sm[j] * xi[j]);   // momdist.py:21
                if (a.traj) {
#pragma unroll
                    for (int j = 0; j < D; ++j) {
                        a.traj[((size_t)i * TS + j) * a.N + n] = z[j];
                        a.traj[((size_t)i * TS + D + j) * a.N + n] = rho[j];
                        a.traj[((size_t)i * TS + 2 * D + j) * a.N + n] = rr[j];
                    }
                }
#pragma unroll
                for (int j = 0; j < D; ++j) {   // half kick + first drift
                    const float sq = -((z[j] - mu[j]) / sig[j]) / sig[j];
                    const float gU = -(beta * sp[j] + (1.0f - beta) * sq);
                    p[j] = rr[j] - eps * gU / 2.0f;
                    z[j] = z[j] + eps * ((p[j] / sm[j]) / sm[j]);
                }
                for (int l = 1; l < L; ++l) {   // alternate full steps (ais_utils.py:31-36,46-49)
                    lp = target_eval<D, false>(a.tgt, sTp, z, sp, dummy, dummy);
#pragma unroll
                    for (int j = 0; j < D; ++j) {
                        const float sq = -((z[j] - mu[j]) / sig[j]) / sig[j];
                        const float gU = -(beta * sp[j] + (1.0f - beta) * sq);
                        p[j] = p[j] - eps * gU;
                        z[j] = z[j] + eps * ((p[j] / sm[j]) / sm[j]);
                    }
                }
                lp = target_eval<D, false>(a.tgt, sTp, z, sp, dummy, dummy);
#pragma unroll
                for (int j = 0; j < D; ++j) {   // final half kick
                    const float sq = -((z[j] - mu[j]) / sig[j]) / sig[j];
                    const float gU = -(beta * sp[j] + (1.0f - beta) * sq);
                    rho[j] = p[j] - eps * gU / 2.0f;
                }
                wm = wm + uha_md_logprob<D>(rho, sm) - uha_md_logprob<D>(rr, sm);   // ais_utils.py:21
            }
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

// one block's partial slice: beta [K] | rows [3K] | mu [D] | ls [2D]
struct UhaLayout { int beta, rows, mu, ls, P; };
static UhaLayout uha_layout(int D, int K) {
    UhaLayout l;
    const int k = K > 0 ? K : 1;
    l.beta = 0; l.rows = k; l.mu = 4 * k; l.ls = 4 * k + D; l.P = (4 * k + 3 * D + 3) & ~3;
    return l;
}

template <int D>
__global__ void __launch_bounds__(UHA_PB, 2) bridge_uha_bwd_kernel(const BridgeArgs a, int L, const float* __restrict__ cot_negw,
                                                                   float* __restrict__ partials, const UhaLayout Y) {
    extern __shared__ float4 smem4[];
    float* sTp = reinterpret_cast<float*>(smem4);
    const int tid = threadIdx.x;
    const int ntp = (a.tgt.kind == TGT_GMM || a.tgt.kind == TGT_MANY_GMM) ? a.tgt.ncomp * MIX_STRIDE : 0;
    for (int i = tid; i < ntp; i += blockDim.x) sTp[i] = a.tgt.mix[i];
    __syncthreads();
    float* part = partials + (size_t)blockIdx.x * Y.P;
    const int K = a.K;
    const size_t TS = (size_t)3 * D;

    float mu[D], sig[D], ivar[D], sm[D], im2[D];
#pragma unroll
    for (int j = 0; j < D; ++j) {
        mu[j] = a.vd_mean[j]; sig[j] = expf(a.vd_logdiag[j]); ivar[j] = 1.0f / (sig[j] * sig[j]);
        sm[j] = expf(a.vd_logdiag[D + j]); im2[j] = 1.0f / (sm[j] * sm[j]);
    }

    const long long ntiles = (a.N + UHA_PB - 1) / UHA_PB;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long n_raw = tile * UHA_PB + tid;
        const bool active = n_raw < a.N;
        const long long n = active ? n_raw : a.N - 1;   // tail lanes shadow the last particle with zero cotangent
        const float c = active ? -cot_negw[n] : 0.f;    // dL/dw_n
        float zb[D], rb[D], gmu[D], gls[D], gmd[D], sp[D], hv[D], zero[D], rnew[D], zK[D];
#pragma unroll
        for (int j = 0; j < D; ++j) {
            zK[j] = a.traj[((size_t)K * TS + j) * a.N + n];
            rnew[j] = a.traj[((size_t)K * TS + D + j) * a.N + n];
            gmu[j] = 0.f; gls[j] = 0.f; gmd[j] = 0.f; zero[j] = 0.f; rb[j] = 0.f;
        }
        target_eval<D, false>(a.tgt, sTp, zK, sp, zero, hv);
#pragma unroll
        for (int j = 0; j < D; ++j) zb[j] = c * sp[j];   // w += log p(z_K) (boundingmachine.py:101)

        for (int i = K - 1; i >= 0; --i) {
            const float beta = __ldg(a.betas + i), eps = __ldg(a.eps + i);
            const float ca = __ldg(a.eps + K + i), cs = __ldg(a.eps + 2 * K + i);
            const float omb = 1.0f - beta, he = 0.5f * eps;
            float zs[UHA_LMAX + 1][D], ps[UHA_LMAX][D], rho[D], rr[D];
#pragma unroll
            for (int j = 0; j < D; ++j) {
                zs[0][j] = a.traj[((size_t)i * TS + j) * a.N + n];
                rho[j] = a.traj[((size_t)i * TS + D + j) * a.N + n];
                rr[j] = a.traj[((size_t)i * TS + 2 * D + j) * a.N + n];
            }
            // ---- recompute the leapfrog positions and momenta (same operation order as the forward kernel)
            for (int l = 0; l < L; ++l) {
                target_eval<D, false>(a.tgt, sTp, zs[l], sp, zero, hv);
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    const float sq = -((zs[l][j] - mu[j]) / sig[j]) / sig[j];
                    const float gU = -(beta * sp[j] + omb * sq);
                    ps[l][j] = (l == 0) ? rr[j] - eps * gU / 2.0f : ps[l - 1][j] - eps * gU;
                    zs[l + 1][j] = zs[l][j] + eps * ((ps[l][j] / sm[j]) / sm[j]);
                }
            }
            float gbeta = 0.f, geps = 0.f, gca = 0.f, gcs = 0.f;
            // ---- weight: w += log N(rho_new; 0, sigma_m) - log N(rho_r; 0, sigma_m)   (rnew = rho_{i+1}, stored)
            float pb[D], rrb[D];
#pragma unroll
            for (int j = 0; j < D; ++j) {
                const float vn = rnew[j] / sm[j], vr = rr[j] / sm[j];
                pb[j] = rb[j] - c * (vn / sm[j]);
                rrb[j] = c * (vr / sm[j]);
                gmd[j] = fmaf(c, vn * vn - vr * vr, gmd[j]);
            }
            // ---- leapfrog backwards: kick at zs[L] (eps/2), then (drift, kick) pairs down to zs[0]
            for (int l = L; l >= 0; --l) {
                if (l < L) {   // drift zs[l+1] = zs[l] + eps ps[l] / sigma_m^2
#pragma unroll
                    for (int j = 0; j < D; ++j) {
                        pb[j] = fmaf(eps * im2[j], zb[j], pb[j]);
                        geps = fmaf(zb[j], ps[l][j] * im2[j], geps);
                        gmd[j] = fmaf(-2.0f * eps * im2[j] * ps[l][j], zb[j], gmd[j]);
                    }
                }
                const bool half = (l == 0 || l == L);
                const float ck = half ? he : eps;
                float gb[D];
#pragma unroll
                for (int j = 0; j < D; ++j) gb[j] = -ck * pb[j];
                target_eval<D, true>(a.tgt, sTp, zs[l], sp, gb, hv);
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    const float sq = -(zs[l][j] - mu[j]) * ivar[j];
                    const float gU = -(beta * sp[j] + omb * sq);
                    zb[j] = zb[j] - beta * hv[j] + omb * ivar[j] * gb[j];
                    geps = fmaf((half ? -0.5f : -1.0f) * pb[j], gU, geps);
                    gbeta = fmaf(gb[j], -(sp[j] - sq), gbeta);
                    gmu[j] = fmaf(-omb * ivar[j], gb[j], gmu[j]);
                    gls[j] = fmaf(2.0f * omb * sq, gb[j], gls[j]);
                }
            }
            // ---- momentum refresh rho_r = a rho + s sigma_m xi
#pragma unroll
            for (int j = 0; j < D; ++j) {
                rrb[j] += pb[j];
                const float noise = rr[j] - ca * rho[j];   // s sigma_m xi
                gca = fmaf(rrb[j], rho[j], gca);
                gcs = fmaf(rrb[j], noise / cs, gcs);
                gmd[j] = fmaf(rrb[j], noise, gmd[j]);
                rb[j] = ca * rrb[j];
                rnew[j] = rho[j];
            }
            gbeta = uha_warp_sum(gbeta); geps = uha_warp_sum(geps); gca = uha_warp_sum(gca); gcs = uha_warp_sum(gcs);
            if ((tid & 31) == 0) {
                atomicAdd(part + Y.beta + i, gbeta);
                atomicAdd(part + Y.rows + i, geps);
                atomicAdd(part + Y.rows + K + i, gca);
                atomicAdd(part + Y.rows + 2 * K + i, gcs);
            }
#pragma unroll
            for (int j = 0; j < D; ++j) zK[j] = zs[0][j];
        }
        // initial: rho_0 = sigma_m xi_r (rnew = rho_0 here), z_0 = mu + sigma xi0, w_0 = -log q(z_0)   (zK = z_0 here)
#pragma unroll
        for (int j = 0; j < D; ++j) {
            if (K >= 1) gmd[j] = fmaf(rb[j], rnew[j], gmd[j]);
            gmu[j] += zb[j];
            gls[j] += zb[j] * (zK[j] - mu[j]) + c;
            const float m1 = uha_warp_sum(gmu[j]), m2 = uha_warp_sum(gls[j]), m3 = uha_warp_sum(gmd[j]);
            if ((tid & 31) == 0) {
                atomicAdd(part + Y.mu + j, m1);
                atomicAdd(part + Y.ls + j, m2);
                atomicAdd(part + Y.ls + D + j, m3);
            }
        }
    }
}

struct UhaOut { float *beta, *rows, *mu, *ls; };
__global__ void uha_reduce_kernel(const float* __restrict__ partials, int nblocks, UhaLayout Y, UhaOut o, int D, int K) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= Y.P) return;
    float s = 0.f;
    for (int b = 0; b < nblocks; ++b) s += partials[(size_t)b * Y.P + k];
    auto put = [&](float* dst, int off, int len) { if (dst && k >= off && k < off + len) dst[k - off] = s; };
    put(o.beta, Y.beta, K); put(o.rows, Y.rows, 3 * K); put(o.mu, Y.mu, D); put(o.ls, Y.ls, 2 * D);
}

static size_t uha_smem() { return (size_t)(MIX_MAX * MIX_STRIDE + 8) * sizeof(float); }
static int uha_grid(long long N, int num_sms, int per_sm) {
    const long long ntiles = (N + UHA_PB - 1) / UHA_PB;
    long long g = (long long)num_sms * per_sm;
    if (g > ntiles) g = ntiles;
    return (int)(g < 1 ? 1 : g);
}

template <int D>
static int launch_uha_fwd_d(const BridgeArgs& a, int L, cudaStream_t st, int num_sms) {
    bridge_uha_fwd_kernel<D><<<uha_grid(a.N, num_sms, 4), UHA_PB, uha_smem(), st>>>(a, L);
    CMCD_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_bridge_uha_fwd(const BridgeArgs& a, int D, int lfsteps, cudaStream_t st, int num_sms) {
    if (lfsteps < 1 || lfsteps > UHA_LMAX) { set_error("bridge_uha: lfsteps=%d outside 1..%d", lfsteps, UHA_LMAX); return 2; }
    switch (D) {
        case 2: return launch_uha_fwd_d<2>(a, lfsteps, st, num_sms);
        case 10: return launch_uha_fwd_d<10>(a, lfsteps, st, num_sms);
        default: set_error("bridge_uha_fwd: dim=%d has no instantiation (supported: 2, 10)", D); return 2;
    }
}

size_t bridge_uha_bwd_workspace_bytes(int D, int K, int num_sms) {
    return (size_t)num_sms * 2 * uha_layout(D, K).P * sizeof(float);
}

template <int D>
static int launch_uha_bwd_d(const BridgeArgs& a, int L, cudaStream_t st, int num_sms, const float* cot, const UhaOut& o,
                            void* ws, size_t ws_bytes) {
    const UhaLayout Y = uha_layout(D, a.K);
    const int grid = uha_grid(a.N, num_sms, 2);
    const size_t need = (size_t)grid * Y.P * sizeof(float);
    if (ws_bytes < need || !ws) { set_error("bridge_uha_bwd: workspace too small (%zu < %zu)", ws_bytes, need); return 2; }
    CMCD_CUDA_OK(cudaMemsetAsync(ws, 0, need, st));
    bridge_uha_bwd_kernel<D><<<grid, UHA_PB, uha_smem(), st>>>(a, L, cot, (float*)ws, Y);
    CMCD_CUDA_OK(cudaGetLastError());
    uha_reduce_kernel<<<(Y.P + 255) / 256, 256, 0, st>>>((const float*)ws, grid, Y, o, D, a.K);
    CMCD_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_bridge_uha_bwd(const BridgeArgs& a, int D, int lfsteps, cudaStream_t st, int num_sms, const float* cot_negw,
                          float* g_vd_mean, float* g_vd_logdiag, float* g_betas, float* g_eps, void* ws, size_t ws_bytes) {
    if (lfsteps < 1 || lfsteps > UHA_LMAX) { set_error("bridge_uha: lfsteps=%d outside 1..%d", lfsteps, UHA_LMAX); return 2; }
    UhaOut o{g_betas, g_eps, g_vd_mean, g_vd_logdiag};
    switch (D) {
        case 2: return launch_uha_bwd_d<2>(a, lfsteps, st, num_sms, cot_negw, o, ws, ws_bytes);
        case 10: return launch_uha_bwd_d<10>(a, lfsteps, st, num_sms, cot_negw, o, ws, ws_bytes);
        default: set_error("bridge_uha_bwd: dim=%d has no instantiation (supported: 2, 10)", D); return 2;
    }
}

}  // namespace cmcd
