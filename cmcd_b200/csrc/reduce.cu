// Reductions over the per-particle losses and PRNG test hooks.
//
// Replaces batch_log_elbos.mean() / .var(ddof=0) (src/mcdboundingmachine.py:205,231) and the
// ELBO / ln Z estimators of src/utils.py:227-237 (logsumexp(-loss) - log n), which the
// reference evaluates after a per-element .item() host loop (src/opt.py:193).
#include "common.cuh"

namespace cmcd {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// online logsumexp state combine: (m, s) represents s * exp(m)
__device__ __forceinline__ void lse_combine(float& m, float& s, float m2, float s2) {
    const float mm = fmaxf(m, m2);
    if (mm == -CUDART_INF_F) { s = 0.f; m = mm; return; }
    s = s * expf(m - mm) + s2 * expf(m2 - mm);
    m = mm;
}

// One block; out4 = {sum l, sum l^2, max(-l), sum exp(-l - max)}.  Sums in double.
__global__ void __launch_bounds__(1024) loss_stats_kernel(const float* __restrict__ l, long long n, float* out4) {
    double s1 = 0.0, s2 = 0.0;
    float m = -CUDART_INF_F, s = 0.f;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = l[i];
        s1 += v; s2 += (double)v * v;
        lse_combine(m, s, -v, 1.f);
    }
    __shared__ double sh1[32], sh2[32];
    __shared__ float shm[32], shs[32];
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o), sx = __shfl_xor_sync(0xffffffffu, s, o);
        lse_combine(m, s, m2, sx);
    }
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sh1[w] = s1; sh2[w] = s2; shm[w] = m; shs[w] = s; }
    __syncthreads();
    if (w == 0) {
        const int nw = blockDim.x >> 5;
        s1 = lane < nw ? sh1[lane] : 0.0; s2 = lane < nw ? sh2[lane] : 0.0;
        m = lane < nw ? shm[lane] : -CUDART_INF_F; s = lane < nw ? shs[lane] : 0.f;
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            const float m2 = __shfl_xor_sync(0xffffffffu, m, o), sx = __shfl_xor_sync(0xffffffffu, s, o);
            lse_combine(m, s, m2, sx);
        }
        if (lane == 0) { out4[0] = (float)s1; out4[1] = (float)s2; out4[2] = m; out4[3] = s; }
    }
}

// grid = batches; losses[b][n] -> elbo[b] = -mean, lnz[b] = logsumexp(-l) - log n
__global__ void __launch_bounds__(256) batched_elbo_lnz_kernel(const float* __restrict__ losses, int n, float* elbo, float* lnz) {
    const float* l = losses + (size_t)blockIdx.x * n;
    float s1 = 0.f, m = -CUDART_INF_F, s = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = l[i];
        s1 += v;
        lse_combine(m, s, -v, 1.f);
    }
    __shared__ float sh1[8], shm[8], shs[8];
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o), sx = __shfl_xor_sync(0xffffffffu, s, o);
        lse_combine(m, s, m2, sx);
    }
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sh1[w] = s1; shm[w] = m; shs[w] = s; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) { s1 += sh1[i]; lse_combine(m, s, shm[i], shs[i]); }
        elbo[blockIdx.x] = -s1 / (float)n;
        lnz[blockIdx.x] = logf(s) + m - logf((float)n);
    }
}

__global__ void threefry_kernel(const uint32_t* key2, const uint32_t* x0, const uint32_t* x1, long long n, uint32_t* y0, uint32_t* y1) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    Key k; k.k0 = key2[0]; k.k1 = key2[1];
    uint32_t a = x0[i], b = x1[i];
    threefry2x32(k, a, b);
    y0[i] = a; y1[i] = b;
}

// every Gaussian a particle consumes (runtime d): xi0[n][d], xi[k][n][d]
__global__ void particle_noise_kernel(const int32_t* seeds, long long n, int d, int K, float* xi0, float* xi) {
    const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= n) return;
    Key k = prng_key(seeds[p]), ka;
    split(k, ka, k);
    for (int j = 0; j < d; ++j) xi0[p * d + j] = bits_to_normal(random_bits_at(ka, j, d));
    if (K < 1) return;
    ka = split_first(k);
    k = split_second(ka);
    for (int i = 0; i < K; ++i) {
        split(k, ka, k);
        for (int j = 0; j < d; ++j) xi[((size_t)i * n + p) * d + j] = bits_to_normal(random_bits_at(ka, j, d));
        k = split_second(k);
    }
}

int launch_loss_stats(cudaStream_t st, const float* negw, long long n, float* out4) {
    loss_stats_kernel<<<1, 1024, 0, st>>>(negw, n, out4);
    CMCD_CUDA_OK(cudaGetLastError());
    return 0;
}
int launch_batched_elbo_lnz(cudaStream_t st, const float* losses, int batches, int n, float* elbo, float* lnz) {
    batched_elbo_lnz_kernel<<<batches, 256, 0, st>>>(losses, n, elbo, lnz);
    CMCD_CUDA_OK(cudaGetLastError());
    return 0;
}
int launch_threefry(cudaStream_t st, const uint32_t* key2, const uint32_t* x0, const uint32_t* x1, long long n, uint32_t* y0, uint32_t* y1) {
    threefry_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(key2, x0, x1, n, y0, y1);
    CMCD_CUDA_OK(cudaGetLastError());
    return 0;
}
int launch_particle_noise(cudaStream_t st, const int32_t* seeds, long long n, int d, int K, float* xi0, float* xi) {
    particle_noise_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(seeds, n, d, K, xi0, xi);
    CMCD_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace cmcd

namespace cmcd {
// log p, score and (optionally) Hessian-vector product of a registry target at x[n][D].
template <int D>
__global__ void target_eval_kernel(TargetDesc t, const float* __restrict__ x, long long n, const float* __restrict__ v,
                                   float* lp, float* score, float* hvp) {
    __shared__ float sTp[MIX_MAX * MIX_STRIDE];
    const int ntp = (t.kind == TGT_GMM || t.kind == TGT_MANY_GMM) ? t.ncomp * MIX_STRIDE : 0;
    for (int i = threadIdx.x; i < ntp; i += blockDim.x) sTp[i] = t.mix[i];
    __syncthreads();
    const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= n) return;
    float z[D], g[D], vv[D], hv[D];
#pragma unroll
    for (int j = 0; j < D; ++j) { z[j] = x[p * D + j]; vv[j] = v ? v[p * D + j] : 0.f; hv[j] = 0.f; }
    const float l = v ? target_eval<D, true>(t, sTp, z, g, vv, hv) : target_eval<D, false>(t, sTp, z, g, vv, hv);
    if (lp) lp[p] = l;
#pragma unroll
    for (int j = 0; j < D; ++j) { if (score) score[p * D + j] = g[j]; if (hvp && v) hvp[p * D + j] = hv[j]; }
}

int launch_target_eval(cudaStream_t st, const TargetDesc& t, int D, const float* x, long long n, const float* v,
                       float* lp, float* score, float* hvp) {
    const unsigned grid = (unsigned)((n + 127) / 128);
    if (D == 2) target_eval_kernel<2><<<grid, 128, 0, st>>>(t, x, n, v, lp, score, hvp);
    else if (D == 10) target_eval_kernel<10><<<grid, 128, 0, st>>>(t, x, n, v, lp, score, hvp);
    else { set_error("target_eval: dim=%d has no small-d instantiation", D); return 2; }
    CMCD_CUDA_OK(cudaGetLastError());
    return 0;
}
}  // namespace cmcd

namespace cmcd {
// FP32 FMA-pipe peak probe (roofline denominator for the compute-bound bridge kernels; SURVEY 8d asks for a
// measured FFMA number because MEASURED_PEAKS.json only carries HBM and bf16-tensor peaks).
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* out, float a, float b, int iters) {
    float c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-9f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fmaf(c[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int launch_ffma_peak(cudaStream_t st, float* scratch, int blocks, int iters) {
    ffma_peak_kernel<<<blocks, 256, 0, st>>>(scratch, 1.0001f, 1e-7f, iters);
    CMCD_CUDA_OK(cudaGetLastError());
    return 0;
}
}  // namespace cmcd
