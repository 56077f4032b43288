// Wide path: d = 1600 (lgcp), few particles.  Each half-step is a dense [N x 1620] x [1620 x 1620] product whose
// operand (10.5 MB of fp32 weights, or the 10.2 MB dense K^-1 of the LGCP prior) is streamed from L2 by a split-K
// skinny GEMM spread over all SMs; small fused "finalize" kernels reduce the split-K partials and apply bias /
// softplus / residual / score / kernel-mean arithmetic.  One launch sequence per bridge step, enqueue-only.
//
// Replaces, for the lgcp target: vmap(compute_log_elbo) (src/mcdboundingmachine.py:126-205) over
// src/mcd_cais.py:46-89 / src/mcd_over_orig.py:18-55 with log_prob_model = LogGaussianCoxPines
// (src/model_handler.py:287-409, src/cp_utils.py:87-155) and apply_fun_sn = the geffner net (src/nn.py:42-72).
//
// LGCP density in K^-1 form: with dv = x - mu0,  -0.5|L^-1 dv|^2 = -0.5 dv^T K^-1 dv, so
//   log p(x) = -0.5 dv.(K^-1 dv) + log_norm + sum_j (x_j counts_j - a exp(x_j)),
//   score(x) = -K^-1 dv + counts - a exp(x),   H v = -K^-1 v - a exp(x) o v
// (the reference does two triangular solves per score; K is well conditioned, cond = 27.6).
#include <cstdlib>

#include "common.cuh"
#include "umma.cuh"

namespace cmcd {

constexpr int WG_THREADS = 128;   // skinny GEMM: 128 threads x 4 columns
constexpr int WG_COLS = WG_THREADS * 4;

// Y_part[s][n][m] = sum_{k in slice s} (X[n][k] - shift) * W[k][m]
// ROWS = particle rows per block: 8, or 24 so that the README lgcp batch (N = 20) is ONE row tile -- with 8-row tiles every
// weight element was fetched from L2 three times and fed 8 FMAs per load instead of 20 (27 us per product, launch list
// gpurun_out/launches_lgcp_warm.csv).  The k loop keeps 8 independent 16-byte weight loads in flight per thread.
template <int ROWS>
__global__ void __launch_bounds__(WG_THREADS) skinny_gemm_kernel(const float* __restrict__ X, int ldx, float shift,
                                                                 const float* __restrict__ W, int ldw, int N, int Kd,
                                                                 int M, int kslice, float* __restrict__ part) {
    extern __shared__ __align__(16) float sx[];  // [ROWS][kslice]
    const int m0 = blockIdx.x * WG_COLS + threadIdx.x * 4;
    const int s = blockIdx.y;
    const int n0 = blockIdx.z * ROWS;
    const int k0 = s * kslice, k1 = min(Kd, k0 + kslice);
    const int kl = k1 - k0;
    // stage the activations, four independent loads in flight per thread (the tail of the last slice is zero-filled: it is
    // read as float4 below)
    for (int base = threadIdx.x; base < ROWS * kslice; base += 4 * WG_THREADS) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = base + u * WG_THREADS;
            const int r = i / kslice, k = i % kslice;
            v[u] = (i < ROWS * kslice && n0 + r < N && k < kl) ? X[(size_t)(n0 + r) * ldx + k0 + k] - shift : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = base + u * WG_THREADS;
            if (i < ROWS * kslice) sx[i] = v[u];
        }
    }
    __syncthreads();
    float acc[ROWS][4];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) { acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f; }
    if (m0 < M) {
        const bool full = (m0 + 3 < M) && ((ldw & 3) == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);   // W may be a view into the flat parameter vector
        constexpr int KU = 8;
        for (int kb = 0; kb < kl; kb += KU) {   // kslice is a multiple of 8: kb + 7 < kslice
            float4 w[KU];
#pragma unroll
            for (int u = 0; u < KU; ++u) {
                const int k = kb + u;
                if (k < kl) {
                    const float* wp = W + (size_t)(k0 + k) * ldw + m0;
                    if (full) w[u] = __ldg(reinterpret_cast<const float4*>(wp));
                    else {
                        w[u].x = __ldg(wp);
                        w[u].y = (m0 + 1 < M) ? __ldg(wp + 1) : 0.f;
                        w[u].z = (m0 + 2 < M) ? __ldg(wp + 2) : 0.f;
                        w[u].w = (m0 + 3 < M) ? __ldg(wp + 3) : 0.f;
                    }
                } else {
                    w[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                const float4 xa = *reinterpret_cast<const float4*>(sx + r * kslice + kb);       // 2 LDS.128 per 32 FMAs
                const float4 xb = *reinterpret_cast<const float4*>(sx + r * kslice + kb + 4);
                const float xs[KU] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
                for (int u = 0; u < KU; ++u) {
                    acc[r][0] = fmaf(xs[u], w[u].x, acc[r][0]);
                    acc[r][1] = fmaf(xs[u], w[u].y, acc[r][1]);
                    acc[r][2] = fmaf(xs[u], w[u].z, acc[r][2]);
                    acc[r][3] = fmaf(xs[u], w[u].w, acc[r][3]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            if (n0 + r < N) {
                float* dst = part + ((size_t)s * N + n0 + r) * M + m0;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (m0 + q < M) dst[q] = acc[r][q];
            }
        }
    }
}

// part[0][e] <- sum_s part[s][e] (ascending s: the order every consumer used when it summed the slices itself).  The
// consumers are one-CTA-per-particle kernels (20 CTAs for the README lgcp batch) that spent most of their time on S dependent
// L2 reads per element; here the same reads are spread over the whole GPU and the consumers see a single slice.
__global__ void __launch_bounds__(256) wide_reduce_partials_kernel(float* __restrict__ part, int S, int nel) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nel) return;
    float v = 0.f;
    for (int s = 0; s < S; ++s) v += part[(size_t)s * nel + e];
    part[e] = v;
}

// ---- the same split-K product on tcgen05 tiles (the lgcp K^-1 "whitening" product and the 1620-wide layers, served from L2)
// One CTA = one 64-column tile of W x one K range; rows of X (particles) are the 128 TMEM lanes.  Per 64-deep K chunk: the W
// block [64 k][64 n] is staged once as tf32 (hi, lo) K-major core-matrix tiles, every thread puts its particle's 64 activations
// into its own TMEM lane as (hi, lo), one elected lane issues the 24 MMAs of the 3-pass split (a_hi b_lo + a_lo b_hi + a_hi b_hi,
// same scheme and accuracy as the bridge kernels, umma.cuh) accumulating in TMEM over the chunks; the next chunk's operands are
// requested from L2 while the tensor core works.  ~6 instructions per weight element instead of 24 FMAs + loads.
constexpr int WTC_KC = 64, WTC_NT = 64;
constexpr uint32_t WTC_A_HI = 0, WTC_A_LO = 64, WTC_D = 128, WTC_COLS = 256;
__global__ void __launch_bounds__(128) skinny_gemm_tc_kernel(const float* __restrict__ X, int ldx, float shift,
                                                             const float* __restrict__ W, int ldw, int N, int Kd, int M,
                                                             int chunks_per_slice, float* __restrict__ part) {
    __shared__ __align__(1024) uint8_t sB[2 * WTC_NT * WTC_KC * 4];   // B_hi | B_lo
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t mbar;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int m0 = blockIdx.x * WTC_NT;
    const int s = blockIdx.y;
    const int n0 = blockIdx.z * 128;
    const int row = n0 + tid;                   // this thread's particle (TMEM lane)
    const bool rowok = row < N;
    if (warp == 0) umma::tmem_alloc(&tmem_slot, WTC_COLS);
    if (tid == 0) umma::mbar_init(&mbar, 1);
    umma::fence_async_smem();
    umma::fence_before();
    __syncthreads();
    umma::fence_after();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const uint64_t dhi = umma::make_desc(umma::smem_u32(sB), 128, 32 * WTC_KC);
    const uint64_t dlo = umma::make_desc(umma::smem_u32(sB + WTC_NT * WTC_KC * 4), 128, 32 * WTC_KC);
    const uint32_t idesc = umma::make_idesc_tf32(128, WTC_NT);
    const int c0 = s * chunks_per_slice;
    const int nchunk = min(chunks_per_slice, (Kd + WTC_KC - 1) / WTC_KC - c0);
    const bool xvec = !(ldx & 3) && !(reinterpret_cast<uintptr_t>(X) & 15);

    // operand registers of one chunk: 32 W elements per thread -- lane l of warp w takes (n % 8, k % 4) = (l / 4, l % 4) of the
    // (n / 8, k / 4) block pair i * 4 + w, i.e. one 8 x 4 patch per warp instruction: its stores into the core-matrix tiles hit
    // 32 different banks (a row-major float4 mapping was an 8-way conflict) and its loads are four 32-byte sectors -- and the X
    // row as 64 floats
    float wreg[32];
    float xreg[WTC_KC];
    const int lane = tid & 31;
    auto request = [&](int c) {
        const int kb = (c0 + c) * WTC_KC;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int pair = i * 4 + warp;                 // 0..127: n-group = pair % 8, k-group = pair / 8
            const int n = (pair & 7) * 8 + (lane >> 2), k = kb + (pair >> 3) * 4 + (lane & 3);
            wreg[i] = (k < Kd && m0 + n < M) ? __ldg(W + (size_t)k * ldw + m0 + n) : 0.f;
        }
        if (rowok) {
            const float* xp = X + (size_t)row * ldx + kb;
            if (xvec && kb + WTC_KC <= Kd) {
#pragma unroll
                for (int q = 0; q < WTC_KC / 4; ++q) {
                    const float4 v = *reinterpret_cast<const float4*>(xp + 4 * q);
                    xreg[4 * q] = v.x; xreg[4 * q + 1] = v.y; xreg[4 * q + 2] = v.z; xreg[4 * q + 3] = v.w;
                }
            } else {
#pragma unroll
                for (int q = 0; q < WTC_KC; ++q) xreg[q] = (kb + q < Kd) ? xp[q] : shift;
            }
        } else {
#pragma unroll
            for (int q = 0; q < WTC_KC; ++q) xreg[q] = shift;
        }
    };
    uint32_t parity = 0;
    if (nchunk > 0) request(0);
    for (int c = 0; c < nchunk; ++c) {
        if (c > 0) { umma::mbar_wait(&mbar, parity); parity ^= 1u; umma::fence_after(); }   // tensor core done with B tiles and A lanes
        {   // W block -> tf32 hi / lo core-matrix tiles: element (n, k) = W[kb + k][m0 + n]
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int pair = i * 4 + warp;
                const int n = (pair & 7) * 8 + (lane >> 2), k = (pair >> 3) * 4 + (lane & 3);
                float hi, lo;
                umma::split_tf32(wreg[i], hi, lo);
                const int off = umma::core_off(n, k, WTC_KC);
                *reinterpret_cast<float*>(sB + off) = hi;
                *reinterpret_cast<float*>(sB + WTC_NT * WTC_KC * 4 + off) = lo;
            }
        }
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {   // X row -> this thread's TMEM lane
            uint32_t hh[16], ll[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                float hi, lo;
                umma::split_tf32(xreg[cc * 16 + e] - shift, hi, lo);
                hh[e] = __float_as_uint(hi); ll[e] = __float_as_uint(lo);
            }
            umma::tmem_st16(tmem_lane + WTC_A_HI + cc * 16, hh);
            umma::tmem_st16(tmem_lane + WTC_A_LO + cc * 16, ll);
        }
        umma::tmem_st_wait();
        umma::fence_before();
        umma::fence_async_smem();
        __syncthreads();
        if (warp == 0) {
            umma::fence_after();
            if (umma::elect_one()) {
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t acol = tmem_base + (pass == 1 ? WTC_A_LO : WTC_A_HI);
                    const uint64_t bd = (pass == 0) ? dlo : dhi;
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        umma::mma_tf32_ts(tmem_base + WTC_D, acol + k * 8, bd + (uint64_t)(k * 16), idesc, (c | pass | k) > 0);
                }
                umma::commit(&mbar);
            }
            __syncwarp();
        }
        if (c + 1 < nchunk) request(c + 1);   // L2 latency of the next chunk overlaps the MMAs
    }
    if (nchunk > 0) { umma::mbar_wait(&mbar, parity); umma::fence_after(); }
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
        uint32_t v[16];
        if (nchunk > 0) { umma::tmem_ld16(tmem_lane + WTC_D + cc * 16, v); umma::tmem_ld_wait(); }
        if (rowok) {
            float* dst = part + ((size_t)s * N + row) * M + m0 + cc * 16;
#pragma unroll
            for (int e = 0; e < 16; ++e)
                if (m0 + cc * 16 + e < M) dst[e] = nchunk > 0 ? __uint_as_float(v[e]) : 0.f;
        }
    }
    umma::fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem_slot, WTC_COLS);
}

struct WideGemm {
    int S, kslice, rows;
};
static WideGemm plan_gemm(int N, int Kd, int M, int num_sms) {
    const int rows = (N > 8 && N <= 24) ? 24 : 8;
    const int cg = (M + WG_COLS - 1) / WG_COLS, rt = (N + rows - 1) / rows;
    int S = (2 * num_sms + cg * rt - 1) / (cg * rt);
    if (S < 1) S = 1;
    const int smax = rows == 24 ? 32 : 64;   // the finalize kernels read S partials per element
    if (S > smax) S = smax;
    int ks = (Kd + S - 1) / S;
    ks = (ks + 7) & ~7;
    S = (Kd + ks - 1) / ks;
    return {S, ks, rows};
}
// reduce = false: the consumer is an element-wise kernel spread over the whole GPU (layer finalizers) and sums the S slices
// itself -- one launch less per product; one-CTA-per-particle consumers get the slices pre-summed (reduce = true).
static int run_gemm(cudaStream_t st, const float* X, int ldx, float shift, const float* W, int ldw, int N, int Kd, int M,
                    int num_sms, float* part, int* S_out, bool reduce = true) {
    // large products: tcgen05 tiles (CMCD_DISABLE_WIDE_TC=1 keeps the FP32-FMA kernel for A/B runs)
    if ((long long)Kd * M >= 256LL * 256 && !std::getenv("CMCD_DISABLE_WIDE_TC")) {
        const int nchunks = (Kd + WTC_KC - 1) / WTC_KC;
        const int ctiles = (M + WTC_NT - 1) / WTC_NT, rtiles = (N + 127) / 128;
        int S = (2 * num_sms + ctiles * rtiles - 1) / (ctiles * rtiles);   // ~2 CTAs per SM = what TMEM lets be resident; measured per
                                                                           // product at 1 / 2 / 3 / 5 per SM: 13.9 / 11.6 / 15.0 / 15.5 us
        if (const char* e = std::getenv("CMCD_WIDE_TC_CTAS_PER_SM")) S = (atoi(e) * num_sms + ctiles * rtiles - 1) / (ctiles * rtiles);
        if (S < 1) S = 1;
        if (S > nchunks) S = nchunks;
        const int cps = (nchunks + S - 1) / S;
        S = (nchunks + cps - 1) / cps;
        skinny_gemm_tc_kernel<<<dim3(ctiles, S, rtiles), 128, 0, st>>>(X, ldx, shift, W, ldw, N, Kd, M, cps, part);
        CMCD_CUDA_OK(cudaGetLastError());
        if (S > 1 && reduce) {
            const int nel = N * M;
            wide_reduce_partials_kernel<<<(nel + 255) / 256, 256, 0, st>>>(part, S, nel);
            CMCD_CUDA_OK(cudaGetLastError());
        }
        *S_out = reduce ? 1 : S;
        return 0;
    }
    const WideGemm g = plan_gemm(N, Kd, M, num_sms);
    dim3 grid((M + WG_COLS - 1) / WG_COLS, g.S, (N + g.rows - 1) / g.rows);
    const size_t smem = (size_t)g.rows * g.kslice * sizeof(float);
    if (g.rows == 24) skinny_gemm_kernel<24><<<grid, WG_THREADS, smem, st>>>(X, ldx, shift, W, ldw, N, Kd, M, g.kslice, part);
    else skinny_gemm_kernel<8><<<grid, WG_THREADS, smem, st>>>(X, ldx, shift, W, ldw, N, Kd, M, g.kslice, part);
    CMCD_CUDA_OK(cudaGetLastError());
    if (g.S > 1 && reduce) {
        const int nel = N * M;
        wide_reduce_partials_kernel<<<(nel + 255) / 256, 256, 0, st>>>(part, g.S, nel);
        CMCD_CUDA_OK(cudaGetLastError());
    }
    *S_out = reduce ? 1 : g.S;
    return 0;
}

// One-block-per-particle kernels (N ~ 20 blocks on 148 SMs): their run time is one thread's serial walk over d / blockDim elements with
// ~12 dependent-latency loads each, so the block is as wide as the register budget allows (d = 1600: 6.25 -> 1.6 elements per thread;
// node kernel 33 -> 14 us, kernel means 19 -> 9-10 us, target finalize 10.5 -> 5.5 us).
constexpr int WIDE_PB_MAX = 1024;
static int wide_pb(int d) { return d >= 1024 ? 1024 : d >= 512 ? 512 : 256; }

__device__ __forceinline__ float block_sum(float v, float* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    float t = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.f;
    if (w == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (l == 0) sh[0] = t;
    }
    __syncthreads();
    return sh[0];
}

__device__ __forceinline__ float softplus_f(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }

// ---- init: keys, z0 = mu + sigma xi0, w0 = -log q(z0).  One block per particle. ---------------------------
__global__ void __launch_bounds__(WIDE_PB_MAX) wide_init_kernel(const int32_t* seeds, int d, const float* mu, const float* logdiag,
                                                        float* z, float* w, uint32_t* keys, float* traj, long long N) {
    __shared__ float sh[32];
    const int n = blockIdx.x;
    Key k = prng_key(seeds[n]), ka;
    split(k, ka, k);
    float lq = 0.f;
    for (int j = threadIdx.x; j < d; j += blockDim.x) {
        const float xi = bits_to_normal(random_bits_at(ka, j, d));
        const float sg = expf(logdiag[j]);
        const float zz = sg * xi + mu[j];
        z[(size_t)n * d + j] = zz;
        if (traj) traj[(size_t)j * N + n] = zz;
        const float v = (zz - mu[j]) / sg;
        lq += -0.5f * v * v - logf(2.5066282746310002f * sg);
    }
    lq = block_sum(lq, sh);
    if (threadIdx.x == 0) {
        w[n] = -lq;
        ka = split_first(k);    // mcdboundingmachine.py:162
        k = split_second(ka);   // mcd_cais.py:94
        keys[2 * n] = k.k0; keys[2 * n + 1] = k.k1;
    }
}

// ---- target finalize: sp = -(K^-1 dv) + counts - a exp(x); lp = -0.5 dv.(K^-1 dv) + log_norm + sum(x c - a e^x) ----
__global__ void __launch_bounds__(WIDE_PB_MAX) wide_target_fin_kernel(const float* __restrict__ part, int S, int N, int d,
                                                              const float* __restrict__ x, const float* __restrict__ counts,
                                                              float mu0, float log_norm, float area, float* sp, float* lp) {
    __shared__ float sh[32];
    const int n = blockIdx.x;
    float acc = 0.f;
    for (int j = threadIdx.x; j < d; j += blockDim.x) {
        float kd = 0.f;
        for (int s = 0; s < S; ++s) kd += part[((size_t)s * N + n) * d + j];
        const float xx = x[(size_t)n * d + j], ex = expf(xx);
        sp[(size_t)n * d + j] = -kd + counts[j] - area * ex;
        acc += -0.5f * (xx - mu0) * kd + (xx * counts[j] - area * ex);
    }
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) lp[n] = acc + log_norm;
}

// ---- network finalize kernels (geffner table form with x folded into the first d hidden units) ----
// stage 1: a1 = softplus(sum part + c1[t]); A1 = a1 + pad(x)
__device__ __forceinline__ float sigmoid_f(float x) {
    const float e = expf(-fabsf(x));
    const float s = 1.0f / (1.0f + e);
    return x >= 0.f ? s : 1.0f - s;
}
// (a1 / a2 outputs: if non-null they receive softplus'(pre) = sigmoid(pre), which is what the reverse pass needs)
__global__ void wide_l1_fin_kernel(const float* __restrict__ part, int S, int N, int HP, int d, const float* __restrict__ c1t,
                                   const float* __restrict__ x, int fold_x, float* a1, float* A1) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * HP) return;
    const int n = i / HP, j = i % HP;
    float p = c1t[j];
    for (int s = 0; s < S; ++s) p += part[((size_t)s * N + n) * HP + j];
    const float a = softplus_f(p);
    if (a1) a1[i] = sigmoid_f(p);
    A1[i] = a + ((fold_x && j < d) ? x[(size_t)n * d + j] : 0.f);
}
// stage 2: a2 = softplus(sum part + c2[t]); A2 = a2 + skip*A1   (A1 already holds a1 + pad(x))
__global__ void wide_l2_fin_kernel(const float* __restrict__ part, int S, int N, int HP, const float* __restrict__ c2t,
                                   const float* __restrict__ A1, float skip, float* a2, float* A2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * HP) return;
    const int n = i / HP, j = i % HP;
    float p = c2t[j];
    for (int s = 0; s < S; ++s) p += part[((size_t)s * N + n) * HP + j];
    const float a = softplus_f(p);
    if (a2) a2[i] = sigmoid_f(p);
    A2[i] = a + skip * A1[i];
}

// ---- forward-kernel mean + sample:  mf = z - eps uf - eps NN ; zn = mf + s xi ; fkterm kept per element ----
// One block per particle.  NN output = out_scale * clamp(sum part3 + c3[t]).  Advances the key chain.
__global__ void __launch_bounds__(WIDE_PB_MAX) wide_fwd_mean_kernel(const float* __restrict__ part3, int S, int N, int d,
                                                            const float* __restrict__ c3t, float out_scale_v, const float* __restrict__ out_scale_dev, float out_clip,
                                                            int use_nn, const float* __restrict__ z, const float* __restrict__ sp,
                                                            const float* __restrict__ mu, const float* __restrict__ logdiag,
                                                            const float* __restrict__ betas, const float* __restrict__ epss, int step,
                                                            float clip_t, float clip_q,
                                                            uint32_t* keys, float* zn, float* mf_out, float* traj_row) {
    const int n = blockIdx.x;
    const float out_scale = out_scale_dev ? __ldg(out_scale_dev) : out_scale_v;
    const float beta = betas[step], eps = epss[step];
    Key k; k.k0 = keys[2 * n]; k.k1 = keys[2 * n + 1];
    Key ka, kn;
    split(k, ka, kn);                     // mcd_cais.py:66
    const float scale = sqrtf(2.0f * eps);
    for (int j = threadIdx.x; j < d; j += blockDim.x) {
        const size_t e = (size_t)n * d + j;
        float nn = 0.f;
        if (use_nn) {
            float o = c3t[j];
            for (int s = 0; s < S; ++s) o += part3[((size_t)s * N + n) * d + j];
            nn = out_scale * fminf(fmaxf(o, -out_clip), out_clip);
        }
        const float sg = expf(logdiag[j]);
        const float sq = -((z[e] - mu[j]) / sg) / sg;
        const float gu = fminf(fmaxf(sp[e], -clip_t), clip_t), gq = fminf(fmaxf(sq, -clip_q), clip_q);
        const float uf = -(beta * gu + (1.0f - beta) * gq);
        const float mf = (z[e] - eps * uf) - eps * nn;
        const float xi = bits_to_normal(random_bits_at(ka, j, d));
        const float znew = mf + scale * xi;
        zn[e] = znew;
        mf_out[e] = mf;
        if (traj_row) traj_row[(size_t)j * N + n] = znew;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const Key k2 = split_second(kn);  // mcd_cais.py:87
        keys[2 * n] = k2.k0; keys[2 * n + 1] = k2.k1;
    }
}

// ---- backward-kernel mean + weight update:  mb = zn - eps ub + eps NN ; w += logN(z; mb, s) - logN(zn; mf, s) ----
__global__ void __launch_bounds__(WIDE_PB_MAX) wide_bwd_mean_kernel(const float* __restrict__ part3, int S, int N, int d,
                                                            const float* __restrict__ c3t, float out_scale_v, const float* __restrict__ out_scale_dev, float out_clip,
                                                            int use_nn, const float* __restrict__ z, const float* __restrict__ zn,
                                                            const float* __restrict__ mf, const float* __restrict__ spn,
                                                            const float* __restrict__ mu, const float* __restrict__ logdiag,
                                                            const float* __restrict__ betas, const float* __restrict__ epss, int step,
                                                            float clip_t, float clip_q, float* w) {
    __shared__ float sh[32];
    const int n = blockIdx.x;
    const float out_scale = out_scale_dev ? __ldg(out_scale_dev) : out_scale_v;
    const float beta = betas[step], eps = epss[step];
    const float scale = sqrtf(2.0f * eps);
    float acc = 0.f;
    for (int j = threadIdx.x; j < d; j += blockDim.x) {
        const size_t e = (size_t)n * d + j;
        float nn = 0.f;
        if (use_nn) {
            float o = c3t[j];
            for (int s = 0; s < S; ++s) o += part3[((size_t)s * N + n) * d + j];
            nn = out_scale * fminf(fmaxf(o, -out_clip), out_clip);
        }
        const float sg = expf(logdiag[j]);
        const float sq = -((zn[e] - mu[j]) / sg) / sg;
        const float gu = fminf(fmaxf(spn[e], -clip_t), clip_t), gq = fminf(fmaxf(sq, -clip_q), clip_q);
        const float ub = -(beta * gu + (1.0f - beta) * gq);
        const float mb = (zn[e] - eps * ub) + eps * nn;
        const float vb = (z[e] - mb) / scale, vf = (zn[e] - mf[e]) / scale;
        // the two log-normalisers -log(sqrt(2 pi) s) cancel exactly per element
        acc += -0.5f * vb * vb + 0.5f * vf * vf;
    }
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) w[n] += acc;
}

__global__ void wide_final_kernel(const float* w, const float* lp, int N, float* out_negw) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < N) out_negw[n] = -(w[n] + lp[n]);
}

// Two independent chains meet at every trajectory point: the target score (dense K^-1 product + finalize) and the drift network
// (three products + finalizers) -- and, in the reverse pass, the Hessian-vector product and the network pull-back.  Each of their
// kernels is a short split-K product that cannot fill the GPU by itself for long (launch ramp, 3 chunks per CTA, tail), so the
// two chains run on two streams and join where the step algebra needs both (fork / join with events: legal under stream
// capture, so opt.run(graph=True) keeps working).  CMCD_WIDE_SERIAL=1 keeps everything on the caller's stream (A/B runs).
struct WideAsync {
    cudaStream_t side = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
static WideAsync* wide_async() {
    constexpr int MAX_DEV = 64;
    static WideAsync tab[MAX_DEV];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV || std::getenv("CMCD_WIDE_SERIAL")) return nullptr;
    WideAsync& w = tab[dev];
    if (!w.side) {
        if (cudaStreamCreateWithFlags(&w.side, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&w.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&w.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    return &w;
}

// workspace layout (floats)
struct WideWs {
    size_t z, zn, mf, sp, lp, w, keys, A1, A2, part, partT, total;
};
static WideWs wide_layout(long long N, int d, int HP) {
    WideWs L;
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 3) & ~(size_t)3; return r; };
    L.z = take(N * d); L.zn = take(N * d); L.mf = take(N * d); L.sp = take(N * d);
    L.lp = take(N); L.w = take(N); L.keys = take(2 * N);
    L.A1 = take(N * (size_t)HP); L.A2 = take(N * (size_t)HP);
    const int mx = HP > d ? HP : d;
    L.part = take((size_t)64 * N * mx);
    L.partT = take((size_t)64 * N * d);     // split-K slices of the target's K^-1 product (runs concurrently with the network's)
    L.total = o;
    return L;
}

size_t wide_fwd_workspace_bytes(long long N, int d, int HP) { return wide_layout(N, d, HP > 0 ? HP : 8).total * sizeof(float); }

int launch_wide_fwd(const BridgeArgs& a, const cmcd_target* tg, int D, cudaStream_t st, int num_sms, void* ws, size_t ws_bytes) {
    const long long N = a.N;
    const int d = D, K = a.K;
    const NetView& nv = a.net;
    const bool has_net = nv.arch != CMCD_ARCH_NONE;
    const int HP = has_net ? nv.HP : 8;
    if (has_net && nv.arch != CMCD_ARCH_GEFFNER) { set_error("wide path (lgcp / callback targets): only nn_arch=geffner is implemented (README.md:63 config)"); return 2; }
    const WideWs L = wide_layout(N, d, HP);
    if (!ws || ws_bytes < L.total * sizeof(float)) { set_error("wide_fwd: workspace too small (%zu < %zu)", ws_bytes, L.total * sizeof(float)); return 2; }
    float* f = (float*)ws;
    float *z = f + L.z, *zn = f + L.zn, *mf = f + L.mf, *sp = f + L.sp, *lp = f + L.lp, *w = f + L.w;
    float *A1 = f + L.A1, *A2 = f + L.A2, *part = f + L.part, *partT = f + L.partT;
    uint32_t* keys = (uint32_t*)(f + L.keys);
    WideAsync* as = (tg->kind == CMCD_TARGET_CALLBACK) ? nullptr : wide_async();   // callbacks run on the caller's stream
    cudaStream_t st_t = as ? as->side : st;                                          // stream of the target chain
    int ST = 1;
    const bool cais = (a.mode == CMCD_MODE_CAIS_SN || a.mode == CMCD_MODE_CAIS_VAR_SN);
    const bool nn_b = (a.mode != CMCD_MODE_ULA) && has_net, nn_f = cais && has_net;
    const float skip = 1.f;
    int S = 1;

    wide_init_kernel<<<(unsigned)N, wide_pb(d), 0, st>>>(a.seeds, d, a.vd_mean, a.vd_logdiag, z, w, keys, a.traj, N);
    CMCD_CUDA_OK(cudaGetLastError());
    auto target_at = [&](const float* x) -> int {
        if (tg->kind == CMCD_TARGET_CALLBACK) {   // generic target: the caller's batched score, enqueued on the same stream
            if (int rc = tg->eval(tg->user, (void*)st, x, (int64_t)N, d, nullptr, lp, sp, nullptr)) { set_error("target callback failed (rc=%d)", rc); return 3; }
            return 0;
        }
        if (as) { CMCD_CUDA_OK(cudaEventRecord(as->fork, st)); CMCD_CUDA_OK(cudaStreamWaitEvent(as->side, as->fork, 0)); }
        if (int rc = run_gemm(st_t, x, d, tg->lgcp_mu0, tg->lgcp_kinv, d, (int)N, d, d, num_sms, partT, &ST)) return rc;
        wide_target_fin_kernel<<<(unsigned)N, wide_pb(d), 0, st_t>>>(partT, ST, (int)N, d, x, tg->lgcp_counts, tg->lgcp_mu0,
                                                              tg->lgcp_log_norm, tg->lgcp_bin_area, sp, lp);
        CMCD_CUDA_OK(cudaGetLastError());
        if (as) CMCD_CUDA_OK(cudaEventRecord(as->join, as->side));
        return 0;
    };
    auto target_join = [&]() -> int {             // the caller's stream waits for the target chain
        if (as) CMCD_CUDA_OK(cudaStreamWaitEvent(st, as->join, 0));
        return 0;
    };
    // NN(x, t) up to the layer-3 split-K partials (left in `part`, S3 slices)
    auto net_at = [&](const float* x, int t, int* S3) -> int {
        const int nel = (int)N * HP, blk = (nel + 255) / 256;
        if (int rc = run_gemm(st, x, d, 0.f, nv.U1, HP, (int)N, d, HP, num_sms, part, &S, false)) return rc;
        wide_l1_fin_kernel<<<blk, 256, 0, st>>>(part, S, (int)N, HP, d, nv.c1 + (size_t)t * HP, x, 1, nullptr, A1);
        CMCD_CUDA_OK(cudaGetLastError());
        if (int rc = run_gemm(st, A1, HP, 0.f, nv.W2, HP, (int)N, HP, HP, num_sms, part, &S, false)) return rc;
        wide_l2_fin_kernel<<<blk, 256, 0, st>>>(part, S, (int)N, HP, nv.c2 + (size_t)t * HP, A1, skip, nullptr, A2);
        CMCD_CUDA_OK(cudaGetLastError());
        if (int rc = run_gemm(st, A2, HP, 0.f, nv.W3, d, (int)N, HP, d, num_sms, part, S3)) return rc;
        return 0;
    };
    if (int rc = target_at(z)) return rc;         // (side stream) ...
    // One network evaluation per trajectory point: in the CAIS modes NN(z', i + 1) of step i's backward-kernel mean
    // (mcd_cais.py:78) is the evaluation step i + 1's forward-kernel mean needs (mcd_cais.py:60); its layer-3 partials stay
    // in `part` between wide_bwd_mean_kernel and the next wide_fwd_mean_kernel, so K + 1 evaluations serve 2K uses.
    int S3 = 1;
    if (nn_f && K > 0) { if (int rc = net_at(z, 0, &S3)) return rc; }   // ... while the network runs on the caller's stream
    if (int rc = target_join()) return rc;
    for (int i = 0; i < K; ++i) {
        wide_fwd_mean_kernel<<<(unsigned)N, wide_pb(d), 0, st>>>(part, S3, (int)N, d, has_net ? nv.c3 + (size_t)i * d : nullptr,
                                                          nv.out_scale, nv.out_scale_dev, nv.out_clip, nn_f ? 1 : 0, z, sp, a.vd_mean, a.vd_logdiag,
                                                          a.betas, a.eps, i, a.clip_t, a.clip_q, keys, zn, mf,
                                                          a.traj ? a.traj + (size_t)(i + 1) * d * N : nullptr);
        CMCD_CUDA_OK(cudaGetLastError());
        if (int rc = target_at(zn)) return rc;   // sp <- score at z' (reused as the next step's forward score)
        const int tb = cais ? i + 1 : i;
        if (nn_b) { if (int rc = net_at(zn, tb, &S3)) return rc; }
        if (int rc = target_join()) return rc;
        wide_bwd_mean_kernel<<<(unsigned)N, wide_pb(d), 0, st>>>(part, S3, (int)N, d, has_net ? nv.c3 + (size_t)tb * d : nullptr,
                                                          nv.out_scale, nv.out_scale_dev, nv.out_clip, nn_b ? 1 : 0, z, zn, mf, sp, a.vd_mean,
                                                          a.vd_logdiag, a.betas, a.eps, i, a.clip_t, a.clip_q, w);
        CMCD_CUDA_OK(cudaGetLastError());
        float* t = z; z = zn; zn = t;
    }
    wide_final_kernel<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(w, lp, (int)N, a.out_negw);
    CMCD_CUDA_OK(cudaGetLastError());
    CMCD_CUDA_OK(cudaMemcpyAsync(a.out_z, z, (size_t)N * d * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
}


// =====================================================================================================================
// Reverse pass of the wide path (lgcp): recompute from the stored z_k trajectory, same per-step algebra as
// bridge_bwd.cu (documented there), every R^d / R^HP vector operation as a [N x .] kernel and every product with a
// 1620-wide weight matrix or the dense K^-1 as a skinny GEMM served from L2.
// Replaces jax.grad of compute_bound (src/main.py:174-176) for log_prob_model = LogGaussianCoxPines
// (src/model_handler.py:287-409): H v = -K^-1 v - a exp(x) o v.
// Table cotangents: the wide path folds x into the first d hidden units (A1 = a1 + pad(x)), so the x-rows of W2 / W3
// receive the U2 / U3 cotangents too (U2 = W2[:d], U3 = W3[:d] for the geffner net, nn.py:45-52); g_U2 = g_U3 = 0.

// Y[n][j] = sum_m X[n][m] W[j][m] (+ addend[n][j]);  Y2[n][j] = Y[n][j] * mult[n][j].  One warp per row j of W.
constexpr int WT_NT = 20;   // particles per register tile (the README lgcp batch in one pass over the weight row)
__global__ void __launch_bounds__(256) wide_gemm_t_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W, int ldw,
                                                          int N, int J, int M, const float* __restrict__ addend, int lda,
                                                          const float* __restrict__ mult, int ldm, float* __restrict__ Y,
                                                          float* __restrict__ Y2, int ldy) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const bool vec = !((M | ldx | ldw) & 3) && !((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(W)) & 15);
    for (int j = blockIdx.x * wpb + (threadIdx.x >> 5); j < J; j += gridDim.x * wpb) {
        const float* __restrict__ wr = W + (size_t)j * ldw;
        for (int n0 = 0; n0 < N; n0 += WT_NT) {
            float acc[WT_NT];
#pragma unroll
            for (int r = 0; r < WT_NT; ++r) acc[r] = 0.f;
            if (vec) {
                for (int m = lane * 4; m < M; m += 128) {
                    const float4 w = __ldg(reinterpret_cast<const float4*>(wr + m));
#pragma unroll
                    for (int r = 0; r < WT_NT; ++r) {
                        if (n0 + r < N) {
                            const float4 x = *reinterpret_cast<const float4*>(X + (size_t)(n0 + r) * ldx + m);
                            acc[r] = fmaf(x.x, w.x, fmaf(x.y, w.y, fmaf(x.z, w.z, fmaf(x.w, w.w, acc[r]))));
                        }
                    }
                }
            } else {   // rows not 16-byte aligned (generic targets with dim % 4 != 0)
                for (int m = lane; m < M; m += 32) {
                    const float w = __ldg(wr + m);
#pragma unroll
                    for (int r = 0; r < WT_NT; ++r)
                        if (n0 + r < N) acc[r] = fmaf(X[(size_t)(n0 + r) * ldx + m], w, acc[r]);
                }
            }
#pragma unroll
            for (int r = 0; r < WT_NT; ++r) {
                float v = acc[r];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0 && n0 + r < N) {
                    const int n = n0 + r;
                    if (addend) v += addend[(size_t)n * lda + j];
                    Y[(size_t)n * ldy + j] = v;
                    if (Y2) Y2[(size_t)n * ldy + j] = v * mult[(size_t)n * ldm + j];
                }
            }
        }
    }
}

// Same product with X staged in shared memory once per CTA (persistent CTAs, one weight row per warp): the global-load
// version above waits on 20 dependent L1/L2 loads of X per weight chunk (ncu: 63 % of its samples on that scoreboard, 18 %
// issue utilisation, 36-55 us per call at N = 20, M = 1620).  Needs N * M floats of shared memory and 16-byte aligned rows.
__global__ void __launch_bounds__(512) wide_gemm_t_smem_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W, int ldw,
                                                               int N, int J, int M, const float* __restrict__ addend, int lda,
                                                               const float* __restrict__ mult, int ldm, float* __restrict__ Y,
                                                               float* __restrict__ Y2, int ldy) {
    extern __shared__ float4 sx4[];
    float* sXs = reinterpret_cast<float*>(sx4);   // [N][M]
    const int M4 = M >> 2;
    for (int base = threadIdx.x; base < N * M4; base += 4 * blockDim.x) {   // four independent 16-byte loads in flight per thread
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = base + u * blockDim.x;
            if (i < N * M4) v[u] = *reinterpret_cast<const float4*>(X + (size_t)(i / M4) * ldx + 4 * (i % M4));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = base + u * blockDim.x;
            if (i < N * M4) sx4[i] = v[u];
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int j = blockIdx.x * wpb + (threadIdx.x >> 5); j < J; j += gridDim.x * wpb) {
        const float* __restrict__ wr = W + (size_t)j * ldw;
        for (int n0 = 0; n0 < N; n0 += WT_NT) {
            float acc[WT_NT];
#pragma unroll
            for (int r = 0; r < WT_NT; ++r) acc[r] = 0.f;
            for (int mb = lane * 4; mb < M; mb += 4 * 128) {   // four weight chunks requested before the first is used
                float4 w[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int m = mb + u * 128;
                    w[u] = (m < M) ? __ldg(reinterpret_cast<const float4*>(wr + m)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int m = mb + u * 128;
                    if (m < M) {
#pragma unroll
                        for (int r = 0; r < WT_NT; ++r) {
                            if (n0 + r < N) {
                                const float4 x = *reinterpret_cast<const float4*>(sXs + (size_t)(n0 + r) * M + m);
                                acc[r] = fmaf(x.x, w[u].x, fmaf(x.y, w[u].y, fmaf(x.z, w[u].z, fmaf(x.w, w[u].w, acc[r]))));
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < WT_NT; ++r) {
                float v = acc[r];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0 && n0 + r < N) {
                    const int n = n0 + r;
                    if (addend) v += addend[(size_t)n * lda + j];
                    Y[(size_t)n * ldy + j] = v;
                    if (Y2) Y2[(size_t)n * ldy + j] = v * mult[(size_t)n * ldm + j];
                }
            }
        }
    }
}

// picks the staged version when X fits shared memory and the rows are 16-byte aligned (same summation order either way)
static int run_gemm_t(cudaStream_t st, int num_sms, const float* X, int ldx, const float* W, int ldw, int N, int J, int M,
                      const float* addend, int lda, const float* mult, int ldm, float* Y, float* Y2, int ldy) {
    const size_t smem = (size_t)N * M * sizeof(float);
    const bool vec = !((M | ldx | ldw) & 3) && !((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(W)) & 15);
    if (vec && smem <= 200 * 1024) {
        CMCD_CUDA_OK(cudaFuncSetAttribute(wide_gemm_t_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        wide_gemm_t_smem_kernel<<<num_sms, 512, smem, st>>>(X, ldx, W, ldw, N, J, M, addend, lda, mult, ldm, Y, Y2, ldy);
    } else {
        wide_gemm_t_kernel<<<2 * num_sms, 256, 0, st>>>(X, ldx, W, ldw, N, J, M, addend, lda, mult, ldm, Y, Y2, ldy);
    }
    CMCD_CUDA_OK(cudaGetLastError());
    return 0;
}

// G[i][j] = sum_r A[r][i] B[r][j] over R stacked rows (all particles of all trajectory points): the weight cotangents of one
// whole reverse pass as ONE product per layer instead of a rank-N read-modify-write of the 10 MB gradient per point.
// 64 x 64 output tile per CTA, 4 x 4 per thread, 16-row slabs of A and B staged in shared memory.  Overwrites G.
constexpr int WGR_T = 64, WGR_K = 16;
__global__ void __launch_bounds__(256) wide_wgrad_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, int R,
                                                         int I, int J, float* __restrict__ G, int ldg) {
    __shared__ float As[WGR_K][WGR_T + 4], Bs[WGR_K][WGR_T + 4];
    const int i0 = blockIdx.y * WGR_T, j0 = blockIdx.x * WGR_T, tid = threadIdx.x;
    const int ti = (tid / 16) * 4, tj = (tid % 16) * 4;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    for (int r0 = 0; r0 < R; r0 += WGR_K) {
        for (int e = tid; e < WGR_K * WGR_T; e += 256) {
            const int r = e / WGR_T, c = e % WGR_T;
            const bool rv = r0 + r < R;
            As[r][c] = (rv && i0 + c < I) ? A[(size_t)(r0 + r) * lda + i0 + c] : 0.f;
            Bs[r][c] = (rv && j0 + c < J) ? B[(size_t)(r0 + r) * ldb + j0 + c] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < WGR_K; ++r) {
            const float4 av = *reinterpret_cast<const float4*>(&As[r][ti]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[r][tj]);
            const float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(a4[a], b4[b], acc[a][b]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
            if (i0 + ti + a < I && j0 + tj + b < J) G[(size_t)(i0 + ti + a) * ldg + j0 + tj + b] = acc[a][b];
}
// g[t][j] = sum_n V[t][n][j] for the stacked per-point buffers (per-step bias-table cotangents); grid (J / 256, points)
__global__ void wide_colsum_rows_kernel(const float* __restrict__ V, int ldv, int N, int J, float* __restrict__ g, int ldg) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, t = blockIdx.y;
    if (j >= J) return;
    float s = 0.f;
    for (int n = 0; n < N; ++n) s += V[((size_t)t * N + n) * ldv + j];
    g[(size_t)t * ldg + j] = s;
}
__global__ void wide_neg_kernel(const float* __restrict__ cot, int N, float* __restrict__ c) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < N) c[n] = -cot[n];
}

// out[n][j] = traj[j][n]  (trajectory rows are stored particle-fastest)
__global__ void wide_gather_kernel(const float* __restrict__ traj_row, int N, int d, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * d) return;
    const int n = i / d, j = i % d;
    out[i] = traj_row[(size_t)j * N + n];
}

// One node of the reverse pass (x = z_j): network output from the layer-3 partials, BOTH kernel means that use this
// point -- backward-kernel mean of step j-1 (B use, needs z_{j-1}) and forward-kernel mean of step j (F use, needs
// z_{j+1}) -- their residuals and cotangents G_B, G_F, the combined output-layer cotangent vo = out_scale (eps_B G_B -
// eps_F G_F) (NN(z_j, j) is one evaluation used twice by the reference, mcd_cais.py:60,78, and its VJP is linear), the
// combined HVP input vm = mk_t (beta_B eps_B G_B + beta_F eps_F G_F), everything of the carried cotangent that needs no
// matrix product (base), and the per-step scalar / vd cotangents.  Same algebra as bridge_bwd.cu.  One block per particle.
struct WideNodeArgs {
    const float *part3, *c3t, *x, *sx, *zprev, *zup, *carry, *c, *mu, *logdiag, *betas, *epss;
    float *base, *vo, *vm, *r, *gmu_acc, *gls_acc, *g_beta, *g_eps, *g_os;
    int S3, N, d, j, K, pathwise, use_nn, nn_f;
    float out_scale, out_clip, clip_t, clip_q;
    const float* out_scale_dev;
};
__global__ void __launch_bounds__(WIDE_PB_MAX) wide_node_kernel(const WideNodeArgs a) {
    __shared__ float sh[32];
    const int n = blockIdx.x;
    const bool hasB = a.j > 0, hasF = a.j < a.K;
    const float out_scale = a.out_scale_dev ? __ldg(a.out_scale_dev) : a.out_scale;
    const float bB = hasB ? a.betas[a.j - 1] : 0.f, eB = hasB ? a.epss[a.j - 1] : 0.f;
    const float bF = hasF ? a.betas[a.j] : 0.f, eF = hasF ? a.epss[a.j] : 0.f;
    const float ombB = 1.0f - bB, ombF = 1.0f - bF, tsB = hasB ? 2.0f * eB : 1.f, tsF = hasF ? 2.0f * eF : 1.f;
    const float eFn = a.nn_f ? eF : 0.f;
    const float c = a.c[n];
    float gbB = 0.f, geB = 0.f, gbF = 0.f, geF = 0.f, gos = 0.f, rr = 0.f, xx = 0.f;
    for (int j = threadIdx.x; j < a.d; j += blockDim.x) {
        const size_t e = (size_t)n * a.d + j;
        float o = 0.f, nn = 0.f;
        if (a.use_nn) {
            o = a.c3t[j];
            for (int s = 0; s < a.S3; ++s) o += a.part3[((size_t)s * a.N + n) * a.d + j];
            nn = out_scale * fminf(fmaxf(o, -a.out_clip), a.out_clip);
        }
        const float sg = expf(a.logdiag[j]), ivar = 1.0f / (sg * sg);
        const float x = a.x[e], sx = a.sx[e];
        const float sq = -(x - a.mu[j]) * ivar;
        const float mk_t = (fabsf(sx) <= a.clip_t) ? 1.f : 0.f, mk_q = (fabsf(sq) <= a.clip_q) ? 1.f : 0.f;
        const float gu = fminf(fmaxf(sx, -a.clip_t), a.clip_t), gq = fminf(fmaxf(sq, -a.clip_q), a.clip_q);
        const float dc = gu - gq;
        const float uB = -(bB * gu + ombB * gq), uF = -(bF * gu + ombF * gq);
        float rB = 0.f, GB = 0.f, xs = 0.f, GF = 0.f;
        if (hasB) {
            const float meanB = (x - eB * uB) + eB * nn;
            rB = (a.zprev[e] - meanB) / tsB;
            GB = c * rB;
        }
        if (hasF) {
            const float meanF = (x - eF * uF) - eFn * nn;
            xs = (a.zup[e] - meanF) / tsF;
            GF = a.pathwise ? a.carry[e] : -c * xs;
        }
        rr = fmaf(rB, rB, rr);
        xx = fmaf(xs, xs, xx);
        const float v = eB * GB - eFn * GF;
        gos = fmaf(v, fminf(fmaxf(o, -a.out_clip), a.out_clip), gos);
        a.vo[e] = (a.use_nn && fabsf(o) <= a.out_clip) ? v * out_scale : 0.f;
        a.vm[e] = mk_t * (bB * eB * GB + bF * eF * GF);
        const float wq = eB * ombB * GB + eF * ombF * GF;
        a.gmu_acc[e] += wq * ivar * mk_q;
        a.gls_acc[e] += wq * mk_q * (-2.0f * sq);
        if (a.pathwise) {
            const float fpart = hasF ? (GF - c * a.r[e]) : c * sx;   // node K: terminal w += log p(z_K) (mcdboundingmachine.py:178)
            a.base[e] = fpart + GB - wq * ivar * mk_q;
        }
        a.r[e] = rB;
        gbF += eF * GF * dc;
        geF += GF * (-uF - (a.nn_f ? nn : 0.f) + (a.pathwise ? xs : 0.f));
        gbB += eB * GB * dc;
        geB += GB * (-uB + nn);
    }
    gbB = block_sum(gbB, sh); geB = block_sum(geB, sh);
    gbF = block_sum(gbF, sh); geF = block_sum(geF, sh);
    gos = block_sum(gos, sh);
    rr = block_sum(rr, sh); xx = block_sum(xx, sh);
    if (threadIdx.x == 0) {
        if (hasF) {
            atomicAdd(a.g_beta + a.j, gbF);
            atomicAdd(a.g_eps + a.j, a.pathwise ? geF : geF - c * xx);   // -c |xi/s|^2 in the log-var mode
        }
        if (hasB) {
            atomicAdd(a.g_beta + a.j - 1, gbB);
            atomicAdd(a.g_eps + a.j - 1, geB + c * rr);                  // c |r|^2
        }
        if (a.use_nn && a.g_os) atomicAdd(a.g_os, gos);
    }
}

// carry_{j-1} = base + H vm + dx,  H vm = -K^-1 vm - a exp(x) vm  (vm already carries the beta eps weights of both uses),
// dx = dA1[:, :d] + dp1 U1^T (already summed into dxnet).
struct WideNodeCombineArgs {
    const float *kpart, *x, *base, *vm, *dxnet;
    float* out;
    int S, N, d, use_nn;
    float area;
    int generic;   // callback targets: kpart holds H vm itself
};
__global__ void wide_node_combine_kernel(const WideNodeCombineArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.N * a.d) return;
    const int n = i / a.d, j = i % a.d;
    float kd = 0.f;
    for (int s = 0; s < a.S; ++s) kd += a.kpart[((size_t)s * a.N + n) * a.d + j];
    const float hv = a.generic ? kd : -kd - a.area * expf(a.x[i]) * a.vm[i];
    a.out[i] = a.base[i] + hv + (a.use_nn ? a.dxnet[i] : 0.f);
}

// g_mu[j] = sum_n (gmu_acc + adj), g_ls[j] = sum_n (gls_acc + adj (z0 - mu) + c)
__global__ void wide_vd_final_kernel(const float* __restrict__ gmu_acc, const float* __restrict__ gls_acc, const float* __restrict__ adj,
                                     const float* __restrict__ z0, const float* __restrict__ c, const float* __restrict__ mu,
                                     int N, int d, int pathwise, float* __restrict__ g_mu, float* __restrict__ g_ls) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= d) return;
    float m = 0.f, l = 0.f;
    for (int n = 0; n < N; ++n) {
        const size_t e = (size_t)n * d + j;
        m += gmu_acc[e]; l += gls_acc[e] + c[n];
        if (pathwise) { m += adj[e]; l += adj[e] * (z0[e] - mu[j]); }
    }
    if (g_mu) g_mu[j] = m;
    if (g_ls) g_ls[j] = l;
}

struct WideBwdWs {
    size_t zp, z, spp, sp, adj, abar, G, vo, vm, r, dx, gmu, gls, lp, c, A1, A2, s1, s2, dA, dp, part, partT, total;
    // per-point stacks [T][N][.] kept for the end-of-pass weight-gradient products: network input x, A1, A2, output cotangent
    // vo, dp2, dp1; dA1 = scratch for the layer-1 cotangent
    size_t Xs, A1s, A2s, VOs, DPs, DP1s, dA1;
};
static WideBwdWs wide_bwd_layout(long long N, int d, int HP, int T = 0) {
    WideBwdWs L;
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 3) & ~(size_t)3; return r; };
    const size_t nd = (size_t)N * d, nh = (size_t)N * HP;
    L.zp = take(nd); L.z = take(nd); L.spp = take(nd); L.sp = take(nd); L.adj = take(nd); L.abar = take(nd);
    L.G = take(nd); L.vo = take(nd); L.vm = take(nd); L.r = take(nd); L.dx = take(nd); L.gmu = take(nd); L.gls = take(nd);
    L.lp = take(N); L.c = take(N);
    L.A1 = take(nh); L.A2 = take(nh); L.s1 = take(nh); L.s2 = take(nh); L.dA = take(nh); L.dp = take(nh);
    const int mx = HP > d ? HP : d;
    L.part = take((size_t)64 * N * mx);
    L.partT = take((size_t)64 * N * d);     // K^-1 products (score, Hessian-vector product) on the side stream
    L.Xs = L.A1s = L.A2s = L.VOs = L.DPs = L.DP1s = L.dA1 = 0;
    if (T > 0) {
        L.Xs = take((size_t)T * nd); L.VOs = take((size_t)T * nd);
        L.A1s = take((size_t)T * nh); L.A2s = take((size_t)T * nh); L.DPs = take((size_t)T * nh); L.DP1s = take((size_t)T * nh);
        L.dA1 = take(nh);
    }
    L.total = o;
    return L;
}
size_t wide_bwd_workspace_bytes(long long N, int d, int HP, int K) { return wide_bwd_layout(N, d, HP > 0 ? HP : 8, HP > 0 ? K + 1 : 0).total * sizeof(float); }

int launch_wide_bwd(const BridgeArgs& a, const cmcd_target* tg, int D, cudaStream_t st, int num_sms, const float* cot_negw,
                    float* g_vd_mean, float* g_vd_logdiag, float* g_betas, float* g_eps, const cmcd_net_grad* g,
                    void* ws, size_t ws_bytes) {
    const long long N = a.N;
    const int d = D, K = a.K;
    const NetView& nv = a.net;
    const bool has_net = nv.arch != CMCD_ARCH_NONE;
    const int HP = has_net ? nv.HP : 8;
    if (has_net && nv.arch != CMCD_ARCH_GEFFNER) { set_error("wide path (lgcp / callback targets): only nn_arch=geffner is implemented (README.md:63 config)"); return 2; }
    if (HP & 3) { set_error("wide path: hidden_pad must be a multiple of 4"); return 2; }
    const WideBwdWs L = wide_bwd_layout(N, d, HP, has_net ? K + 1 : 0);
    if (!ws || ws_bytes < L.total * sizeof(float)) { set_error("wide_bwd: workspace too small (%zu < %zu)", ws_bytes, L.total * sizeof(float)); return 2; }
    float* f = (float*)ws;
    float *zp = f + L.zp, *z = f + L.z, *spp = f + L.spp, *sp = f + L.sp, *adj = f + L.adj, *abar = f + L.abar;
    float *G = f + L.G, *vo = f + L.vo, *vm = f + L.vm, *r = f + L.r, *dx = f + L.dx, *gmu = f + L.gmu, *gls = f + L.gls;
    float *lp = f + L.lp, *c = f + L.c, *A1 = f + L.A1, *A2 = f + L.A2, *s1 = f + L.s1, *s2 = f + L.s2, *dA = f + L.dA, *dp = f + L.dp;
    float *part = f + L.part, *partT = f + L.partT;
    float *Xs = f + L.Xs, *A1s = f + L.A1s, *A2s = f + L.A2s, *VOs = f + L.VOs, *DPs = f + L.DPs, *DP1s = f + L.DP1s, *dA1 = f + L.dA1;
    WideAsync* as = (tg->kind == CMCD_TARGET_CALLBACK) ? nullptr : wide_async();
    cudaStream_t st_t = as ? as->side : st;
    int ST = 1;
    const bool cais = (a.mode == CMCD_MODE_CAIS_SN || a.mode == CMCD_MODE_CAIS_VAR_SN);
    const bool pathwise = a.mode != CMCD_MODE_CAIS_VAR_SN;
    const bool nn_b = (a.mode != CMCD_MODE_ULA) && has_net, nn_f = cais && has_net;
    const int nd = (int)N * d, nh = (int)N * HP;
    const int T = K + 1;
    int S = 1;

    // outputs are overwritten: zero the accumulators
    CMCD_CUDA_OK(cudaMemsetAsync(gmu, 0, (size_t)nd * sizeof(float), st));
    CMCD_CUDA_OK(cudaMemsetAsync(gls, 0, (size_t)nd * sizeof(float), st));
    CMCD_CUDA_OK(cudaMemsetAsync(adj, 0, (size_t)nd * sizeof(float), st));
    if (g_betas && K) CMCD_CUDA_OK(cudaMemsetAsync(g_betas, 0, (size_t)K * sizeof(float), st));
    if (g_eps && K) CMCD_CUDA_OK(cudaMemsetAsync(g_eps, 0, (size_t)K * sizeof(float), st));
    float *gU1 = nullptr, *gW2 = nullptr, *gW3 = nullptr, *gc1 = nullptr, *gc2 = nullptr, *gc3 = nullptr, *gos = nullptr;
    if (g && has_net) {
        gU1 = g->U1; gW2 = g->W2; gW3 = g->W3; gc1 = g->c1; gc2 = g->c2; gc3 = g->c3; gos = g->out_scale;
        if (!gU1 || !gW2 || !gW3 || !gc1 || !gc2 || !gc3 || !gos) { set_error("wide_bwd: every network cotangent buffer is required"); return 2; }
        // gU1 / gW2 / gW3 are overwritten by the end-of-pass products; the bias tables get every row a point used and zeros elsewhere
        CMCD_CUDA_OK(cudaMemsetAsync(gc1, 0, (size_t)T * HP * sizeof(float), st));
        CMCD_CUDA_OK(cudaMemsetAsync(gc2, 0, (size_t)T * HP * sizeof(float), st));
        CMCD_CUDA_OK(cudaMemsetAsync(gc3, 0, (size_t)T * d * sizeof(float), st));
        CMCD_CUDA_OK(cudaMemsetAsync(gos, 0, sizeof(float), st));
        if (g->U2) CMCD_CUDA_OK(cudaMemsetAsync(g->U2, 0, (size_t)d * HP * sizeof(float), st));
        if (g->U3) CMCD_CUDA_OK(cudaMemsetAsync(g->U3, 0, (size_t)d * d * sizeof(float), st));
    }
    // scalar cotangents need a buffer even if the caller passes NULL
    float* gbeta_buf = g_betas ? g_betas : part;   // (NULL only when K == 0: never touched)
    float* geps_buf = g_eps ? g_eps : part;

    auto ew = [&](int n) { return dim3((unsigned)((n + 255) / 256)); };
    const bool generic = tg->kind == CMCD_TARGET_CALLBACK;
    auto fork = [&]() -> int {       // the side stream picks up after everything enqueued on the caller's stream so far
        if (as) { CMCD_CUDA_OK(cudaEventRecord(as->fork, st)); CMCD_CUDA_OK(cudaStreamWaitEvent(as->side, as->fork, 0)); }
        return 0;
    };
    auto join = [&]() -> int {       // the caller's stream waits for the side chain
        if (as) { CMCD_CUDA_OK(cudaEventRecord(as->join, as->side)); CMCD_CUDA_OK(cudaStreamWaitEvent(st, as->join, 0)); }
        return 0;
    };
    auto target_at = [&](const float* x, float* score) -> int {     // on the side stream (between fork and join)
        if (generic) {
            if (int rc = tg->eval(tg->user, (void*)st, x, (int64_t)N, d, nullptr, lp, score, nullptr)) { set_error("target callback failed (rc=%d)", rc); return 3; }
            return 0;
        }
        if (int rc = run_gemm(st_t, x, d, tg->lgcp_mu0, tg->lgcp_kinv, d, (int)N, d, d, num_sms, partT, &ST)) return rc;
        wide_target_fin_kernel<<<(unsigned)N, wide_pb(d), 0, st_t>>>(partT, ST, (int)N, d, x, tg->lgcp_counts, tg->lgcp_mu0,
                                                              tg->lgcp_log_norm, tg->lgcp_bin_area, score, lp);
        CMCD_CUDA_OK(cudaGetLastError());
        return 0;
    };
    // point j keeps its network input and activations in slice j of the stacks (A1 = A1s + j nh, ...)
    auto net_fwd_store = [&](const float* x, int t, int j, int* S3) -> int {
        const int blk = (nh + 255) / 256;
        float *A1j = A1s + (size_t)j * nh, *A2j = A2s + (size_t)j * nh;
        if (int rc = run_gemm(st, x, d, 0.f, nv.U1, HP, (int)N, d, HP, num_sms, part, &S, false)) return rc;
        wide_l1_fin_kernel<<<blk, 256, 0, st>>>(part, S, (int)N, HP, d, nv.c1 + (size_t)t * HP, x, 1, s1, A1j);
        CMCD_CUDA_OK(cudaGetLastError());
        if (int rc = run_gemm(st, A1j, HP, 0.f, nv.W2, HP, (int)N, HP, HP, num_sms, part, &S, false)) return rc;
        wide_l2_fin_kernel<<<blk, 256, 0, st>>>(part, S, (int)N, HP, nv.c2 + (size_t)t * HP, A1j, 1.f, s2, A2j);
        CMCD_CUDA_OK(cudaGetLastError());
        return run_gemm(st, A2j, HP, 0.f, nv.W3, d, (int)N, HP, d, num_sms, part, S3);
    };
    // network VJP for cotangent vo (slice j of VOs) on the raw output at point j: dx -> `dx`; the factors of the parameter
    // cotangents (dp2, dp1 next to x, A1, A2, vo) stay in their slices for the end-of-pass products
    auto net_bwd = [&](int j) -> int {
        float *voj = VOs + (size_t)j * nd, *dpj = DPs + (size_t)j * nh, *dp1j = DP1s + (size_t)j * nh;
        // dA2 = vo W3^T ; dp2 = dA2 * softplus'(pre2)
        if (int rc = run_gemm_t(st, num_sms, voj, d, nv.W3, d, (int)N, HP, d, nullptr, 0, s2, HP, dA, dpj, HP)) return rc;
        // dA1 = dA2 + dp2 W2^T ; dp1 = dA1 * softplus'(pre1)
        if (int rc = run_gemm_t(st, num_sms, dpj, HP, nv.W2, HP, (int)N, HP, HP, dA, HP, s1, HP, dA1, dp1j, HP)) return rc;
        // dx = dA1[:, :d] + dp1 U1^T
        if (int rc = run_gemm_t(st, num_sms, dp1j, HP, nv.U1, HP, (int)N, d, HP, dA1, HP, nullptr, 0, dx, nullptr, d)) return rc;
        return 0;
    };

    wide_neg_kernel<<<ew((int)N), 256, 0, st>>>(cot_negw, (int)N, c);
    // trajectory rows z_j in particle-major form: with a network every point keeps its row in slice j of Xs (the end-of-pass
    // weight-gradient products read them again); without one three rotating buffers do.  `abar` holds the carried cotangent.
    float* rot[3] = {zp, z, spp};
    auto row = [&](int j) -> float* { return has_net ? Xs + (size_t)j * nd : rot[((j % 3) + 3) % 3]; };
    wide_gather_kernel<<<ew(nd), 256, 0, st>>>(a.traj + (size_t)K * d * N, (int)N, d, row(K));
    CMCD_CUDA_OK(cudaGetLastError());
    CMCD_CUDA_OK(cudaMemsetAsync(r, 0, (size_t)nd * sizeof(float), st));
    CMCD_CUDA_OK(cudaMemsetAsync(abar, 0, (size_t)nd * sizeof(float), st));
    const int t0 = cais ? 0 : -1;   // table row of node j: t0 + j  (MCD_ULA_sn: NN(z_j, j-1), mcd_over_orig.py:45)
    float* x = row(K);
    for (int j = K; j >= 0; --j) {
        const bool hasB = j > 0;
        const int t = t0 + j;
        const bool use_nn = has_net && K > 0 && (cais || (nn_b && hasB));
        x = row(j);
        float *zprev = row(j - 1 >= 0 ? j - 1 : j), *zup = row(j + 1 <= K ? j + 1 : j);   // (absent neighbours are never read)
        float* voj = use_nn ? VOs + (size_t)j * nd : vo;
        if (hasB) {
            wide_gather_kernel<<<ew(nd), 256, 0, st>>>(a.traj + (size_t)(j - 1) * d * N, (int)N, d, zprev);
            CMCD_CUDA_OK(cudaGetLastError());
        }
        int S3 = 1;
        if (int rc = fork()) return rc;
        if (int rc = target_at(x, sp)) return rc;                              // side stream: score at x
        if (use_nn) { if (int rc = net_fwd_store(x, t, j, &S3)) return rc; }   // caller's stream: network recompute
        if (int rc = join()) return rc;
        WideNodeArgs h{};
        h.part3 = part; h.c3t = use_nn ? nv.c3 + (size_t)t * d : nullptr; h.x = x; h.sx = sp; h.zprev = zprev; h.zup = zup; h.carry = abar;
        h.c = c; h.mu = a.vd_mean; h.logdiag = a.vd_logdiag; h.betas = a.betas; h.epss = a.eps;
        h.base = G; h.vo = voj; h.vm = vm; h.r = r; h.gmu_acc = gmu; h.gls_acc = gls; h.g_beta = gbeta_buf; h.g_eps = geps_buf; h.g_os = gos;
        h.S3 = S3; h.N = (int)N; h.d = d; h.j = j; h.K = K; h.pathwise = pathwise; h.use_nn = use_nn; h.nn_f = nn_f;
        h.out_scale = nv.out_scale; h.out_scale_dev = nv.out_scale_dev; h.out_clip = nv.out_clip; h.clip_t = a.clip_t; h.clip_q = a.clip_q;
        wide_node_kernel<<<(unsigned)N, wide_pb(d), 0, st>>>(h);
        CMCD_CUDA_OK(cudaGetLastError());
        // the Hessian-vector product K^-1 vm (side stream) overlaps the network pull-back (caller's stream)
        if (pathwise && !generic) {
            if (int rc = fork()) return rc;
            if (int rc = run_gemm(st_t, vm, d, 0.f, tg->lgcp_kinv, d, (int)N, d, d, num_sms, partT, &ST)) return rc;
        }
        if (use_nn) { if (int rc = net_bwd(j)) return rc; }
        if (pathwise) {
            if (generic) {   // H(x) vm from the caller's batched Hessian-vector product
                if (int rc = tg->eval(tg->user, (void*)st, x, (int64_t)N, d, vm, nullptr, nullptr, partT)) { set_error("target callback failed (rc=%d)", rc); return 3; }
                ST = 1;
            } else if (int rc = join()) return rc;
            WideNodeCombineArgs cb{};
            cb.kpart = partT; cb.x = x; cb.base = G; cb.vm = vm; cb.dxnet = dx; cb.out = abar;
            cb.S = ST; cb.N = (int)N; cb.d = d; cb.use_nn = use_nn; cb.area = tg->lgcp_bin_area; cb.generic = generic ? 1 : 0;
            wide_node_combine_kernel<<<ew(nd), 256, 0, st>>>(cb);
            CMCD_CUDA_OK(cudaGetLastError());
        }
    }
    // end of pass: the weight cotangents as one product per layer over all points' stacked factors, the per-step bias tables as
    // column sums of the same stacks
    if (g && has_net) {
        const bool any_nn = K > 0 && (cais || nn_b);
        if (any_nn) {
            const int j_lo = cais ? 0 : 1, P = K + 1 - j_lo, R = P * (int)N;
            const float *A2p = A2s + (size_t)j_lo * nh, *A1p = A1s + (size_t)j_lo * nh, *VOp = VOs + (size_t)j_lo * nd;
            const float *DPp = DPs + (size_t)j_lo * nh, *DP1p = DP1s + (size_t)j_lo * nh, *Xp = Xs + (size_t)j_lo * nd;
            auto tiles = [](int n) { return (unsigned)((n + WGR_T - 1) / WGR_T); };
            wide_wgrad_kernel<<<dim3(tiles(d), tiles(HP)), 256, 0, st>>>(A2p, HP, VOp, d, R, HP, d, gW3, d);
            wide_wgrad_kernel<<<dim3(tiles(HP), tiles(HP)), 256, 0, st>>>(A1p, HP, DPp, HP, R, HP, HP, gW2, HP);
            wide_wgrad_kernel<<<dim3(tiles(HP), tiles(d)), 256, 0, st>>>(Xp, d, DP1p, HP, R, d, HP, gU1, HP);
            wide_colsum_rows_kernel<<<dim3((unsigned)((d + 255) / 256), P), 256, 0, st>>>(VOp, d, (int)N, d, gc3, d);      // rows t0 + j_lo = 0 ...
            wide_colsum_rows_kernel<<<dim3((unsigned)((HP + 255) / 256), P), 256, 0, st>>>(DPp, HP, (int)N, HP, gc2, HP);
            wide_colsum_rows_kernel<<<dim3((unsigned)((HP + 255) / 256), P), 256, 0, st>>>(DP1p, HP, (int)N, HP, gc1, HP);
            CMCD_CUDA_OK(cudaGetLastError());
        } else {
            CMCD_CUDA_OK(cudaMemsetAsync(gU1, 0, (size_t)d * HP * sizeof(float), st));
            CMCD_CUDA_OK(cudaMemsetAsync(gW2, 0, (size_t)HP * HP * sizeof(float), st));
            CMCD_CUDA_OK(cudaMemsetAsync(gW3, 0, (size_t)HP * d * sizeof(float), st));
        }
    }
    adj = abar; zp = row(0);   // after node 0: carried cotangent = dL/dz_0, row(0) = z_0
    wide_vd_final_kernel<<<ew(d), 256, 0, st>>>(gmu, gls, adj, zp, c, a.vd_mean, (int)N, d, pathwise ? 1 : 0, g_vd_mean, g_vd_logdiag);
    CMCD_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace cmcd
