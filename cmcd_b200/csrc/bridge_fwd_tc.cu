// Forward bridge kernel, tensor-core variant (hidden width 64): the 64x64 layer of the drift network runs as
// tcgen05 tiles over 128-particle batches, everything else stays per-thread in registers.
//
// Same contract as bridge_fwd_kernel (bridge_fwd.cu): replaces vmap(compute_log_elbo)
// (src/mcdboundingmachine.py:126-205) over src/mcd_cais.py:46-89 / src/mcd_cais_var.py:56-101 /
// src/mcd_over_orig.py:18-55 with apply_fun_sn = PISNet (src/nn_dds.py:145-164) or the geffner net
// (src/nn.py:66-70) in table form (include/cmcd_b200.h, cmcd_net).
//
// Mapping: one CTA = 128 threads = 128 particles = the 128 TMEM lanes.  Thread p owns particle p of the tile:
//   layer 1   a1 = act(U1^T x + c1[t])  per thread; split into tf32 hi/lo and written with tcgen05.st into the
//             thread's own TMEM lane (columns AH.., AL..) -- this *is* the A operand [128 x 64] of the MMA;
//   layer 2   every thread arrives on an mbarrier; one elected lane of warp 0 issues 24 x tcgen05.mma kind::tf32
//             (M=128, N=64, K=8): D = A_hi B_lo + A_lo B_hi + A_hi B_hi
//             with B = W2^T split hi/lo once per CTA in shared memory (K-major core-matrix tiles);
//   epilogue  each thread reads its lane of D with tcgen05.ld (64 columns), adds c2[t], activation, and folds
//             the 64 x d output layer in registers.
// While the MMA batch runs (~1.6k cycles) the thread does the work that does not depend on the network output:
// threefry split + Gaussian for the step, the target score at z', the key advance.  Two CTAs per SM (256 TMEM
// columns each) alternate so the CUDA cores always have one tile's epilogue / layer 1 to run.
#include <cuda_bf16.h>

#include <cstdlib>

#include "common.cuh"
#include "umma.cuh"

namespace cmcd {

constexpr int TC_PB = 128;   // particles per CTA (TMEM lanes)
constexpr int TC_H = 64;     // hidden width of this path
// TMEM: 128 columns (accumulator D, A_hi as tf32) + 32 columns (A_lo as packed bf16 pairs) = 160 per CTA -> three CTAs per SM
constexpr uint32_t TC_COL_D = 0, TC_COL_AH = 64, TC_COLS_MAIN = 128, TC_COLS_LO = 32;
constexpr int TC_B_BYTES = TC_H * TC_H * 4;   // one 64x64 fp32 operand tile
constexpr int TC_B16_BYTES = TC_H * TC_H * 2; // the bf16 copy of B_hi for the A_lo pass
constexpr int TC_CTAS_PER_SM = 3;

template <int D>
struct TcCtx {
    const float *sU1, *sU2, *sW3, *sU3, *sW3T;   // shared memory (sW3T: W3 transposed, [D][64])
    const float *c1, *c2, *c3;            // global per-step tables [T][64], [T][64], [T][D]
    const float* tab;                     // this warp's staged rows c1[t] | c2[t] of the current half-step (shared memory)
    float out_scale, out_clip;
    uint32_t tmem_base, tmem_lane;        // main allocation (D | A_hi): base; base + this warp's lane quarter
    uint32_t tmem_lo_base, tmem_lo_lane;  // second allocation: A_lo as packed bf16 pairs
    uint64_t bhi, blo, bhi16;             // shared-memory descriptors of the B tiles (first K block)
    uint64_t* mbar;                       // MMA batch complete
    uint64_t* mbar_ready;                 // A operand staged by all 128 threads
    uint32_t parity, parity_ready;
    // TMA staging of the table rows (a.tab_tma): CTA-shared double buffer, one mbarrier per buffer, phase bit per buffer
    float* tab_shared;                    // [2][2 * 64]
    uint64_t* tab_bar;                    // [2]
    uint32_t tab_parity;                  // bit b = phase of buffer b
    int tma;
};

// numpyro Normal.log_prob summed over dims (src/mcd_utils.py:19-21)
template <int D>
__device__ __forceinline__ float gauss_logprob_tc(const float (&x)[D], const float (&mean)[D], float scale, float lognorm) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        const float v = (x[j] - mean[j]) / scale;
        s += -0.5f * v * v - lognorm;
    }
    return s;
}

// layer 1 -> TMEM A operand -> issue the MMA batch.  skipacc += a1 W3 (geffner residual), 0 otherwise.
template <int D, int ACT>
__device__ __forceinline__ void tc_net_issue(TcCtx<D>& cx, int t, int next_t, const float (&x)[D], float (&skipacc)[D]) {
    constexpr bool skip = (ACT == ACT_SOFTPLUS);
    const float4* __restrict__ c1v = reinterpret_cast<const float4*>(cx.tab);
    f32x2_t xb[D];
#pragma unroll
    for (int a = 0; a < D; ++a) xb[a] = pk2(x[a], x[a]);
#pragma unroll
    for (int m = 0; m < D; ++m) skipacc[m] = 0.f;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
        uint32_t h[16], l[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 cc = c1v[c * 4 + q];
            f32x2_t p01 = pk2(cc.x, cc.y), p23 = pk2(cc.z, cc.w);   // two hidden units per FFMA2 (x broadcast pairs built once per node)
#pragma unroll
            for (int a = 0; a < D; ++a) {
                const float4 u = *reinterpret_cast<const float4*>(cx.sU1 + a * TC_H + c * 16 + q * 4);
                p01 = fma2(xb[a], pk2(u.x, u.y), p01);
                p23 = fma2(xb[a], pk2(u.z, u.w), p23);
            }
            float p[4];
            upk2(p01, p[0], p[1]); upk2(p23, p[2], p[3]);
            float av[4];
            if constexpr (ACT == ACT_GELU) {   // two activations per instruction slot (FFMA2)
                upk2(gelu_fast2(p[0], p[1]), av[0], av[1]);
                upk2(gelu_fast2(p[2], p[3]), av[2], av[3]);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) av[e] = act_tc<ACT>(p[e]);
            }
            float lov[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float a1 = av[e];
                if (skip) {
                    const int j = c * 16 + q * 4 + e;
#pragma unroll
                    for (int m = 0; m < D; ++m) skipacc[m] = fmaf(a1, cx.sW3[j * D + m], skipacc[m]);
                }
                float hi;
                umma::split_tf32(a1, hi, lov[e]);
                h[q * 4 + e] = __float_as_uint(hi);
            }
            // the low part only has to carry ~9 more bits: bf16, two K elements per TMEM column (even k in the low half)
            const __nv_bfloat162 l01 = __floats2bfloat162_rn(lov[0], lov[1]), l23 = __floats2bfloat162_rn(lov[2], lov[3]);
            l[q * 2 + 0] = *reinterpret_cast<const uint32_t*>(&l01);
            l[q * 2 + 1] = *reinterpret_cast<const uint32_t*>(&l23);
        }
        umma::tmem_st16(cx.tmem_lane + TC_COL_AH + c * 16, h);
        umma::tmem_st8(cx.tmem_lo_lane + c * 8, l);
    }
    umma::tmem_st_wait();
    umma::fence_before();
    umma::mbar_arrive(cx.mbar_ready);   // only the issuing thread waits for the 128 arrivals; the other warps move on
    if (threadIdx.x < 32) {   // warp 0 (converged): wait for the tile's 128 arrivals, one elected lane issues the batch
        umma::mbar_wait(cx.mbar_ready, cx.parity_ready);
        cx.parity_ready ^= 1u;
        umma::fence_after();
        if (umma::elect_one()) {
            if (cx.tma && next_t >= 0) {
                // every thread of the CTA has arrived for node t, i.e. is done with the rows of node t - 1: their buffer is free.
                // One thread requests both rows of node t + 1 through the TMA engine (two 256-byte bulk copies, one mbarrier).
                const int b = next_t & 1;
                umma::mbar_arrive_expect_tx(cx.tab_bar + b, 2 * TC_H * sizeof(float));
                umma::bulk_copy_g2s(cx.tab_shared + b * (2 * TC_H), cx.c1 + (size_t)next_t * TC_H, TC_H * sizeof(float), cx.tab_bar + b);
                umma::bulk_copy_g2s(cx.tab_shared + b * (2 * TC_H) + TC_H, cx.c2 + (size_t)next_t * TC_H, TC_H * sizeof(float), cx.tab_bar + b);
            }
            const uint32_t idesc = umma::make_idesc_tf32(128, TC_H), idesc16 = umma::make_idesc_bf16_k(128, TC_H);
            const uint32_t dcol = cx.tmem_base + TC_COL_D, ahi = cx.tmem_base + TC_COL_AH;
            // D = A_hi B_lo + A_lo B_hi + A_hi B_hi, small terms first: the fp32 accumulator truncates on every accumulate
            // (tools/umma_probe2.cu, test 3).  Next K block of a B tile: +256 B = +16 in the descriptor's address field.
#pragma unroll
            for (int k = 0; k < 8; ++k) umma::mma_tf32_ts(dcol, ahi + k * 8, cx.blo + (uint64_t)(k * 16), idesc, k > 0);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma::mma_f16_ts(dcol, cx.tmem_lo_base + k * 8, cx.bhi16 + (uint64_t)(k * 16), idesc16, 1);
#pragma unroll
            for (int k = 0; k < 8; ++k) umma::mma_tf32_ts(dcol, ahi + k * 8, cx.bhi + (uint64_t)(k * 16), idesc, 1);
            umma::commit(cx.mbar);
        }
        __syncwarp();
    }
}

// wait for the MMA batch, epilogue: out = out_scale * clamp(W3^T (act(D + c2[t] + U2^T x) [+ a1]) + U3^T x + c3[t])
template <int D, int ACT>
__device__ __forceinline__ void tc_net_finish(TcCtx<D>& cx, int t, const float (&x)[D], const float (&skipacc)[D], float (&out)[D]) {
    constexpr bool has_u = (ACT == ACT_SOFTPLUS);   // geffner: U2, U3 present
    float o[D];
#pragma unroll
    for (int m = 0; m < D; ++m) {
        float p = __ldg(cx.c3 + (size_t)t * D + m) + skipacc[m];
        if (has_u) {
#pragma unroll
            for (int a = 0; a < D; ++a) p = fmaf(x[a], cx.sU3[a * D + m], p);
        }
        o[m] = p;
    }
    const float4* __restrict__ c2v = reinterpret_cast<const float4*>(cx.tab + TC_H);
    f32x2_t O2[D];                        // dds: output-layer partial sums over (even, odd) hidden units
#pragma unroll
    for (int m = 0; m < D; ++m) O2[m] = pk2(0.f, 0.f);
    umma::mbar_wait(cx.mbar, cx.parity);
    cx.parity ^= 1u;
    umma::fence_after();
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
        uint32_t v[16];
        umma::tmem_ld16(cx.tmem_lane + TC_COL_D + c * 16, v);
        umma::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 cc = c2v[c * 4 + q];
            if constexpr (ACT == ACT_GELU) {
                // dds: no U2 term; bias add, two exact-erf GELUs and the output-layer products all on packed pairs of hidden units
                float p0, p1, p2, p3;
                upk2(add2(pk2(__uint_as_float(v[q * 4 + 0]), __uint_as_float(v[q * 4 + 1])), pk2(cc.x, cc.y)), p0, p1);
                upk2(add2(pk2(__uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3])), pk2(cc.z, cc.w)), p2, p3);
                const f32x2_t A01 = gelu_fast2(p0, p1), A23 = gelu_fast2(p2, p3);
#pragma unroll
                for (int m = 0; m < D; ++m) {
                    const float4 w = *reinterpret_cast<const float4*>(cx.sW3T + m * TC_H + c * 16 + q * 4);
                    O2[m] = fma2(A01, pk2(w.x, w.y), O2[m]);
                    O2[m] = fma2(A23, pk2(w.z, w.w), O2[m]);
                }
            } else {
                float p[4] = {__uint_as_float(v[q * 4 + 0]) + cc.x, __uint_as_float(v[q * 4 + 1]) + cc.y,
                              __uint_as_float(v[q * 4 + 2]) + cc.z, __uint_as_float(v[q * 4 + 3]) + cc.w};
                if (has_u) {
#pragma unroll
                    for (int a = 0; a < D; ++a) {
                        const float4 u = *reinterpret_cast<const float4*>(cx.sU2 + a * TC_H + c * 16 + q * 4);
                        p[0] = fmaf(x[a], u.x, p[0]); p[1] = fmaf(x[a], u.y, p[1]);
                        p[2] = fmaf(x[a], u.z, p[2]); p[3] = fmaf(x[a], u.w, p[3]);
                    }
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float av = act_tc<ACT>(p[e]);
                    const int j = c * 16 + q * 4 + e;
#pragma unroll
                    for (int m = 0; m < D; ++m) o[m] = fmaf(av, cx.sW3[j * D + m], o[m]);
                }
            }
        }
    }
    if constexpr (ACT == ACT_GELU) {
#pragma unroll
        for (int m = 0; m < D; ++m) {
            float e0, e1;
            upk2(O2[m], e0, e1);
            o[m] += e0 + e1;
        }
    }
    umma::fence_before();   // orders these tcgen05.ld before the next batch's writes to D (via the next mbarrier arrive / wait)
#pragma unroll
    for (int m = 0; m < D; ++m) out[m] = cx.out_scale * fminf(fmaxf(o[m], -cx.out_clip), cx.out_clip);
}

template <int D, int ACT, bool EV>
__global__ void __launch_bounds__(TC_PB, TC_CTAS_PER_SM) bridge_fwd_tc_kernel(const BridgeArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint32_t tmem_slot, tmem_slot_lo;
    __shared__ __align__(8) uint64_t mbar, mbar_ready, tab_bar[2];
    __shared__ __align__(128) float tab_shared[2][2 * TC_H];
    const int tid = threadIdx.x, warp = tid >> 5;
    const NetView& nv = a.net;
    uint8_t* sBhi = smem_raw;
    uint8_t* sBlo = smem_raw + TC_B_BYTES;
    uint8_t* sBhi16 = smem_raw + 2 * TC_B_BYTES;
    float* sf = reinterpret_cast<float*>(smem_raw + 2 * TC_B_BYTES + TC_B16_BYTES);
    float* sU1 = sf;
    float* sU2 = sU1 + D * TC_H;
    float* sW3 = sU2 + D * TC_H;
    float* sU3 = sW3 + TC_H * D;
    float* sW3T = sU3 + ((D * D + 3) & ~3);
    float* sTp = sW3T + D * TC_H;
    // B[n = j][k = i] = W2[i][j], split into tf32 hi / lo
    for (int idx = tid; idx < TC_H * TC_H; idx += TC_PB) {
        const int i = idx / TC_H, j = idx % TC_H;
        float hi, lo;
        umma::split_tf32(nv.W2[idx], hi, lo);
        const int off = umma::core_off(j, i, TC_H);
        *reinterpret_cast<float*>(sBhi + off) = hi;
        *reinterpret_cast<float*>(sBlo + off) = lo;
        *reinterpret_cast<__nv_bfloat16*>(sBhi16 + umma::core_off16(j, i, TC_H)) = __float2bfloat16(hi);
    }
    for (int i = tid; i < D * TC_H; i += TC_PB) { sU1[i] = nv.U1[i]; sU2[i] = nv.U2 ? nv.U2[i] : 0.f; }
    for (int i = tid; i < TC_H * D; i += TC_PB) { sW3[i] = nv.W3[i]; sW3T[(i % D) * TC_H + i / D] = nv.W3[i]; }
    for (int i = tid; i < D * D; i += TC_PB) sU3[i] = nv.U3 ? nv.U3[i] : 0.f;
    const int ntp = (a.tgt.kind == TGT_GMM || a.tgt.kind == TGT_MANY_GMM) ? a.tgt.ncomp * MIX_STRIDE : 0;
    for (int i = tid; i < ntp; i += TC_PB) sTp[i] = a.tgt.mix[i];
    float2* sMu = reinterpret_cast<float2*>(sTp + MIX_MAX * MIX_STRIDE);   // many_gmm: dense component means
    // per-warp double buffer for the per-step table rows c1[t] | c2[t] (cp.async one half-step ahead)
    float* sTab = reinterpret_cast<float*>(sMu + MIX_MAX) + warp * (2 * 2 * TC_H);
    const int lane_ = tid & 31;
    auto stage_tab = [&](int t, int buf) {   // 32 lanes x 16 B = c1 row (256 B) + c2 row (256 B)
        const float* src = (lane_ < 16 ? nv.c1 : nv.c2) + (size_t)t * TC_H + (lane_ & 15) * 4;
        umma::cp_async16(sTab + buf * (2 * TC_H) + lane_ * 4, src);
        umma::cp_async_commit();
    };
    const bool fast_gmm = (D == 2) && (a.tgt.kind == TGT_MANY_GMM);
    if (fast_gmm)
        many_gmm_stage_means(a.tgt, sMu, tid, TC_PB);
    const ManyGmmConst gc = many_gmm_const(a.tgt);
    if (warp == 0) { umma::tmem_alloc(&tmem_slot, TC_COLS_MAIN, false); umma::tmem_alloc(&tmem_slot_lo, TC_COLS_LO, true); }
    if (tid == 0) { umma::mbar_init(&mbar, 1); umma::mbar_init(&mbar_ready, TC_PB); umma::mbar_init(&tab_bar[0], 1); umma::mbar_init(&tab_bar[1], 1); }
    umma::fence_async_smem();   // generic-proxy writes of the B tiles -> visible to the tensor core (async proxy)
    umma::fence_before();
    __syncthreads();
    umma::fence_after();

    TcCtx<D> cx;
    cx.sU1 = sU1; cx.sU2 = sU2; cx.sW3 = sW3; cx.sU3 = sU3; cx.sW3T = sW3T;
    cx.c1 = nv.c1; cx.c2 = nv.c2; cx.c3 = nv.c3;
    cx.out_scale = net_out_scale(nv); cx.out_clip = nv.out_clip;
    cx.tmem_base = tmem_slot;
    cx.tmem_lane = tmem_slot + ((uint32_t)(warp * 32) << 16);
    cx.tmem_lo_base = tmem_slot_lo;
    cx.tmem_lo_lane = tmem_slot_lo + ((uint32_t)(warp * 32) << 16);
    cx.bhi16 = umma::make_desc(umma::smem_u32(sBhi16), 128, 16 * TC_H);
    cx.bhi = umma::make_desc(umma::smem_u32(sBhi), 128, 32 * TC_H);
    cx.blo = umma::make_desc(umma::smem_u32(sBlo), 128, 32 * TC_H);
    cx.mbar = &mbar; cx.parity = 0u;
    cx.mbar_ready = &mbar_ready; cx.parity_ready = 0u;
    cx.tab_shared = &tab_shared[0][0]; cx.tab_bar = tab_bar; cx.tab_parity = 0u; cx.tma = a.tab_tma;

    const bool cais = (a.mode == CMCD_MODE_CAIS_SN || a.mode == CMCD_MODE_CAIS_VAR_SN);
    const bool nn_b = (a.mode != CMCD_MODE_ULA);
    const bool nn_f = cais;
    const int K = a.K;

    float mu[D], sig[D];
#pragma unroll
    for (int j = 0; j < D; ++j) { mu[j] = a.vd_mean[j]; sig[j] = expf(a.vd_logdiag[j]); }

    const long long ntiles = (a.N + TC_PB - 1) / TC_PB;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long n_raw = tile * TC_PB + tid;
        const bool active = n_raw < a.N;
        const long long n = active ? n_raw : a.N - 1;   // tail lanes shadow the last particle (all 128 lanes take part in the MMA)
        constexpr bool ev = EV;   // mcd_utils.evolve entry: (z, rng_key_gen) given by the caller (mcd_utils.py:24-33), own instantiation
        Key k, ka;
        float z[D], xi[D];
        float w = 0.f;
        if constexpr (ev) {
#pragma unroll
            for (int j = 0; j < D; ++j) z[j] = a.z0[n * D + j];
            ka.k0 = a.keys[2 * n]; ka.k1 = a.keys[2 * n + 1];
        } else {   // z0 = sigma*xi + mu ; w = -log q(z0)   (vardist/diag_gauss.py:26-33,44-62)
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
        }
        if (a.traj && active) {
#pragma unroll
            for (int j = 0; j < D; ++j) a.traj[((size_t)0 * D + j) * a.N + n] = z[j];
        }
        float sp[D], dummy[D];
        float lp = 0.f;
        if (!ev) ka = split_first(k);   // mcdboundingmachine.py:162 (evolve entry: ka is the caller's rng_key_gen)
        k = split_second(ka);           // mcd_cais.py:94
        float wm = 0.f;
        // K + 1 nodes z_0 .. z_K, ONE network evaluation per node: in the CAIS modes NN(z_j, j) serves both the
        // backward-kernel mean of step j-1 (mcd_cais.py:78, called there as NN(z', i + 1)) and the forward-kernel mean of
        // step j (mcd_cais.py:60) -- same point, same time index, so the reference's 2K evaluations are K + 1 distinct
        // ones.  MCD_ULA_sn (mcd_over_orig.py:45): NN(z_j, j - 1), backward-kernel mean only.
        const int t0 = cais ? 0 : -1;
        if (K > 0 && nn_b) {                  // first evaluation: node 0 (CAIS) or node 1 (ULA_sn), t = 0 either way
            if (cx.tma) {
                __syncthreads();              // the previous tile's last node may still be reading buffer 0 in a lagging warp
                if (tid == 0) {
                    umma::mbar_arrive_expect_tx(&tab_bar[0], 2 * TC_H * sizeof(float));
                    umma::bulk_copy_g2s(&tab_shared[0][0], nv.c1, TC_H * sizeof(float), &tab_bar[0]);
                    umma::bulk_copy_g2s(&tab_shared[0][TC_H], nv.c2, TC_H * sizeof(float), &tab_bar[0]);
                }
            } else stage_tab(0, 0);
        }
        float x[D], mf[D], nnv[D], skipacc[D];
#pragma unroll
        for (int j = 0; j < D; ++j) { x[j] = z[j]; mf[j] = 0.f; }
        float beta = 0.f, eps = 0.f, scale = 1.f;
        for (int nd = 0; nd <= K; ++nd) {
            const int t = t0 + nd;
            const bool use_nn = nn_b && (cais || nd > 0) && K > 0;
            if (use_nn) {
                // table rows of this node were requested one node ago; request the next ones
                if (cx.tma) {
                    const int b = t & 1;
                    umma::mbar_wait(&tab_bar[b], (cx.tab_parity >> b) & 1u);
                    cx.tab_parity ^= 1u << b;
                    cx.tab = &tab_shared[b][0];
                } else {
                    umma::cp_async_wait_all();
                    __syncwarp();
                    cx.tab = sTab + (t & 1) * (2 * TC_H);
                    if (nd < K) stage_tab(t + 1, (t + 1) & 1);
                }
                tc_net_issue<D, ACT>(cx, t, nd < K ? t + 1 : -1, x, skipacc);
            }
            // ---- work that does not depend on the network output overlaps the MMA batch ----
            if (fast_gmm) { float d0, d1; lp = many_gmm_eval<false>(gc, sMu, x[0], x[1], sp[0], sp[1], 0.f, 0.f, d0, d1); }
            else lp = target_eval<D, false>(a.tgt, sTp, x, sp, dummy, dummy);
            float gu[D], gq[D];
#pragma unroll
            for (int j = 0; j < D; ++j) {
                const float sq = -((x[j] - mu[j]) / sig[j]) / sig[j];
                gu[j] = fminf(fmaxf(sp[j], -a.clip_t), a.clip_t);
                gq[j] = fminf(fmaxf(sq, -a.clip_q), a.clip_q);
            }
            if (nd < K) step_keys_and_normal<D>(k, xi);   // split + Gaussian + key advance of step nd (two interleaved threefry batches)
#pragma unroll
            for (int j = 0; j < D; ++j) nnv[j] = 0.f;
            if (use_nn) tc_net_finish<D, ACT>(cx, t, x, skipacc, nnv);
            if (nd > 0) {   // backward kernel of step nd - 1 (beta, eps, scale, mf still hold that step's values), weight update
                float mb[D];
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    const float ub = -(beta * gu[j] + (1.0f - beta) * gq[j]);
                    mb[j] = x[j] - eps * ub;
                    mb[j] = mb[j] + eps * nnv[j];
                }
                const float lognorm = logf(2.5066282746310002f * scale);
                const float fk = gauss_logprob_tc<D>(x, mf, scale, lognorm);
                const float bk = gauss_logprob_tc<D>(z, mb, scale, lognorm);
                wm += bk - fk;
#pragma unroll
                for (int j = 0; j < D; ++j) z[j] = x[j];
                if (a.traj && active) {
#pragma unroll
                    for (int j = 0; j < D; ++j) a.traj[((size_t)nd * D + j) * a.N + n] = z[j];
                }
            }
            if (nd < K) {   // forward kernel of step nd: mean, sample z_{nd+1}
                beta = __ldg(a.betas + nd); eps = __ldg(a.eps + nd);
                scale = sqrtf(2.0f * eps);
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    const float uf = -(beta * gu[j] + (1.0f - beta) * gq[j]);
                    mf[j] = z[j] - eps * uf;
                    if (nn_f) mf[j] = mf[j] - eps * nnv[j];
                    x[j] = mf[j] + scale * xi[j];
                }
            }
        }
        w += wm;
        w += lp;
        if (active) {
            // evolve returns the steps' log-ratio sum (mcd_cais.py:98-99); compute_log_elbo adds -log q(z_0) and log p(z_K)
            a.out_negw[n] = ev ? wm : -w;
#pragma unroll
            for (int j = 0; j < D; ++j) a.out_z[n * D + j] = z[j];
        }
    }
    umma::fence_before();
    __syncthreads();
    if (warp == 0) { umma::tmem_dealloc(cx.tmem_base, TC_COLS_MAIN); umma::tmem_dealloc(cx.tmem_lo_base, TC_COLS_LO); }
}

template <int D, int ACT>
static int launch_fwd_tc_t(const BridgeArgs& a_in, cudaStream_t st, int num_sms) {
    const bool ev = a_in.z0 != nullptr;
    // request > 227/4 KB so that at most three CTAs (3 x 160 TMEM columns) share an SM
    size_t smem = 2 * TC_B_BYTES + TC_B16_BYTES + (2 * D * TC_H + 2 * TC_H * D + D * D + 4 + MIX_MAX * MIX_STRIDE + 2 * MIX_MAX + 4 * 4 * TC_H + 8) * sizeof(float);
    if (smem < 58 * 1024) smem = 58 * 1024;
    auto kern = ev ? bridge_fwd_tc_kernel<D, ACT, true> : bridge_fwd_tc_kernel<D, ACT, false>;
    CMCD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // TMA bulk copies need 16-byte aligned rows; CMCD_TAB_TMA=0 keeps the per-warp cp.async staging (A/B runs)
    BridgeArgs a = a_in;
    const char* env = std::getenv("CMCD_TAB_TMA");
    a.tab_tma = (!env || env[0] != '0') && !((reinterpret_cast<uintptr_t>(a.net.c1) | reinterpret_cast<uintptr_t>(a.net.c2)) & 15);
    const long long ntiles = (a.N + TC_PB - 1) / TC_PB;
    long long grid = (long long)TC_CTAS_PER_SM * num_sms;
    if (grid > ntiles) grid = ntiles;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, TC_PB, smem, st>>>(a);
    CMCD_CUDA_OK(cudaGetLastError());
    return 0;
}

// hidden_pad == 64 networks (dds; geffner with x_dim + emb_dim in 57..64); K >= 1 and a network in use
bool fwd_tc_supported(const BridgeArgs& a, int D) {
    return a.net.arch != CMCD_ARCH_NONE && a.net.HP == TC_H && a.K >= 1 && a.mode != CMCD_MODE_ULA && (D == 2 || D == 10);
}

int launch_bridge_fwd_tc(const BridgeArgs& a, int D, cudaStream_t st, int num_sms) {
    const bool dds = a.net.arch == CMCD_ARCH_DDS;
    if (D == 2) return dds ? launch_fwd_tc_t<2, ACT_GELU>(a, st, num_sms) : launch_fwd_tc_t<2, ACT_SOFTPLUS>(a, st, num_sms);
    if (D == 10) return dds ? launch_fwd_tc_t<10, ACT_GELU>(a, st, num_sms) : launch_fwd_tc_t<10, ACT_SOFTPLUS>(a, st, num_sms);
    set_error("bridge_fwd_tc: dim=%d has no instantiation", D);
    return 2;
}

}  // namespace cmcd
