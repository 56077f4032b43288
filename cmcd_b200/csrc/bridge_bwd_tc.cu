// Reverse-mode (adjoint) bridge kernel, tensor-core variant for the dds network (hidden width 64, d = 2).
//
// Same contract as bridge_bwd_kernel (bridge_bwd.cu): replaces the transposed scan of
// jax.grad(compute_bound / compute_bound_var) (src/main.py:174-176 over src/mcdboundingmachine.py:126-231,
// src/mcd_cais.py:46-89, src/mcd_cais_var.py:56-101, src/mcd_over_orig.py:18-55) with apply_fun_sn = PISNet
// (src/nn_dds.py:145-164) in table form.  The per-step algebra is the one documented at the top of bridge_bwd.cu.
//
// Mapping: one CTA per SM, 256 threads = two independent 128-particle tiles (one warpgroup each, thread = particle =
// TMEM lane) that alternate between CUDA-core phases and tensor-core batches:
//   GEMM1  pre2 = a1 W2            A = a1 (tf32 hi/lo, TMEM, written by the owning threads), B = W2^T tile (smem)
//   GEMM2  da1  = dp2 W2^T         A = dp2 (tf32 hi/lo, TMEM),                              B = W2 tile (smem)
//   WGRAD  gW2 += a1^T dp2         kind::f16, bf16 hi+lo operands staged row-major per particle in shared memory
//                                  (MN-major, 128B swizzle), M = 128 = [a1_hi ; a1_lo] stacked, N = 64, K = 128 particles;
//                                  accumulates in TMEM across steps, flushed to a thread-private row every few steps
//                                  (the tensor-core accumulator truncates, tools/umma_probe2.cu test 3).
// TMEM per tile (256 columns): A_hi | A_lo | D (pre2, then da1) | gW2 accumulator.  Between GEMM1 and GEMM2 the dead
// A region doubles as per-thread scratch for a2 / act'(pre2).  The skinny gradients (per-step bias tables, U1, W3)
// are reduced over the 32 particles of a warp by a transposition through TMEM (32x32b store, 16x256b load) plus three
// shuffle stages, then added atomically (step-indexed tables) or kept in per-lane register accumulators (U1, W3).
#include <cuda_bf16.h>

#include "common.cuh"
#include "umma.cuh"

namespace cmcd {

constexpr int BT_H = 64;
constexpr int BT_THREADS = 256;
constexpr int BT_PB = 128;                    // particles per tile
constexpr uint32_t BT_A_HI = 0, BT_A_LO = 64, BT_D12 = 128, BT_D3 = 192, BT_TILE_COLS = 256;
constexpr int BT_T16K = 16384;                // one 64x64 fp32 tile, or one [128][64] bf16 tile
constexpr int BT_OFF_BF_HI = 0, BT_OFF_BF_LO = BT_T16K, BT_OFF_BD_HI = 2 * BT_T16K, BT_OFF_BD_LO = 3 * BT_T16K;
constexpr int BT_OFF_TILE = 4 * BT_T16K;      // + wg * 4 * BT_T16K : X_H1, X_H2 (a1), Z_G1, Z_G2 (dp2)
constexpr int BT_OFF_SMALL = 12 * BT_T16K;    // 196608
constexpr int BT_FLUSH_NODES = 4;             // weight-gradient batches (nodes) between flushes of the TMEM gW2 accumulator

struct BtLayout {  // offsets (floats) into one CTA's partial-gradient slice
    int W2rows, c1, c2, c3, U1, W3, os, beta, eps, mu, ls, P;
};

static BtLayout bt_make_layout(int D, int K) {
    BtLayout l;
    int o = 0;
    const int T = K + 1;
    l.W2rows = o; o += 2 * BT_PB * BT_H;
    l.c1 = o; o += T * BT_H;
    l.c2 = o; o += T * BT_H;
    l.c3 = o; o += T * D;
    l.U1 = o; o += D * BT_H;
    l.W3 = o; o += BT_H * D;
    l.os = o; o += 1;
    l.beta = o; o += K;
    l.eps = o; o += K;
    l.mu = o; o += D;
    l.ls = o; o += D;
    l.P = (o + 3) & ~3;
    return l;
}

// A global load the compiler must issue HERE: plain loads / __ldg whose value is first used one trajectory point later get sunk to
// the use and expose a full L2 round trip per point (ncu r2b: 3.6 % of the adjoint's samples on the step-constant loads, 1.2 % on
// the trajectory rows).  volatile asm keeps its position among the other volatile asm statements of the loop body.
__device__ __forceinline__ float bt_ldg_now(const float* p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ float bt_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Branch-free pair exchange for the reduction stages: with m = all-ones in the upper half of the lane group and 0 in the
// lower half, send = m ? a : b and keep = m ? b : a as two 3-input logic ops (LOP3) -- selects written with ?: compile to
// pairs of predicated moves here, which made MOV the most executed opcode of the kernel.
__device__ __forceinline__ void bt_pick(float a, float b, uint32_t m, float& send, float& keep) {
    const uint32_t ua = __float_as_uint(a), ub = __float_as_uint(b);
    const uint32_t us = (ua & m) | (ub & ~m);
    send = __uint_as_float(us);
    keep = __uint_as_float(ua ^ ub ^ us);
}

// Column sums over the 32 particles of a warp through TMEM: every thread stores its 16 values into its own lane
// (32x32b), the warp reads the block back as 16x256b fragments -- thread t then holds 4 lanes x 4 columns -- sums its
// lanes, and three exchange stages over lane bits 4, 3, 2 finish the reduction.  Half the instruction slots of the
// shuffle butterfly.  `scratch` = 48 free TMEM columns of this warp's lane quarter.  Result: the total of element
// bt_red_col(lane) in every lane (lanes differing in bit 2 hold the same element).
__device__ __forceinline__ int bt_red_col(int lane) { return ((lane >> 4) & 1) * 8 + (lane & 3) * 2 + ((lane >> 3) & 1); }
__device__ __forceinline__ void bt_tmem_reduce16x3(uint32_t sc0, uint32_t sc1, uint32_t sc2, const float (&v0)[16], const float (&v1)[16],
                                                   const float (&v2)[16], int lane, float& r0, float& r1, float& r2) {
    const uint32_t sc[3] = {sc0, sc1, sc2};
    {
        uint32_t w[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) w[e] = __float_as_uint(v0[e]);
        umma::tmem_st16(sc0, w);
#pragma unroll
        for (int e = 0; e < 16; ++e) w[e] = __float_as_uint(v1[e]);
        umma::tmem_st16(sc1, w);
#pragma unroll
        for (int e = 0; e < 16; ++e) w[e] = __float_as_uint(v2[e]);
        umma::tmem_st16(sc2, w);
    }
    umma::tmem_st_wait();
    uint32_t a[3][8], b[3][8];
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        umma::tmem_ld_16x256b_x2(sc[s], a[s]);
        umma::tmem_ld_16x256b_x2(sc[s] + (16u << 16), b[s]);
    }
    umma::tmem_ld_wait();
    float p[3][4];
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = (j >> 1) * 4 + (j & 1);   // registers r, r + 2 of both loads hold column 8 (j / 2) + 2 (t % 4) + j % 2
            p[s][j] = (__uint_as_float(a[s][r]) + __uint_as_float(a[s][r + 2])) + (__uint_as_float(b[s][r]) + __uint_as_float(b[s][r + 2]));
        }
    float q[3][2], d[3];
    {
        const uint32_t m = (lane & 16) ? 0xFFFFFFFFu : 0u;
#pragma unroll
        for (int s = 0; s < 3; ++s)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float send, keep;
                bt_pick(p[s][j], p[s][j + 2], m, send, keep);
                q[s][j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
    }
    {
        const uint32_t m = (lane & 8) ? 0xFFFFFFFFu : 0u;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            float send, keep;
            bt_pick(q[s][0], q[s][1], m, send, keep);
            d[s] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
#pragma unroll
    for (int s = 0; s < 3; ++s) d[s] += __shfl_xor_sync(0xffffffffu, d[s], 4);
    r0 = d[0]; r1 = d[1]; r2 = d[2];
}

// 16 values of row q (hidden units 16c..16c+15) -> bf16 hi + bf16 lo parts in two MN-major SW128 tiles
// (h1 / h2: the packed (even, odd) bf16 pairs, also the TMEM A operand of the kind::f16 recompute / input-gradient products)
__device__ __forceinline__ void bt_stage_bf16x2(uint8_t* t1, uint8_t* t2, int q, int c, const float (&v)[16], uint32_t (&h1)[8], uint32_t (&h2)[8]) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const __nv_bfloat162 hi = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
        const float2 f = __bfloat1622float2(hi);
        const __nv_bfloat162 lo = __floats2bfloat162_rn(v[2 * k] - f.x, v[2 * k + 1] - f.y);
        h1[k] = *reinterpret_cast<const uint32_t*>(&hi);
        h2[k] = *reinterpret_cast<const uint32_t*>(&lo);
    }
    const int row = q * 128, sw = q & 7;
    const int o0 = row + (((2 * c) ^ sw) << 4), o1 = row + (((2 * c + 1) ^ sw) << 4);
    *reinterpret_cast<uint4*>(t1 + o0) = make_uint4(h1[0], h1[1], h1[2], h1[3]);
    *reinterpret_cast<uint4*>(t1 + o1) = make_uint4(h1[4], h1[5], h1[6], h1[7]);
    *reinterpret_cast<uint4*>(t2 + o0) = make_uint4(h2[0], h2[1], h2[2], h2[3]);
    *reinterpret_cast<uint4*>(t2 + o1) = make_uint4(h2[4], h2[5], h2[6], h2[7]);
}
__device__ __forceinline__ void bt_stage_bf16x2(uint8_t* t1, uint8_t* t2, int q, int c, const float (&v)[16]) {
    uint32_t h1[8], h2[8];
    bt_stage_bf16x2(t1, t2, q, c, v, h1, h2);
}

// 24 x tcgen05.mma kind::tf32: D = A_hi B_lo + A_lo B_hi + A_hi B_hi (small terms first: the accumulator truncates)
// (dhi / dlo: shared-memory descriptors of the B tile's first K block; one K block further = +256 B = +16 in the address field)
#ifndef BT_GEMM_TF32
// Default since round 2: GEMM1 / GEMM2 as 12 x tcgen05.mma kind::f16 on the bf16 (hi, lo) pairs the weight-gradient staging
// computes anyway -- A from TMEM as packed pairs (column c = elements 2c, 2c+1; 32 + 32 columns), B = W2 as bf16 (hi, lo) K-major
// core-matrix tiles -- D = A_hi B_lo + A_lo B_hi + A_hi B_hi (small terms first).  Operand precision 2^-17 (hi + lo carry 16
// mantissa bits) instead of 2^-22: used only for the adjoint's recompute / cotangent products, whose errors average over
// particles and steps (parity tolerances unchanged, tests/test_gpu_parity_bwd.py); the forward kernel keeps tf32 x 3.
// Saves the tf32 splits (384 instructions per node), half of the operand STTM traffic and 3/4 of the MMA time per product.
// -DBT_GEMM_TF32 restores the tf32 3-pass form.
__device__ __forceinline__ void bt_issue_gemm(uint32_t tmem_base, uint32_t d_col, uint64_t dhi, uint64_t dlo) {
    const uint32_t idesc = umma::make_idesc_bf16_k(128, BT_H);
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
        const uint32_t acol = tmem_base + (pass == 1 ? BT_A_LO : BT_A_HI);
        const uint64_t bd = (pass == 0) ? dlo : dhi;
#pragma unroll
        for (int k = 0; k < 4; ++k)   // K = 16 per MMA = 8 packed TMEM columns, two 8 x 16 B core matrices of B (+256 B)
            umma::mma_f16_ts(tmem_base + d_col, acol + k * 8, bd + (uint64_t)(k * 16), idesc, (pass | k) > 0);
    }
}
#else
__device__ __forceinline__ void bt_issue_gemm(uint32_t tmem_base, uint32_t d_col, uint64_t dhi, uint64_t dlo) {
    const uint32_t idesc = umma::make_idesc_tf32(128, BT_H);
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
        const uint32_t acol = tmem_base + (pass == 1 ? BT_A_LO : BT_A_HI);
        const uint64_t bd = (pass == 0) ? dlo : dhi;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            umma::mma_tf32_ts(tmem_base + d_col, acol + k * 8, bd + (uint64_t)(k * 16), idesc, (pass | k) > 0);
    }
}
#endif
// 16 x tcgen05.mma kind::f16: D3[128 x 64] (+)= [X_H1 ; X_H2]^T (Z_G1 + Z_G2) over K = 128 particles
// (dx / dz: descriptors of the first 16-particle K block of the a1 / dp2 staging tiles; +2048 B per K block, +16 KB for the low part)
__device__ __forceinline__ void bt_issue_wgrad(uint32_t tmem_base, uint64_t dx, uint64_t dz, bool fresh) {
    const uint32_t idesc = umma::make_idesc_bf16_mn(128, BT_H);
#pragma unroll
    for (int part = 0; part < 2; ++part) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            umma::mma_f16_ss(tmem_base + BT_D3, dx + (uint64_t)(k * 128), dz + (uint64_t)(part * (BT_T16K >> 4) + k * 128), idesc,
                             (part | k) > 0 || !fresh);
    }
}

// timing experiments (tools/ab_variants.sh; results are WRONG with any of these defined): which phase sits on the critical path?
#ifdef BT_X_NOG1
#define BT_G1S(c, k) xb[0]
#else
#define BT_G1S(c, k) g1s[c][k]
#endif

template <int D>
__global__ void __launch_bounds__(BT_THREADS, 1) bridge_bwd_tc_kernel(const BridgeArgs a, const float* __restrict__ cot_negw,
                                                                      float* __restrict__ partials, const BtLayout L) {
    constexpr int ACT = ACT_GELU;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t mbars[2][4];   // per tile: GEMM1 done, GEMM2 done, WGRAD done, operands staged (128 arrivals)
    const int tid = threadIdx.x, wg = tid >> 7, q = tid & 127, warp = tid >> 5, lane = tid & 31;
    const NetView& nv = a.net;
    float* sf = reinterpret_cast<float*>(smem + BT_OFF_SMALL);
    float* sU1 = sf;                        // [D][64]
    float* sW3 = sU1 + D * BT_H;            // [D][64]: W3 TRANSPOSED (row m = output dim), so that pairs of hidden units are packed operands
    // step-independent skinny gradients (U1, W3, and -- dds: c2 = st2.b, c3 = out.b are the same row for every step -- c2, c3):
    // one PRIVATE accumulator block per warp, [c2 64 | W3 64 x D | U1 D x 64]: after a reduction the 16 lanes with bit 2 clear hold
    // 16 distinct hidden units, so plain load-add-store suffices (ATOMS on CTA-shared accumulators showed up as 10 % of the
    // kernel's stall samples, profiles/r2_bwd_tc_hot_lines.txt).  Summed over the warps at the end; c2 / c3 go to table row 0.
    constexpr int BT_WACC = BT_H + 2 * BT_H * D;
    float* sAccAll = sW3 + BT_H * D;                          // [8 warps][BT_WACC]
    float* wAcc = sAccAll + warp * BT_WACC;
    float* sAccC3 = sAccAll + (BT_THREADS / 32) * BT_WACC;    // [D] c3, [D] = out_scale  (rare: one atomic per warp and tile)
    float* sTp = sAccC3 + 8;                // mixture parameters (generic layout, MIX_STRIDE floats per component)
    float2* sMu = reinterpret_cast<float2*>(sTp + MIX_MAX * MIX_STRIDE);   // many_gmm: dense component means
    // per-warp double buffer for the per-step table rows c1[t] | c2[t] (cp.async one half-step ahead)
    float* sTab = reinterpret_cast<float*>(sMu + MIX_MAX) + warp * (2 * 2 * BT_H);
    auto stage_tab = [&](int t, int buf) {   // 32 lanes x 16 B = c1 row (256 B) + c2 row (256 B)
        const float* src = (lane < 16 ? nv.c1 : nv.c2) + (size_t)t * BT_H + (lane & 15) * 4;
        umma::cp_async16(sTab + buf * (2 * BT_H) + lane * 4, src);
        umma::cp_async_commit();
    };
    for (int idx = tid; idx < BT_H * BT_H; idx += BT_THREADS) {
        const int i = idx / BT_H, j = idx % BT_H;
#ifndef BT_GEMM_TF32
        const float w = nv.W2[idx];
        const __nv_bfloat16 hi = __float2bfloat16(w), lo = __float2bfloat16(w - __bfloat162float(hi));
        const int of = umma::core_off16(j, i, BT_H);   // GEMM1: B[n = j][k = i] = W2[i][j]
        const int od = umma::core_off16(i, j, BT_H);   // GEMM2: B[n = i][k = j] = W2[i][j]
        *reinterpret_cast<__nv_bfloat16*>(smem + BT_OFF_BF_HI + of) = hi;
        *reinterpret_cast<__nv_bfloat16*>(smem + BT_OFF_BF_LO + of) = lo;
        *reinterpret_cast<__nv_bfloat16*>(smem + BT_OFF_BD_HI + od) = hi;
        *reinterpret_cast<__nv_bfloat16*>(smem + BT_OFF_BD_LO + od) = lo;
#else
        float hi, lo;
        umma::split_tf32(nv.W2[idx], hi, lo);
        const int of = umma::core_off(j, i, BT_H);   // GEMM1: B[n = j][k = i] = W2[i][j]
        const int od = umma::core_off(i, j, BT_H);   // GEMM2: B[n = i][k = j] = W2[i][j]
        *reinterpret_cast<float*>(smem + BT_OFF_BF_HI + of) = hi;
        *reinterpret_cast<float*>(smem + BT_OFF_BF_LO + of) = lo;
        *reinterpret_cast<float*>(smem + BT_OFF_BD_HI + od) = hi;
        *reinterpret_cast<float*>(smem + BT_OFF_BD_LO + od) = lo;
#endif
    }
    for (int i = tid; i < D * BT_H; i += BT_THREADS) sU1[i] = nv.U1[i];
    for (int i = tid; i < BT_H * D; i += BT_THREADS) sW3[(i % D) * BT_H + i / D] = nv.W3[i];
    for (int i = tid; i < (BT_THREADS / 32) * BT_WACC + 8; i += BT_THREADS) sAccAll[i] = 0.f;
    const int ntp = (a.tgt.kind == TGT_GMM || a.tgt.kind == TGT_MANY_GMM) ? a.tgt.ncomp * MIX_STRIDE : 0;
    for (int i = tid; i < ntp; i += BT_THREADS) sTp[i] = a.tgt.mix[i];
    const bool fast_gmm = (a.tgt.kind == TGT_MANY_GMM);
    if (fast_gmm)
        many_gmm_stage_means(a.tgt, sMu, tid, BT_THREADS);
    const ManyGmmConst gc = many_gmm_const(a.tgt);
    if (warp == 0) umma::tmem_alloc(&tmem_slot, 512);
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) umma::mbar_init(&mbars[0][0] + i, (i & 3) == 3 ? BT_PB : 1);
    }
    umma::fence_async_smem();
    umma::fence_before();
    __syncthreads();
    umma::fence_after();

    const uint32_t tmem_base = tmem_slot + (uint32_t)wg * BT_TILE_COLS;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
#ifndef BT_GEMM_TF32
    constexpr uint32_t BT_B_SBO = 16 * BT_H;   // bf16 K-major core matrices: 8 rows x 16 B, next 8-row group 16 * K bytes further
#else
    constexpr uint32_t BT_B_SBO = 32 * BT_H;
#endif
    const uint64_t bf_hi = umma::make_desc(umma::smem_u32(smem + BT_OFF_BF_HI), 128, BT_B_SBO);
    const uint64_t bf_lo = umma::make_desc(umma::smem_u32(smem + BT_OFF_BF_LO), 128, BT_B_SBO);
    const uint64_t bd_hi = umma::make_desc(umma::smem_u32(smem + BT_OFF_BD_HI), 128, BT_B_SBO);
    const uint64_t bd_lo = umma::make_desc(umma::smem_u32(smem + BT_OFF_BD_LO), 128, BT_B_SBO);
    uint8_t* tX1 = smem + BT_OFF_TILE + wg * 4 * BT_T16K;
    uint8_t* tX2 = tX1 + BT_T16K;
    uint8_t* tZ1 = tX1 + 2 * BT_T16K;
    uint8_t* tZ2 = tX1 + 3 * BT_T16K;
    const uint64_t x_desc = umma::make_desc_sw128(umma::smem_u32(tX1), BT_T16K, 1024);
    const uint64_t z_desc = umma::make_desc_sw128(umma::smem_u32(tZ1), BT_T16K, 1024);
    uint64_t* mb1 = &mbars[wg][0];
    uint64_t* mb2 = &mbars[wg][1];
    uint64_t* mb3 = &mbars[wg][2];
    uint64_t* mbR = &mbars[wg][3];
    uint32_t par1 = 0, par2 = 0, par3 = 0, parR = 0;

    float* part = partials + (size_t)blockIdx.x * L.P;
    float* w2row = part + L.W2rows + (size_t)(wg * BT_PB + q) * BT_H;   // this thread's private row of the stacked gW2 tile

    const bool cais = (a.mode == CMCD_MODE_CAIS_SN || a.mode == CMCD_MODE_CAIS_VAR_SN);
    const bool pathwise = a.mode != CMCD_MODE_CAIS_VAR_SN;
    const bool nn_b = (a.mode != CMCD_MODE_ULA);
    const bool nn_f = cais;
    const int K = a.K;
    const float out_scale = net_out_scale(nv);

    float mu[D], sig[D], ivar[D];
#pragma unroll
    for (int j = 0; j < D; ++j) { mu[j] = a.vd_mean[j]; sig[j] = expf(a.vd_logdiag[j]); ivar[j] = 1.0f / (sig[j] * sig[j]); }

    // per-lane accumulators of the skinny gradients that are not indexed by the step: after a butterfly, lane l holds the
    // warp's sum for hidden unit 16 cc + bt_red_col(l); summed over the whole kernel, flushed once at the end
    // (the step-independent skinny gradients -- U1, W3, c2, c3 -- go into shared-memory accumulators of the CTA with one
    //  ATOMS per reduced element and warp: the select chains that kept per-chunk register accumulators statically indexed cost
    //  16 instructions per 16-unit chunk)

    bool wgrad_pending = false;   // a WGRAD batch has been committed to mb3 and not yet waited for
    bool d3_fresh = true;         // the TMEM gW2 accumulator holds nothing (next WGRAD starts with accumulate = 0)
    int d3_steps = 0;

    // flush: gW2 accumulator (this thread's TMEM lane, 64 columns) += into the thread-private global row
    auto flush_d3 = [&]() {
        if (wgrad_pending) { umma::mbar_wait(mb3, par3); par3 ^= 1u; wgrad_pending = false; }
        umma::fence_after();
        if (!d3_fresh) {
            float4* g = reinterpret_cast<float4*>(w2row);
            float4 r[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) r[k] = g[k];      // the whole row in flight at once (L2-resident, written 2 steps ago)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t v[16];
                umma::tmem_ld16(tmem_lane + BT_D3 + c * 16, v);
                umma::tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float4 t = r[c * 4 + k];
                    t.x += __uint_as_float(v[4 * k + 0]); t.y += __uint_as_float(v[4 * k + 1]);
                    t.z += __uint_as_float(v[4 * k + 2]); t.w += __uint_as_float(v[4 * k + 3]);
                    g[c * 4 + k] = t;
                }
            }
        }
        d3_fresh = true;
        d3_steps = 0;
        umma::fence_before();
    };

    const long long ntiles = (a.N + BT_PB - 1) / BT_PB;
    for (long long tile = (long long)wg * gridDim.x + blockIdx.x; tile < ntiles; tile += 2LL * gridDim.x) {   // few tiles: one per CTA
        const long long n_raw = tile * BT_PB + q;
        const bool active = n_raw < a.N;
        const long long n = active ? n_raw : a.N - 1;   // tail lanes shadow the last particle with zero cotangent
        const float c = active ? -cot_negw[n] : 0.f;    // dL/dw_n
        // Node form: K + 1 nodes z_K .. z_0, ONE network evaluation / VJP / weight-gradient batch per node.  In the CAIS
        // modes NN(z_j, j) is used twice by the reference -- backward-kernel mean of step j-1 (mcd_cais.py:78) and
        // forward-kernel mean of step j (mcd_cais.py:60) -- and the VJP is linear in the output cotangent, so both uses
        // share one recompute and one pull-back of  v = eps_{j-1} G_B - eps_j G_F.  Likewise one score / Hessian of the
        // target per node serves both kernel means (beta differs, the point does not).
        //   B use (j > 0, step iB = j-1): mean_b = x - eps_B u_B + eps_B nn ; r = (z_{j-1} - mean_b) / (2 eps_B) ; G_B = c r
        //   F use (j < K, step iF = j)  : mean_f = x - eps_F u_F - eps_F nn ; xs = (z_{j+1} - mean_f) / (2 eps_F) ;
        //                                 G_F = abar_j (pathwise: carried from node j+1) or -c xs (log-var mode)
        //   carry_{j-1} = [-c r_j + G_F (1 - eps_F (1-beta_F) mk_q / sigma^2)] + [G_B (1 - eps_B (1-beta_B) mk_q / sigma^2)]
        //                 + H_p(x) mk_t (beta_F eps_F G_F + beta_B eps_B G_B) + J_x^T v          (node K: first bracket = c grad log p)
        float x[D], zup[D], zprev[D], zpre2[D], carry[D], rS[D], gmu[D], gls[D], zero[D], hv[D], sx[D];
        float hx[3] = {0.f, 0.f, 0.f};   // many_gmm fast path: Hessian of log p at x (h00, h01, h11)
#pragma unroll
        for (int j = 0; j < D; ++j) {
            x[j] = a.traj[((size_t)K * D + j) * a.N + n];
            zprev[j] = a.traj[((size_t)(K - 1) * D + j) * a.N + n];
            zpre2[j] = (K > 1) ? a.traj[((size_t)(K - 2) * D + j) * a.N + n] : 0.f;
            gmu[j] = 0.f; gls[j] = 0.f; zero[j] = 0.f; carry[j] = 0.f; rS[j] = 0.f; zup[j] = 0.f; sx[j] = 0.f; hv[j] = 0.f;
        }
        f32x2_t g1s[4][8];              // act'(pre1) of the current evaluation as packed pairs, chunk-indexed with static indices only
        float cgb = 0.f, cge = 0.f;     // beta_i / eps_i gradient of step j: B-use part, carried from node j+1 to node j
        float c3Acc[D], gosAcc = 0.f;   // cotangents of the (step-independent) output bias and of out_scale, summed over the nodes
#pragma unroll
        for (int j = 0; j < D; ++j) c3Acc[j] = 0.f;

        const int t0 = cais ? 0 : -1;   // table row of node j: t0 + j
        stage_tab(t0 + K, (t0 + K) & 1);
        float bnext = __ldg(a.betas + K - 1), enext = __ldg(a.eps + K - 1);   // step K-1 (B use of node K)
        float bprev = 0.f, eprev = 0.f;                                       // step j (F use): none at node K
        for (int j = K; j >= 0; --j) {
            const bool hasB = j > 0, hasF = j < K;
            const int t = t0 + j;
            f32x2_t xb[D];                    // (x_d, x_d): broadcast operands of the packed fp32x2 products
            float c3v[D];                     // output-bias row of this node, requested now, used at the end of epilogue 1
#pragma unroll
            for (int d = 0; d < D; ++d) { xb[d] = pk2(x[d], x[d]); c3v[d] = (cais || j > 0) ? bt_ldg_now(nv.c3 + (size_t)(cais ? j : j - 1) * D + d) : 0.f; }
            const bool use_nn = cais || hasB;
            // step constants of both uses; an absent use gets eps = 0, c = 0 so that all of its terms vanish
            // step constants: (beta, eps) of step j-1 were requested one node ago; those of step j are last node's B-use values
            const float bB = hasB ? bnext : 0.f, eB = hasB ? enext : 0.f;
            const float bF = bprev, eF = eprev;
            if (j > 1) { bnext = bt_ldg_now(a.betas + j - 2); enext = bt_ldg_now(a.eps + j - 2); }
            const float tsB = hasB ? 2.0f * eB : 1.f, tsF = hasF ? 2.0f * eF : 1.f;
            const float ombB = 1.0f - bB, ombF = 1.0f - bF;
            const float cB = hasB ? c : 0.f, cF = hasF ? c : 0.f;
            const float eFn = nn_f ? eF : 0.f;   // MCD_ULA_sn: the forward-kernel mean has no network term
            const float* tab = sTab + (t & 1) * (2 * BT_H);
            if (use_nn) {
                // table rows of this node were requested one node ago; request the next ones
                umma::cp_async_wait_all();
                __syncwarp();
                if (j > 0 && (cais || j > 1)) stage_tab(t - 1, (t - 1) & 1);
            }

            // ---------------- layer 1, A operand, a1 staging; GEMM1 ----------------
            if (use_nn) {
                if (d3_steps >= BT_FLUSH_NODES) flush_d3();
                if (wgrad_pending) { umma::mbar_wait(mb3, par3); par3 ^= 1u; wgrad_pending = false; }   // a1 / dp2 staging tiles are free again
                const float4* __restrict__ c1v = reinterpret_cast<const float4*>(tab);
#pragma unroll 1
                for (int cc = 0; cc < 4; ++cc) {
                    float a1[16];
                    f32x2_t g1[8];
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {
                        const float4 cv = c1v[cc * 4 + qq];
                        f32x2_t p01 = pk2(cv.x, cv.y), p23 = pk2(cv.z, cv.w);   // two hidden units per FFMA2 (x broadcast pairs built once per node)
#pragma unroll
                        for (int d = 0; d < D; ++d) {
                            const float4 u = *reinterpret_cast<const float4*>(sU1 + d * BT_H + cc * 16 + qq * 4);
                            p01 = fma2(xb[d], pk2(u.x, u.y), p01);
                            p23 = fma2(xb[d], pk2(u.z, u.w), p23);
                        }
                        {   // two activations (+ derivatives) per instruction slot (FFMA2)
                            float q0, q1, q2, q3;
                            upk2(p01, q0, q1); upk2(p23, q2, q3);
                            f32x2_t A, DA;
                            gelu_fast_grad2(q0, q1, A, DA);
                            upk2(A, a1[qq * 4 + 0], a1[qq * 4 + 1]); g1[qq * 2 + 0] = DA;
                            gelu_fast_grad2(q2, q3, A, DA);
                            upk2(A, a1[qq * 4 + 2], a1[qq * 4 + 3]); g1[qq * 2 + 1] = DA;
                        }
                    }
#ifndef BT_X_NOG1
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {   // predicated moves: static register indices, no control-flow merge
#pragma unroll
                        for (int e = 0; e < 8; ++e) g1s[k4][e] = (k4 == cc) ? g1[e] : g1s[k4][e];
                    }
#endif
#ifndef BT_GEMM_TF32
                    {
                        uint32_t h1[8], h2[8];
                        bt_stage_bf16x2(tX1, tX2, q, cc, a1, h1, h2);
                        umma::tmem_st8(tmem_lane + BT_A_HI + cc * 8, h1);
                        umma::tmem_st8(tmem_lane + BT_A_LO + cc * 8, h2);
                    }
#else
#ifndef BT_X_NOSPLIT
                    uint32_t hh[16], ll[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        float hi, lo;
                        umma::split_tf32(a1[e], hi, lo);
                        hh[e] = __float_as_uint(hi); ll[e] = __float_as_uint(lo);
                    }
                    umma::tmem_st16(tmem_lane + BT_A_HI + cc * 16, hh);
                    umma::tmem_st16(tmem_lane + BT_A_LO + cc * 16, ll);
#endif
#ifndef BT_X_NOSTAGE
                    bt_stage_bf16x2(tX1, tX2, q, cc, a1);
#else
                    if (a1[3] == 12345.678f) bt_stage_bf16x2(tX1, tX2, q, cc, a1);
#endif
#endif
                }
                umma::tmem_st_wait();
                umma::fence_before();
                umma::fence_async_smem();
                umma::mbar_arrive(mbR);     // only the issuing thread waits for the tile's 128 arrivals; the rest moves on
                if (q < 32) {   // first warp of the tile (converged): wait for the 128 arrivals, one elected lane issues
                    umma::mbar_wait(mbR, parR); parR ^= 1u;
                    umma::fence_after();
                    if (umma::elect_one()) {
                        bt_issue_gemm(tmem_base, BT_D12, bf_hi, bf_lo);
                        umma::commit(mb1);
                    }
                    __syncwarp();
                }
            }

            // ---------------- independent of the network: target score (+ Hessian) and score-side terms at x ----------------
#ifdef BT_X_NOSCORE
            if (fast_gmm) { sx[0] = -x[0]; sx[1] = -x[1]; hx[0] = -1.f; hx[1] = 0.f; hx[2] = -1.f; }
#else
            if (fast_gmm) many_gmm_eval_hess(gc, sMu, x[0], x[1], sx[0], sx[1], hx[0], hx[1], hx[2]);   // every HVP at x comes from it
#endif
            else target_eval<D, false>(a.tgt, sTp, x, sx, zero, hv);
            float sq[D], mk_t[D], mk_q[D], uB[D], uF[D], mB[D], mF[D], dc[D], nn[D], dx[D];
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
                mB[d] = x[d] - eB * uB[d];
                mF[d] = x[d] - eF * uF[d];
                nn[d] = 0.f; dx[d] = 0.f;
            }

            // ---------------- epilogue 1: a2, act'(pre2), raw network output ----------------
            float o[D];
#pragma unroll
            for (int m = 0; m < D; ++m) o[m] = 0.f;
            if (use_nn) {
#pragma unroll
                for (int m = 0; m < D; ++m) o[m] = c3v[m];
                const float4* __restrict__ c2v = reinterpret_cast<const float4*>(tab + BT_H);
                umma::mbar_wait(mb1, par1); par1 ^= 1u;
                umma::fence_after();
                f32x2_t O2[D];                // output-layer partial sums over (even, odd) hidden units
#pragma unroll
                for (int m = 0; m < D; ++m) O2[m] = pk2(0.f, 0.f);
#pragma unroll 1
                for (int cc = 0; cc < 4; ++cc) {
                    uint32_t v[16], a2u[16], g2u[16];
                    umma::tmem_ld16(tmem_lane + BT_D12 + cc * 16, v);
                    umma::tmem_ld_wait();
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {
                        const float4 cv = c2v[cc * 4 + qq];
                        const f32x2_t P01 = add2(pk2(__uint_as_float(v[qq * 4 + 0]), __uint_as_float(v[qq * 4 + 1])), pk2(cv.x, cv.y));
                        const f32x2_t P23 = add2(pk2(__uint_as_float(v[qq * 4 + 2]), __uint_as_float(v[qq * 4 + 3])), pk2(cv.z, cv.w));
                        float q0, q1, q2, q3;
                        upk2(P01, q0, q1); upk2(P23, q2, q3);
                        f32x2_t A01, D01, A23, D23;
                        gelu_fast_grad2(q0, q1, A01, D01);
                        gelu_fast_grad2(q2, q3, A23, D23);
#pragma unroll
                        for (int m = 0; m < D; ++m) {
                            const float4 w = *reinterpret_cast<const float4*>(sW3 + m * BT_H + cc * 16 + qq * 4);
                            O2[m] = fma2(A01, pk2(w.x, w.y), O2[m]);
                            O2[m] = fma2(A23, pk2(w.z, w.w), O2[m]);
                        }
                        float f0, f1;
                        upk2(A01, f0, f1); a2u[qq * 4 + 0] = __float_as_uint(f0); a2u[qq * 4 + 1] = __float_as_uint(f1);
                        upk2(A23, f0, f1); a2u[qq * 4 + 2] = __float_as_uint(f0); a2u[qq * 4 + 3] = __float_as_uint(f1);
                        upk2(D01, f0, f1); g2u[qq * 4 + 0] = __float_as_uint(f0); g2u[qq * 4 + 1] = __float_as_uint(f1);
                        upk2(D23, f0, f1); g2u[qq * 4 + 2] = __float_as_uint(f0); g2u[qq * 4 + 3] = __float_as_uint(f1);
                    }
                    // the A region is dead between GEMM1 and GEMM2: park a2 / act'(pre2) in this thread's lane
                    umma::tmem_st16(tmem_lane + BT_A_HI + cc * 16, a2u);
                    umma::tmem_st16(tmem_lane + BT_A_LO + cc * 16, g2u);
                }
#pragma unroll
                for (int m = 0; m < D; ++m) {
                    float e0, e1;
                    upk2(O2[m], e0, e1);
                    o[m] += e0 + e1;
                }
                umma::tmem_st_wait();
#pragma unroll
                for (int m = 0; m < D; ++m) nn[m] = out_scale * fminf(fmaxf(o[m], -nv.out_clip), nv.out_clip);
            }

            // ---------------- kernel means, residuals, cotangents on the two means ----------------
            float GB[D], GF[D], rB[D], xs[D], vv[D], wq[D];
            float rr = 0.f, xx = 0.f;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const float meanB = mB[d] + eB * nn[d];
                const float meanF = mF[d] - eFn * nn[d];
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

            // ---------------- output-layer VJP, dp2, GEMM2 + WGRAD ----------------
            if (use_nn) {
                float vo[D];
#pragma unroll
                for (int m = 0; m < D; ++m) {
                    const float v = vv[m];
                    const float oc = fminf(fmaxf(o[m], -nv.out_clip), nv.out_clip);
                    gosAcc = fmaf(v, oc, gosAcc);
                    vo[m] = (fabsf(o[m]) <= nv.out_clip) ? v * out_scale : 0.f;
                    c3Acc[m] += vo[m];
                }
                f32x2_t vob[D];
#pragma unroll
                for (int m = 0; m < D; ++m) vob[m] = pk2(vo[m], vo[m]);
#pragma unroll 1
                for (int cc = 0; cc < 4; ++cc) {
                    uint32_t a2u[16], g2u[16];
                    umma::tmem_ld16(tmem_lane + BT_A_HI + cc * 16, a2u);
                    umma::tmem_ld16(tmem_lane + BT_A_LO + cc * 16, g2u);
                    umma::tmem_ld_wait();
                    float dp2[16];
                    uint32_t hh[16], ll[16];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {     // dp2 = (W3 vo) * act'(pre2), two hidden units per packed instruction
                        f32x2_t S = mul2(*reinterpret_cast<const f32x2_t*>(sW3 + cc * 16 + 2 * k), vob[0]);
#pragma unroll
                        for (int m = 1; m < D; ++m) S = fma2(*reinterpret_cast<const f32x2_t*>(sW3 + m * BT_H + cc * 16 + 2 * k), vob[m], S);
                        const f32x2_t DP = mul2(S, pk2(__uint_as_float(g2u[2 * k]), __uint_as_float(g2u[2 * k + 1])));
                        upk2(DP, dp2[2 * k], dp2[2 * k + 1]);
                    }
#ifndef BT_GEMM_TF32
                    {   // packed pairs of chunk cc go to columns [8 cc, 8 cc + 8) of the two A regions: inside parking chunks <= cc / 2,
                        // which this loop has already consumed
                        uint32_t h1[8], h2[8];
                        bt_stage_bf16x2(tZ1, tZ2, q, cc, dp2, h1, h2);
                        umma::tmem_st8(tmem_lane + BT_A_HI + cc * 8, h1);
                        umma::tmem_st8(tmem_lane + BT_A_LO + cc * 8, h2);
                    }
#else
#ifndef BT_X_NOSPLIT
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        float hi, lo;
                        umma::split_tf32(dp2[e], hi, lo);
                        hh[e] = __float_as_uint(hi); ll[e] = __float_as_uint(lo);
                    }
                    umma::tmem_st16(tmem_lane + BT_A_HI + cc * 16, hh);
                    umma::tmem_st16(tmem_lane + BT_A_LO + cc * 16, ll);
#endif
#ifndef BT_X_NOSTAGE
                    bt_stage_bf16x2(tZ1, tZ2, q, cc, dp2);
#else
                    if (dp2[3] == 12345.678f) bt_stage_bf16x2(tZ1, tZ2, q, cc, dp2);
#endif
#endif
                    // gc2[t][j] += sum_p dp2 ; gW3[j][m] += sum_p a2 vo[m]   (three butterflies in lock step; D == 2)
                    {
                        static_assert(D == 2, "the interleaved reductions assume d = 2");
                        float t0v[16], t1v[16];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const f32x2_t A2 = pk2(__uint_as_float(a2u[2 * k]), __uint_as_float(a2u[2 * k + 1]));
                            upk2(mul2(A2, vob[0]), t0v[2 * k], t0v[2 * k + 1]);
                            upk2(mul2(A2, vob[1]), t1v[2 * k], t1v[2 * k + 1]);
                        }
                        float s2, s30, s31;
#ifdef BT_X_NOREDUCE
                        s2 = dp2[0]; s30 = t0v[1]; s31 = t1v[2];
#else
                        bt_tmem_reduce16x3(tmem_lane + BT_D12, tmem_lane + BT_D12 + 16, tmem_lane + BT_D12 + 32, dp2, t0v, t1v, lane, s2, s30, s31);
#endif   // D is free between GEMM1's epilogue and GEMM2
                        if (!(lane & 4)) {
                            const int jc = cc * 16 + bt_red_col(lane);
                            wAcc[jc] += s2;
                            wAcc[BT_H + jc * D + 0] += s30;
                            wAcc[BT_H + jc * D + 1] += s31;
                        }
                    }
                }
                umma::tmem_st_wait();
                umma::fence_before();
                umma::fence_async_smem();
                umma::mbar_arrive(mbR);
                if (q < 32) {
                    umma::mbar_wait(mbR, parR); parR ^= 1u;
                    umma::fence_after();
                    if (umma::elect_one()) {
                        bt_issue_gemm(tmem_base, BT_D12, bd_hi, bd_lo);
                        umma::commit(mb2);
                        bt_issue_wgrad(tmem_base, x_desc, z_desc, d3_fresh);
                        umma::commit(mb3);
                    }
                    __syncwarp();
                }
                wgrad_pending = true;
                d3_fresh = false;
                ++d3_steps;
            }

            // ---------------- independent of the network: target Hessian-vector product at x, scalar gradients ----------------
            if (pathwise) {
                float vm[D], dummy[D];
#pragma unroll
                for (int d = 0; d < D; ++d) vm[d] = mk_t[d] * (bB * eB * GB[d] + bF * eF * GF[d]);
                if (fast_gmm) {
                    hv[0] = fmaf(hx[0], vm[0], hx[1] * vm[1]);
                    hv[1] = fmaf(hx[1], vm[0], hx[2] * vm[1]);
                } else target_eval<D, true>(a.tgt, sTp, x, dummy, vm, hv);
            }
            {
                float gb = cgb, ge = cge;      // step j: B-use part from node j+1, F-use part here
                float ngb = 0.f, nge = cB * rr;   // step j-1: B-use part, completed at node j-1
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
                    const float gbs = bt_warp_sum(gb), ges = bt_warp_sum(ge);
                    if (lane == 0) { atomicAdd(part + L.beta + j, gbs); atomicAdd(part + L.eps + j, ges); }
                }
                cgb = ngb; cge = nge;
            }

            // ---------------- epilogue 2: dp1 = da1 * act'(pre1), dx = U1 dp1, layer-1 gradients ----------------
            if (use_nn) {
                umma::mbar_wait(mb2, par2); par2 ^= 1u;
                umma::fence_after();
                f32x2_t DX2[D];               // input cotangent partial sums over (even, odd) hidden units
#pragma unroll
                for (int d = 0; d < D; ++d) DX2[d] = pk2(0.f, 0.f);
#pragma unroll 1
                for (int cc = 0; cc < 4; ++cc) {
                    uint32_t v[16];
                    umma::tmem_ld16(tmem_lane + BT_D12 + cc * 16, v);
                    umma::tmem_ld_wait();
                    float dp1[16];
                    f32x2_t DP1[8];
                    switch (cc) {
                        case 0:
#pragma unroll
                            for (int k = 0; k < 8; ++k) DP1[k] = mul2(pk2(__uint_as_float(v[2 * k]), __uint_as_float(v[2 * k + 1])), BT_G1S(0, k));
                            break;
                        case 1:
#pragma unroll
                            for (int k = 0; k < 8; ++k) DP1[k] = mul2(pk2(__uint_as_float(v[2 * k]), __uint_as_float(v[2 * k + 1])), BT_G1S(1, k));
                            break;
                        case 2:
#pragma unroll
                            for (int k = 0; k < 8; ++k) DP1[k] = mul2(pk2(__uint_as_float(v[2 * k]), __uint_as_float(v[2 * k + 1])), BT_G1S(2, k));
                            break;
                        default:
#pragma unroll
                            for (int k = 0; k < 8; ++k) DP1[k] = mul2(pk2(__uint_as_float(v[2 * k]), __uint_as_float(v[2 * k + 1])), BT_G1S(3, k));
                            break;
                    }
                    float t0v[16], t1v[16];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
#pragma unroll
                        for (int d = 0; d < D; ++d)
                            DX2[d] = fma2(*reinterpret_cast<const f32x2_t*>(sU1 + d * BT_H + cc * 16 + 2 * k), DP1[k], DX2[d]);
                        upk2(DP1[k], dp1[2 * k], dp1[2 * k + 1]);
                        upk2(mul2(DP1[k], xb[0]), t0v[2 * k], t0v[2 * k + 1]);
                        upk2(mul2(DP1[k], xb[1]), t1v[2 * k], t1v[2 * k + 1]);
                    }
                    {
                        float s1, s40, s41;
#ifdef BT_X_NOREDUCE
                        s1 = dp1[0]; s40 = t0v[1]; s41 = t1v[2];
#else
                        bt_tmem_reduce16x3(tmem_lane + BT_A_HI, tmem_lane + BT_A_HI + 16, tmem_lane + BT_A_HI + 32, dp1, t0v, t1v, lane, s1, s40, s41);
#endif   // the A region is dead after GEMM2
                        if (!(lane & 4)) {
                            const int jc = cc * 16 + bt_red_col(lane);
                            atomicAdd(part + L.c1 + (size_t)t * BT_H + jc, s1);
                            wAcc[BT_H + BT_H * D + 0 * BT_H + jc] += s40;
                            wAcc[BT_H + BT_H * D + 1 * BT_H + jc] += s41;
                        }
                    }
                }
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    float e0, e1;
                    upk2(DX2[d], e0, e1);
                    dx[d] = e0 + e1;
                }
                umma::fence_before();
            }

            // ---------------- combine: cotangent carried to node j-1 (abar_{j-1}); at node 0 it is dL/dz_0 ----------------
            if (pathwise) {
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    const float fpart = hasF ? (GF[d] - cF * rS[d]) : c * sx[d];   // node K: terminal w += log p(z_K) (mcdboundingmachine.py:178)
                    carry[d] = fpart + GB[d] - wq[d] * ivar[d] * mk_q[d] + hv[d] + dx[d];
                }
            }
            bprev = bB; eprev = eB;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                rS[d] = rB[d];
                zup[d] = x[d]; x[d] = zprev[d]; zprev[d] = zpre2[d];   // trajectory rows are loaded two nodes ahead (HBM latency off the critical path)
                if (j > 2) zpre2[d] = bt_ldg_now(a.traj + ((size_t)(j - 3) * D + d) * a.N + n);
            }
        }
        // initial: z0 = mu + sigma xi0, w0 = -log q(z0) = 0.5|xi0|^2 + sum log(sqrt(2pi) sigma)   (zup = z_0 after the last shift)
#pragma unroll
        for (int j = 0; j < D; ++j) {
            if (pathwise) { gmu[j] += carry[j]; gls[j] += carry[j] * (zup[j] - mu[j]); }
            gls[j] += c;
            const float m1 = bt_warp_sum(gmu[j]), m2 = bt_warp_sum(gls[j]), m3 = bt_warp_sum(c3Acc[j]);
            if (lane == 0) { atomicAdd(part + L.mu + j, m1); atomicAdd(part + L.ls + j, m2); atomicAdd(sAccC3 + j, m3); }
        }
        gosAcc = bt_warp_sum(gosAcc);
        if (lane == 0 && gosAcc != 0.f) atomicAdd(sAccC3 + D, gosAcc);
    }
    flush_d3();
    __syncthreads();
    for (int i = tid; i < BT_WACC; i += BT_THREADS) {
        float v = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < BT_THREADS / 32; ++w8) v += sAccAll[w8 * BT_WACC + i];
        if (i < BT_H) part[L.c2 + i] = v;                                            // row 0 of the c2 table carries the sum over the steps
        else if (i < BT_H + BT_H * D) part[L.W3 + (i - BT_H)] = v;                   // [64][D]
        else part[L.U1 + (i - BT_H - BT_H * D)] = v;                                 // [D][64]
    }
    if (tid < D) part[L.c3 + tid] = sAccC3[tid];                                     // likewise c3
    if (tid == D) part[L.os] = sAccC3[D];
    umma::fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem_slot, 512);
}

// out[k] = sum over CTAs of the partial slices; gW2[i][j] = sum of the stacked rows i and 64 + i of both tiles
struct BtOut {
    float *W2, *U1, *W3, *c1, *c2, *c3, *os, *beta, *eps, *mu, *ls;
};

__global__ void bwd_tc_reduce_kernel(const float* __restrict__ partials, int nblocks, BtLayout L, BtOut o, int D, int K) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int nW2 = BT_H * BT_H;
    if (k < nW2) {
        float s = 0.f;
        for (int b = 0; b < nblocks; ++b) {
            const float* p = partials + (size_t)b * L.P + L.W2rows;
            s += (p[k] + p[nW2 + k]) + (p[2 * nW2 + k] + p[3 * nW2 + k]);   // tile 0 rows i, 64+i; tile 1 rows i, 64+i
        }
        if (o.W2) o.W2[k] = s;
        return;
    }
    const int kk = L.c1 + (k - nW2);
    if (kk >= L.P) return;
    float s = 0.f;
    for (int b = 0; b < nblocks; ++b) s += partials[(size_t)b * L.P + kk];
    auto put = [&](float* dst, int off, int len) { if (dst && kk >= off && kk < off + len) dst[kk - off] = s; };
    put(o.c1, L.c1, L.c2 - L.c1); put(o.c2, L.c2, L.c3 - L.c2); put(o.c3, L.c3, L.U1 - L.c3);
    put(o.U1, L.U1, L.W3 - L.U1); put(o.W3, L.W3, L.os - L.W3); put(o.os, L.os, 1);
    put(o.beta, L.beta, K); put(o.eps, L.eps, K); put(o.mu, L.mu, D); put(o.ls, L.ls, D);
}

static size_t bt_smem_bytes(int D) {
    return (size_t)BT_OFF_SMALL + (size_t)(2 * D * BT_H + 8 * (BT_H + 2 * BT_H * D) + 8 + MIX_MAX * MIX_STRIDE + 2 * MIX_MAX + 8 * 4 * BT_H + 8) * sizeof(float);
}

bool bwd_tc_supported(const BridgeArgs& a, int D) {
    return a.net.arch == CMCD_ARCH_DDS && a.net.HP == BT_H && a.K >= 1 && a.mode != CMCD_MODE_ULA && D == 2;
}

size_t bridge_bwd_tc_workspace_bytes(int D, int K, int num_sms) {
    return (size_t)num_sms * bt_make_layout(D, K).P * sizeof(float);
}

int launch_bridge_bwd_tc(const BridgeArgs& a, int D, cudaStream_t st, int num_sms, const float* cot_negw,
                         float* g_vd_mean, float* g_vd_logdiag, float* g_betas, float* g_eps,
                         const cmcd_net_grad* g, void* ws, size_t ws_bytes) {
    if (D != 2) { set_error("bridge_bwd_tc: dim=%d has no instantiation", D); return 2; }
    const BtLayout L = bt_make_layout(D, a.K);
    const long long ntiles = (a.N + BT_PB - 1) / BT_PB;
    long long grid = ntiles;   // up to num_sms tiles: one tile per CTA (the second tile slot idles, the tile has the SM to itself)
    if (grid > num_sms) grid = num_sms;
    if (grid < 1) grid = 1;
    const size_t need = (size_t)grid * L.P * sizeof(float);
    if (!ws || ws_bytes < need) { set_error("bridge_bwd_tc: workspace too small (%zu < %zu)", ws_bytes, need); return 2; }
    CMCD_CUDA_OK(cudaMemsetAsync(ws, 0, need, st));
    const size_t smem = bt_smem_bytes(D);
    auto kern = bridge_bwd_tc_kernel<2>;
    CMCD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)grid, BT_THREADS, smem, st>>>(a, cot_negw, (float*)ws, L);
    CMCD_CUDA_OK(cudaGetLastError());
    BtOut o{};
    if (g) { o.W2 = g->W2; o.U1 = g->U1; o.W3 = g->W3; o.c1 = g->c1; o.c2 = g->c2; o.c3 = g->c3; o.os = g->out_scale; }
    o.beta = g_betas; o.eps = g_eps; o.mu = g_vd_mean; o.ls = g_vd_logdiag;
    const int nout = BT_H * BT_H + (L.P - L.c1);
    bwd_tc_reduce_kernel<<<(nout + 255) / 256, 256, 0, st>>>((const float*)ws, (int)grid, L, o, D, a.K);
    CMCD_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace cmcd
