// Forward bridge kernel, tensor-core variant with FOUR threads per particle, for the two cases in which one 128-particle tile has an
// SM to itself:
//   HT = 144  the WIDE geffner net of README.md:30,34 (emb_dim 130 + d = 2 -> hidden_pad 136; any hidden_pad in 129..144): layer 2 as
//             tcgen05 tiles of 128 particles x 144 x 144 -- 203 KB of operand tiles (W2^T as tf32 hi, tf32 lo and bf16 hi) and 360 of the
//             512 TMEM columns, so no second tile fits;
//   HT = 64   hidden_pad 64 (dds, narrow geffner nets) with at most one tile per SM (README sizes, small shards of a sharded batch):
//             the three-CTAs-per-SM kernel of bridge_fwd_tc.cu then has nothing to overlap.
//
// Same contract as bridge_fwd_kernel (bridge_fwd.cu) / bridge_fwd_tc_kernel (bridge_fwd_tc.cu): replaces
// vmap(compute_log_elbo) (src/mcdboundingmachine.py:126-205) over src/mcd_cais.py:46-89 / src/mcd_cais_var.py:56-101 /
// src/mcd_over_orig.py:18-55 with apply_fun_sn = the geffner net (src/nn.py:42-72) or PISNet (src/nn_dds.py:145-164) in table form
// (include/cmcd_b200.h, cmcd_net); also serves cmcd_bridge_evolve (a.z0 / a.keys given).
//
// What differs from bridge_fwd_tc.cu, and why:
//   * with one tile per SM a lone warp per scheduler walks the activation chains at ~0.25 IPC.  The CTA therefore has 512 threads:
//     group g (warps 4g .. 4g+3) owns the 16-unit chunks {0,1} / {2,3} / {4,5} / {6,7,8} of the 144 hidden units (HT = 64: chunk g).
//     Warps w, w+4, w+8, w+12 address the same TMEM lane quarter, so all four threads of a particle store into / read from its lane.
//     The per-particle algebra is spread too -- on one thread it was the critical path once the network was split: group 0 keeps the
//     particle's state (z, kernel means, log-weight), group 1 runs the key chain and the Gaussians (independent of the trajectory),
//     group 2 evaluates the target score at every node.  The four exchange the network input (group 0 -> others), the partial
//     output sums (others -> group 0, added in a fixed order), xi and (score, log p) through shared memory with two 128-thread
//     named barriers per node.
//   * HT = 144: the MMA batch (45 instructions per half of N, ~3.4k cycles of tensor pipe) would be exposed between layer 1 and the
//     epilogue, so it is issued in two K stages (stage 0 as soon as every group has stored its first chunk(s): {0,2,4,6,7}; stage 1:
//     {1,3,5,8}) and committed per half of N (columns 0..63 for groups 0-1 first, 64..143 for groups 2-3).  HT = 64: one stage of 20.
//   * the per-step table rows arrive by TMA bulk copy into a CTA-wide double buffer.
// Precision scheme as in bridge_fwd_tc.cu: D = A_hi B_lo + A_lo B_hi + A_hi B_hi with A_hi / B_hi / B_lo tf32 and A_lo bf16.
#include <cuda_bf16.h>

#include <cstdlib>

#include "common.cuh"
#include "umma.cuh"

namespace cmcd {

constexpr int TW_PB = 128;        // particles per CTA (TMEM lanes)
constexpr int TW_G = 4;           // threads per particle (groups of 128 threads)
constexpr int TW_THREADS = TW_G * TW_PB;
// Tile widths HT (the MMA's N and K): 144 = hidden_pad 129..144 (the wide geffner net), 64 = hidden_pad 64 at FEW particles (one tile
// per SM at most: the README-size runs and the small shards of the strong-scaling sweep, where a pass is 257 x one node's latency).
template <int HT>
struct TwGeo {
    static constexpr uint32_t COL_D = 0, COL_AH = HT, COL_AL = 2 * HT, COLS = HT == 144 ? 512 : 256;
    static constexpr int B_BYTES = HT * HT * 4, B16_BYTES = HT * HT * 2;
    static constexpr int NSTAGE = HT == 144 ? 2 : 1;   // K stages of the MMA batch
    static constexpr int NHALF = HT == 144 ? 2 : 1;    // halves of N committed separately
    static constexpr int NA = HT == 144 ? 64 : HT;     // columns of the first half (groups 0, 1); second half: 80 (groups 2, 3)
    // 16-unit chunks of group g.  144: [2g, 2g + 2) (group 3: [6, 9)), stage 0 = the first chunk (group 3: the first two);
    // 64: chunk g, one stage.
    static __device__ __forceinline__ int chunk_begin(int g) { return HT == 144 ? 2 * g : g; }
    static __device__ __forceinline__ int chunk_split(int g) { return HT == 144 ? (g == 3 ? 8 : 2 * g + 1) : g + 1; }
    static __device__ __forceinline__ int chunk_end(int g) { return HT == 144 ? (g == 3 ? 9 : 2 * g + 2) : g + 1; }
    static __device__ __forceinline__ int done_bar(int g) { return HT == 144 ? g >> 1 : 0; }
};

template <int D>
struct TwCtx {
    const float *sU1, *sU2, *sW3T, *sU3;  // shared memory, rows padded to HT with zeros; sW3T = W3 transposed, [D][HT]
    const float *c1, *c2, *c3;            // global per-step tables [T][HP], [T][HP], [T][D]
    const float* tab;                     // staged rows c1[t] | c2[t] of the current node (shared memory)
    float out_scale, out_clip;
    int HP, grp;                          // group 0..3 of this thread
    uint32_t tmem_base, tmem_lane;
    uint64_t bhi, blo, bhi16;             // shared-memory descriptors of the B tiles (row n = 0, first K block)
    uint64_t *mbar_ready, *mbar_done;     // [2] A-operand stage complete (512 arrivals); [2] MMA half of N complete
    uint32_t parity;                      // every barrier above completes once per network evaluation
    float* tab_shared;                    // [2][2 * HT]
    uint64_t* tab_bar;                    // [2]
    uint32_t tab_parity;
};

template <int D>
__device__ __forceinline__ float tw_gauss_logprob(const float (&x)[D], const float (&mean)[D], float scale, float lognorm) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        const float v = (x[j] - mean[j]) / scale;
        s += -0.5f * v * v - lognorm;
    }
    return s;
}

__device__ __forceinline__ void tw_quad_sync(int warp) {   // warps w, w + 4, w + 8, w + 12: the four threads of a particle
    asm volatile("bar.sync %0, 128;" :: "r"(1 + (warp & 3)) : "memory");
}

// one 16-unit chunk of layer 1: a1 = act(U1^T x + c1[t]) -> tf32 hi (16 columns) + bf16 lo (8 packed columns) of the thread's
// TMEM lane; geffner: acc += a1 W3 (residual skip, nn.py:70)
template <int D, int ACT, int HT>
__device__ __forceinline__ void tw_layer1_chunk(const TwCtx<D>& cx, int c, const f32x2_t (&xb)[D], f32x2_t (&acc)[D]) {
    using G = TwGeo<HT>;
    const float4* __restrict__ c1v = reinterpret_cast<const float4*>(cx.tab);
    uint32_t h[16], l[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 cc = c1v[c * 4 + q];
        f32x2_t p01 = pk2(cc.x, cc.y), p23 = pk2(cc.z, cc.w);
#pragma unroll
        for (int a = 0; a < D; ++a) {
            const float4 u = *reinterpret_cast<const float4*>(cx.sU1 + a * HT + c * 16 + q * 4);
            p01 = fma2(xb[a], pk2(u.x, u.y), p01);
            p23 = fma2(xb[a], pk2(u.z, u.w), p23);
        }
        float p[4], av[4], lov[4];
        upk2(p01, p[0], p[1]); upk2(p23, p[2], p[3]);
        if constexpr (ACT == ACT_GELU) {   // two activations per instruction slot (FFMA2)
            upk2(gelu_fast2(p[0], p[1]), av[0], av[1]);
            upk2(gelu_fast2(p[2], p[3]), av[2], av[3]);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) av[e] = softplus_fast(p[e]);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float hi;
            umma::split_tf32(av[e], hi, lov[e]);
            h[q * 4 + e] = __float_as_uint(hi);
        }
        if constexpr (ACT == ACT_SOFTPLUS) {
#pragma unroll
            for (int m = 0; m < D; ++m) {   // acc[m] = (sum over even units, sum over odd units) of a1 W3[., m]
                const float4 w = *reinterpret_cast<const float4*>(cx.sW3T + m * HT + c * 16 + q * 4);
                acc[m] = fma2(pk2(av[0], av[1]), pk2(w.x, w.y), acc[m]);
                acc[m] = fma2(pk2(av[2], av[3]), pk2(w.z, w.w), acc[m]);
            }
        }
        const __nv_bfloat162 l01 = __floats2bfloat162_rn(lov[0], lov[1]), l23 = __floats2bfloat162_rn(lov[2], lov[3]);
        l[q * 2 + 0] = *reinterpret_cast<const uint32_t*>(&l01);
        l[q * 2 + 1] = *reinterpret_cast<const uint32_t*>(&l23);
    }
    umma::tmem_st16(cx.tmem_lane + G::COL_AH + c * 16, h);
    umma::tmem_st8(cx.tmem_lane + G::COL_AL + c * 8, l);
}

// the MMAs of one K stage for one half of N; no accumulation on a half's very first instruction (stage 0).
// 144: stage 0 = chunks {0, 2, 4, 6, 7}, stage 1 = {1, 3, 5, 8}; 64: one stage {0, 1, 2, 3}, one "half" of 64 columns.
// A chunk = two tf32 K blocks (2c, 2c + 1) = one bf16 K block.
template <int HT, int HALF, int STAGE>
__device__ __forceinline__ void tw_issue_half(uint32_t tmem_base, uint64_t bhi, uint64_t blo, uint64_t bhi16) {
    using G = TwGeo<HT>;
    constexpr int n0 = HALF ? G::NA : 0, nn = HALF ? HT - G::NA : G::NA;
    constexpr int NC = HT == 144 ? (STAGE ? 4 : 5) : 4;
    constexpr int chunks[3][5] = {{0, 2, 4, 6, 7}, {1, 3, 5, 8, 0}, {0, 1, 2, 3, 0}};
    constexpr int row = HT == 144 ? STAGE : 2;
    const uint32_t idesc = umma::make_idesc_tf32(128, nn), idesc16 = umma::make_idesc_bf16_k(128, nn);
    const uint32_t dcol = tmem_base + G::COL_D + n0, ahi = tmem_base + G::COL_AH, alo = tmem_base + G::COL_AL;
    // rows n0.. of a K-major core-matrix tile: (n0 / 8) * SBO bytes further (SBO = 32 K bytes for tf32, 16 K for bf16), >> 4 in the descriptor
    constexpr uint64_t o32 = (uint64_t)((n0 / 8) * 32 * HT >> 4), o16 = (uint64_t)((n0 / 8) * 16 * HT >> 4);
    const uint64_t dlo = blo + o32, dhi = bhi + o32, dhi16 = bhi16 + o16;
    // Next K block of a tile: +256 B = +16 in the descriptor's address field.  Small terms first within a stage.
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        const int c = chunks[row][i];
#pragma unroll
        for (int k = 2 * c; k < 2 * c + 2; ++k) umma::mma_tf32_ts(dcol, ahi + k * 8, dlo + (uint64_t)(k * 16), idesc, (STAGE || i || k > 2 * c) ? 1u : 0u);
    }
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        const int c = chunks[row][i];
        umma::mma_f16_ts(dcol, alo + c * 8, dhi16 + (uint64_t)(c * 16), idesc16, 1);
    }
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        const int c = chunks[row][i];
#pragma unroll
        for (int k = 2 * c; k < 2 * c + 2; ++k) umma::mma_tf32_ts(dcol, ahi + k * 8, dhi + (uint64_t)(k * 16), idesc, 1);
    }
}

// one stage of layer 1; warp 0 issues the stage's MMAs once all 512 threads have stored theirs
template <int D, int ACT, int HT, int STAGE>
__device__ __forceinline__ void tw_stage(TwCtx<D>& cx, int next_t, const f32x2_t (&xb)[D], f32x2_t (&acc)[D]) {
    using G = TwGeo<HT>;
    const int cb = STAGE ? G::chunk_split(cx.grp) : G::chunk_begin(cx.grp);
    const int ce = STAGE ? G::chunk_end(cx.grp) : G::chunk_split(cx.grp);
#pragma unroll 1
    for (int c = cb; c < ce; ++c) tw_layer1_chunk<D, ACT, HT>(cx, c, xb, acc);
    umma::tmem_st_wait();
    umma::fence_before();
    umma::mbar_arrive(cx.mbar_ready + STAGE);
    if (threadIdx.x < 32) {   // warp 0 (converged): wait for the stage's 512 arrivals, one elected lane issues
        umma::mbar_wait(cx.mbar_ready + STAGE, cx.parity);
        umma::fence_after();
        if (umma::elect_one()) {
            if (STAGE == 0 && next_t >= 0) {
                // every thread of the CTA has arrived for this node, i.e. is done with the rows of the node before: their buffer is
                // free.  Both rows of the next node through the TMA engine (two bulk copies, one mbarrier).
                const int b = next_t & 1;
                const uint32_t row = (uint32_t)cx.HP * sizeof(float);
                umma::mbar_arrive_expect_tx(cx.tab_bar + b, 2 * row);
                umma::bulk_copy_g2s(cx.tab_shared + b * (2 * HT), cx.c1 + (size_t)next_t * cx.HP, row, cx.tab_bar + b);
                umma::bulk_copy_g2s(cx.tab_shared + b * (2 * HT) + HT, cx.c2 + (size_t)next_t * cx.HP, row, cx.tab_bar + b);
            }
            constexpr bool last = STAGE == G::NSTAGE - 1;
            tw_issue_half<HT, 0, STAGE>(cx.tmem_base, cx.bhi, cx.blo, cx.bhi16);
            if (last) umma::commit(cx.mbar_done + 0);
            if constexpr (G::NHALF == 2) {
                tw_issue_half<HT, 1, STAGE>(cx.tmem_base, cx.bhi, cx.blo, cx.bhi16);
                if (last) umma::commit(cx.mbar_done + 1);
            }
        }
        __syncwarp();
    }
}

// layer 1 of this thread's units in one or two stages.  acc receives this group's share of a1 W3 (as even / odd unit partial sums).
template <int D, int ACT, int HT>
__device__ __forceinline__ void tw_net_issue(TwCtx<D>& cx, int next_t, const float (&x)[D], f32x2_t (&acc)[D]) {
    f32x2_t xb[D];
#pragma unroll
    for (int a = 0; a < D; ++a) xb[a] = pk2(x[a], x[a]);
#pragma unroll
    for (int m = 0; m < D; ++m) acc[m] = pk2(0.f, 0.f);
    tw_stage<D, ACT, HT, 0>(cx, next_t, xb, acc);
    if constexpr (TwGeo<HT>::NSTAGE == 2) tw_stage<D, ACT, HT, 1>(cx, next_t, xb, acc);
}

// wait for this group's half of the MMA, epilogue over its units: acc += W3^T act(D + c2[t] + U2^T x)   (dds: no U2 term)
template <int D, int ACT, int HT>
__device__ __forceinline__ void tw_net_finish(TwCtx<D>& cx, const float (&x)[D], f32x2_t (&acc)[D]) {
    using G = TwGeo<HT>;
    const float4* __restrict__ c2v = reinterpret_cast<const float4*>(cx.tab + HT);
    const int cb = G::chunk_begin(cx.grp), ce = G::chunk_end(cx.grp);
    umma::mbar_wait(cx.mbar_done + G::done_bar(cx.grp), cx.parity);
    cx.parity ^= 1u;
    umma::fence_after();
#pragma unroll 1
    for (int c = cb; c < ce; ++c) {
        uint32_t v[16];
        umma::tmem_ld16(cx.tmem_lane + G::COL_D + c * 16, v);
        umma::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 cc = c2v[c * 4 + q];
            f32x2_t p01 = add2(pk2(__uint_as_float(v[q * 4 + 0]), __uint_as_float(v[q * 4 + 1])), pk2(cc.x, cc.y));
            f32x2_t p23 = add2(pk2(__uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3])), pk2(cc.z, cc.w));
            f32x2_t A01, A23;
            if constexpr (ACT == ACT_GELU) {
                float p0, p1, p2, p3;
                upk2(p01, p0, p1); upk2(p23, p2, p3);
                A01 = gelu_fast2(p0, p1); A23 = gelu_fast2(p2, p3);
            } else {
#pragma unroll
                for (int a = 0; a < D; ++a) {
                    const float4 u = *reinterpret_cast<const float4*>(cx.sU2 + a * HT + c * 16 + q * 4);
                    const f32x2_t xa = pk2(x[a], x[a]);
                    p01 = fma2(xa, pk2(u.x, u.y), p01);
                    p23 = fma2(xa, pk2(u.z, u.w), p23);
                }
                float p[4];
                upk2(p01, p[0], p[1]); upk2(p23, p[2], p[3]);
                A01 = pk2(softplus_fast(p[0]), softplus_fast(p[1])); A23 = pk2(softplus_fast(p[2]), softplus_fast(p[3]));
            }
#pragma unroll
            for (int m = 0; m < D; ++m) {
                const float4 w = *reinterpret_cast<const float4*>(cx.sW3T + m * HT + c * 16 + q * 4);
                acc[m] = fma2(A01, pk2(w.x, w.y), acc[m]);
                acc[m] = fma2(A23, pk2(w.z, w.w), acc[m]);
            }
        }
    }
    umma::fence_before();   // orders these tcgen05.ld before the next evaluation's writes to D (via the next mbarrier arrive / wait)
}

template <int D, int ACT, int HT>
__global__ void __launch_bounds__(TW_THREADS, 1) bridge_fwd_tcw_kernel(const BridgeArgs a) {
    using G = TwGeo<HT>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t mbar_ready[2], mbar_done[2], tab_bar[2];
    __shared__ __align__(128) float tab_shared[2][2 * HT];
    // exchange between the four threads of a particle: network input (group 0 -> others), partial outputs (others -> group 0),
    // the step's Gaussians (group 1 -> 0), target score and log-density (group 2 -> 0)
    __shared__ __align__(16) float sX[TW_PB * D], sO[(TW_G - 1) * TW_PB * D], sXi[TW_PB * D], sSp[TW_PB * (D + 1)];
    const int tid = threadIdx.x, warp = tid >> 5, grp = tid >> 7, pl = tid & (TW_PB - 1);
    const NetView& nv = a.net;
    const int HP = nv.HP;   // rows of the network arrays are HP long; columns HP..HT-1 of every staged copy are zero
    uint8_t* sBhi = smem_raw;
    uint8_t* sBlo = smem_raw + G::B_BYTES;
    uint8_t* sBhi16 = smem_raw + 2 * G::B_BYTES;
    float* sf = reinterpret_cast<float*>(smem_raw + 2 * G::B_BYTES + G::B16_BYTES);
    float* sU1 = sf;
    float* sU2 = sU1 + D * HT;
    float* sW3T = sU2 + D * HT;
    float* sU3 = sW3T + HT * D;
    float* sTp = sU3 + ((D * D + 3) & ~3);
    // B[n = j][k = i] = W2[i][j], split into tf32 hi / lo
    for (int base = 0; base < HT * HT; base += 8 * TW_THREADS) {   // eight loads in flight per thread
        float wv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = base + u * TW_THREADS + tid, i = idx / HT, j = idx % HT;
            wv[u] = (idx < HT * HT && i < HP && j < HP) ? __ldg(nv.W2 + i * HP + j) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = base + u * TW_THREADS + tid, i = idx / HT, j = idx % HT;
            if (idx >= HT * HT) continue;
            float hi, lo;
            umma::split_tf32(wv[u], hi, lo);
            const int off = umma::core_off(j, i, HT);
            *reinterpret_cast<float*>(sBhi + off) = hi;
            *reinterpret_cast<float*>(sBlo + off) = lo;
            *reinterpret_cast<__nv_bfloat16*>(sBhi16 + umma::core_off16(j, i, HT)) = __float2bfloat16(hi);
        }
    }
    for (int idx = tid; idx < D * HT; idx += TW_THREADS) {
        const int r = idx / HT, j = idx % HT;
        sU1[idx] = j < HP ? nv.U1[r * HP + j] : 0.f;
        sU2[idx] = (nv.U2 && j < HP) ? nv.U2[r * HP + j] : 0.f;
    }
    // padded hidden units: softplus(0) != 0, their output weights are 0
    for (int idx = tid; idx < HT * D; idx += TW_THREADS) sW3T[(idx % D) * HT + idx / D] = idx < HP * D ? nv.W3[idx] : 0.f;
    for (int i = tid; i < D * D; i += TW_THREADS) sU3[i] = nv.U3 ? nv.U3[i] : 0.f;
    const int ntp = (a.tgt.kind == TGT_GMM || a.tgt.kind == TGT_MANY_GMM) ? a.tgt.ncomp * MIX_STRIDE : 0;
    for (int i = tid; i < ntp; i += TW_THREADS) sTp[i] = a.tgt.mix[i];
    float2* sMu = reinterpret_cast<float2*>(sTp + MIX_MAX * MIX_STRIDE);   // many_gmm: dense component means
    for (int i = tid; i < 2 * 2 * HT; i += TW_THREADS) (&tab_shared[0][0])[i] = 0.f;
    const bool fast_gmm = (D == 2) && (a.tgt.kind == TGT_MANY_GMM);
    if (fast_gmm)
        many_gmm_stage_means(a.tgt, sMu, tid, TW_THREADS);
    const ManyGmmConst gc = many_gmm_const(a.tgt);
    if (warp == 0) umma::tmem_alloc(&tmem_slot, G::COLS, true);
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { umma::mbar_init(&mbar_ready[i], TW_THREADS); umma::mbar_init(&mbar_done[i], 1); umma::mbar_init(&tab_bar[i], 1); }
    }
    umma::fence_async_smem();   // generic-proxy writes of the B tiles (and the zeroed table buffers) -> visible to the async proxy
    umma::fence_before();
    __syncthreads();
    umma::fence_after();

    TwCtx<D> cx;
    cx.sU1 = sU1; cx.sU2 = sU2; cx.sW3T = sW3T; cx.sU3 = sU3;
    cx.c1 = nv.c1; cx.c2 = nv.c2; cx.c3 = nv.c3;
    cx.out_scale = net_out_scale(nv); cx.out_clip = nv.out_clip;
    cx.HP = HP; cx.grp = grp;
    cx.tmem_base = tmem_slot;
    cx.tmem_lane = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16);   // lane quarter = warp % 4
    cx.bhi16 = umma::make_desc(umma::smem_u32(sBhi16), 128, 16 * HT);
    cx.bhi = umma::make_desc(umma::smem_u32(sBhi), 128, 32 * HT);
    cx.blo = umma::make_desc(umma::smem_u32(sBlo), 128, 32 * HT);
    cx.mbar_ready = mbar_ready; cx.mbar_done = mbar_done; cx.parity = 0u;
    cx.tab_shared = &tab_shared[0][0]; cx.tab_bar = tab_bar; cx.tab_parity = 0u;

    const bool cais = (a.mode == CMCD_MODE_CAIS_SN || a.mode == CMCD_MODE_CAIS_VAR_SN);
    const bool nn_f = cais;
    const int K = a.K;
    const int t0 = cais ? 0 : -1;   // node form, see bridge_fwd_tc.cu: NN(z_j, j) (CAIS) / NN(z_j, j - 1) (MCD_ULA_sn)

    float mu[D], sig[D];
#pragma unroll
    for (int j = 0; j < D; ++j) { mu[j] = a.vd_mean[j]; sig[j] = expf(a.vd_logdiag[j]); }

    const long long ntiles = (a.N + TW_PB - 1) / TW_PB;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        __syncthreads();   // the previous tile's last node may still be reading table buffer 0 / the exchange buffers in a lagging warp
        if (tid == 0) {    // rows of the first evaluation (t = 0 in both node forms)
            const uint32_t row = (uint32_t)HP * sizeof(float);
            umma::mbar_arrive_expect_tx(&tab_bar[0], 2 * row);
            umma::bulk_copy_g2s(&tab_shared[0][0], nv.c1, row, &tab_bar[0]);
            umma::bulk_copy_g2s(&tab_shared[0][HT], nv.c2, row, &tab_bar[0]);
        }
        auto wait_tab = [&](int t) {
            const int b = t & 1;
            umma::mbar_wait(&tab_bar[b], (cx.tab_parity >> b) & 1u);
            cx.tab_parity ^= 1u << b;
            cx.tab = &tab_shared[b][0];
        };
        const long long n_raw = tile * TW_PB + pl;
        const bool active = n_raw < a.N;
        const long long n = active ? n_raw : a.N - 1;   // tail lanes shadow the last particle (all 128 lanes take part in the MMA)
        // Roles besides the hidden units: group 0 = state of the particle (z, log-weight, kernel means); group 1 = its key chain and
        // Gaussians (independent of the trajectory); group 2 = target score at every node; group 3 = a third chunk of hidden units.
        const bool ev = a.z0 != nullptr;   // mcd_utils.evolve entry: (z, rng_key_gen) given by the caller (mcd_utils.py:24-33)
        Key k, ka;
        if (ev) { ka.k0 = a.keys[2 * n]; ka.k1 = a.keys[2 * n + 1]; }
        else { k = prng_key(a.seeds[n]); split(k, ka, k); }
        float z[D], xi[D], x[D], mf[D];
        float w = 0.f, wm = 0.f, lp = 0.f;
        if (grp == 0) {
            if (ev) {
#pragma unroll
                for (int j = 0; j < D; ++j) z[j] = a.z0[n * D + j];
            } else {   // z0 = sigma*xi + mu ; w = -log q(z0)   (vardist/diag_gauss.py:26-33,44-62)
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
#pragma unroll
            for (int j = 0; j < D; ++j) { x[j] = z[j]; mf[j] = 0.f; }
            if (a.traj && active) {
#pragma unroll
                for (int j = 0; j < D; ++j) a.traj[((size_t)0 * D + j) * a.N + n] = z[j];
            }
        }
        if (!ev) ka = split_first(k);   // mcdboundingmachine.py:162 (evolve entry: ka is the caller's rng_key_gen)
        k = split_second(ka);           // mcd_cais.py:94      (advanced by group 1 only)
        float beta = 0.f, eps = 0.f, scale = 1.f;
        // Group 0's serial section between the two barriers of consecutive nodes is on every thread's critical path, so only what
        // the next node's input needs stays there (combine, kernel mean, sample); the log-weight update of the step just closed
        // (two Gaussian log-densities, a logf, the trajectory store) is parked in these registers and done in the next node's
        // MMA shadow (or after the last node).
        bool pend = false;
        int pend_nd = 0;
        float pw_x[D], pw_z[D], pw_mf[D], pw_nn[D], pw_gu[D], pw_gq[D], pw_beta = 0.f, pw_eps = 0.f, pw_scale = 1.f;
        auto weight_update = [&]() {   // backward kernel of step pend_nd - 1, weight update (statement order of bridge_fwd_tc_kernel)
            float mb[D];
#pragma unroll
            for (int j = 0; j < D; ++j) {
                const float ub = -(pw_beta * pw_gu[j] + (1.0f - pw_beta) * pw_gq[j]);
                mb[j] = pw_x[j] - pw_eps * ub;
                mb[j] = mb[j] + pw_eps * pw_nn[j];
            }
            const float lognorm = logf(2.5066282746310002f * pw_scale);
            const float fk = tw_gauss_logprob<D>(pw_x, pw_mf, pw_scale, lognorm);
            const float bk = tw_gauss_logprob<D>(pw_z, mb, pw_scale, lognorm);
            wm += bk - fk;
            if (a.traj && active) {
#pragma unroll
                for (int j = 0; j < D; ++j) a.traj[((size_t)pend_nd * D + j) * a.N + n] = pw_x[j];
            }
            pend = false;
        };
        // K + 1 nodes z_0 .. z_K, ONE network evaluation per node (node form, see bridge_fwd_tc.cu)
        for (int nd = 0; nd <= K; ++nd) {
            const int t = t0 + nd;
            const bool use_nn = cais || nd > 0;
            f32x2_t acc[D];
            float c3v[D];
            float bnx = 0.f, enx = 0.f;   // step constants of this node's forward kernel, requested before layer 1
            if (grp == 0) {
#pragma unroll
                for (int j = 0; j < D; ++j) sX[pl * D + j] = x[j];
                if (nd < K) { bnx = __ldg(a.betas + nd); enx = __ldg(a.eps + nd); }
                if (use_nn) {
#pragma unroll
                    for (int m = 0; m < D; ++m) c3v[m] = __ldg(cx.c3 + (size_t)t * D + m);   // requested before layer 1, used after it
                }
            }
            tw_quad_sync(warp);                      // x of this node is in sX
            if (grp != 0) {
#pragma unroll
                for (int j = 0; j < D; ++j) x[j] = sX[pl * D + j];
            }
            if (use_nn) {
                wait_tab(t);
                tw_net_issue<D, ACT, HT>(cx, nd < K ? t + 1 : -1, x, acc);
            }
            // ---- per-particle work that does not depend on the network output, in the shadow of the MMA batch ----
            float gq[D];
            if (grp == 0) {
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    const float sq = -((x[j] - mu[j]) / sig[j]) / sig[j];
                    gq[j] = fminf(fmaxf(sq, -a.clip_q), a.clip_q);
                }
                if (pend) weight_update();
                if (use_nn) {
                    // out = out_scale * clamp(W3^T (a2 + a1) + U3^T x + c3[t]): the terms outside the hidden units
#pragma unroll
                    for (int m = 0; m < D; ++m) {
                        float p = c3v[m];
#pragma unroll
                        for (int q = 0; q < D; ++q) p = fmaf(x[q], sU3[q * D + m], p);
                        acc[m] = add2(acc[m], pk2(p, 0.f));
                    }
                }
            } else if (grp == 1) {
                if (nd < K) {   // split + Gaussian + key advance of step nd
                    step_keys_and_normal<D>(k, xi);
#pragma unroll
                    for (int j = 0; j < D; ++j) sXi[pl * D + j] = xi[j];
                }
            } else if (grp == 2) {
                float sp[D], dummy[D];
                float lpn;
                if (fast_gmm) { float d0, d1; lpn = many_gmm_eval<false>(gc, sMu, x[0], x[1], sp[0], sp[1], 0.f, 0.f, d0, d1); }
                else lpn = target_eval<D, false>(a.tgt, sTp, x, sp, dummy, dummy);
#pragma unroll
                for (int j = 0; j < D; ++j) sSp[pl * (D + 1) + j] = sp[j];
                sSp[pl * (D + 1) + D] = lpn;
            }
            if (use_nn) {
                tw_net_finish<D, ACT, HT>(cx, x, acc);
                if (grp != 0) {
#pragma unroll
                    for (int j = 0; j < D; ++j) {
                        float e0, e1;
                        upk2(acc[j], e0, e1);
                        sO[((grp - 1) * TW_PB + pl) * D + j] = e0 + e1;
                    }
                }
            }
            tw_quad_sync(warp);                      // partial outputs, Gaussians and target score handed to group 0
            if (grp != 0) continue;
            float nnv[D], gu[D];
#pragma unroll
            for (int m = 0; m < D; ++m) {
                nnv[m] = 0.f;
                if (use_nn) {
                    float o, o1;
                    upk2(acc[m], o, o1);
                    o += o1;
#pragma unroll
                    for (int g = 0; g < TW_G - 1; ++g) o += sO[(g * TW_PB + pl) * D + m];
                    nnv[m] = cx.out_scale * fminf(fmaxf(o, -cx.out_clip), cx.out_clip);
                }
                gu[m] = fminf(fmaxf(sSp[pl * (D + 1) + m], -a.clip_t), a.clip_t);
                if (nd < K) xi[m] = sXi[pl * D + m];
            }
            lp = sSp[pl * (D + 1) + D];
            if (nd > 0) {   // the step nd - 1 just closed: park its log-weight update (beta, eps, scale, mf still hold that step's values)
                pend = true; pend_nd = nd; pw_beta = beta; pw_eps = eps; pw_scale = scale;
#pragma unroll
                for (int j = 0; j < D; ++j) { pw_x[j] = x[j]; pw_z[j] = z[j]; pw_mf[j] = mf[j]; pw_nn[j] = nnv[j]; pw_gu[j] = gu[j]; pw_gq[j] = gq[j]; }
#pragma unroll
                for (int j = 0; j < D; ++j) z[j] = x[j];
                // (144-wide tiles: group 3's three chunks are the critical path, not this section -- measured 1.62 vs 1.65 ms --
                //  and the parked state costs the registers the epilogue needs: update in place)
                if constexpr (HT == 144) weight_update();
            }
            if (nd < K) {   // forward kernel of step nd: mean, sample z_{nd+1}
                beta = bnx; eps = enx;
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
        if (grp == 0 && pend) weight_update();
        if (grp == 0 && active) {
            w += wm;
            w += lp;
            // evolve returns the steps' log-ratio sum (mcd_cais.py:98-99); compute_log_elbo adds -log q(z_0) and log p(z_K)
            a.out_negw[n] = ev ? wm : -w;
#pragma unroll
            for (int j = 0; j < D; ++j) a.out_z[n * D + j] = z[j];
        }
    }
    umma::fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(cx.tmem_base, G::COLS);
}

template <int HT>
static size_t tw_smem_bytes(int D) {
    using G = TwGeo<HT>;
    return 2 * G::B_BYTES + G::B16_BYTES + (2 * D * HT + HT * D + D * D + 4 + MIX_MAX * MIX_STRIDE + 2 * MIX_MAX + 8) * sizeof(float);
}

// (1) geffner net with hidden_pad in 129..144 at d = 2 (README.md:30,34), any N; (2) hidden_pad 64 (dds, narrow geffner nets) at d = 2
// when the batch is at most one 128-particle tile per SM -- there the three-CTAs-per-SM kernel of bridge_fwd_tc.cu has one thread walk
// a node's whole dependency chain (5.5 us per node) and four threads per particle are faster.  Both need a network evaluation on every
// node that has one (CAIS modes, MCD_ULA_sn) and 16-byte aligned table rows for the bulk copies.
// CMCD_TC_WIDE=0 / CMCD_TC_QUAD=0 keep the other mappings (A/B runs, tests).
static int tw_tile_width(const BridgeArgs& a, int D, int num_sms) {
    if (D != 2 || a.net.arch == CMCD_ARCH_NONE || a.K < 1 || a.mode == CMCD_MODE_ULA) return 0;
    if ((reinterpret_cast<uintptr_t>(a.net.c1) | reinterpret_cast<uintptr_t>(a.net.c2)) & 15) return 0;
    const char* ew = std::getenv("CMCD_TC_WIDE");
    const char* eq = std::getenv("CMCD_TC_QUAD");
    if (a.net.arch == CMCD_ARCH_GEFFNER && a.net.HP > 128 && a.net.HP <= 144 && !(a.net.HP & 3)) return (ew && ew[0] == '0') ? 0 : 144;
    if (a.net.HP == 64 && (a.N + TW_PB - 1) / TW_PB <= num_sms) return (eq && eq[0] == '0') ? 0 : 64;
    return 0;
}
bool fwd_tcw_supported(const BridgeArgs& a, int D, int num_sms) { return tw_tile_width(a, D, num_sms) != 0; }

template <int ACT, int HT>
static int launch_fwd_tcw_t(const BridgeArgs& a, cudaStream_t st, int num_sms) {
    const size_t smem = tw_smem_bytes<HT>(2);
    auto kern = bridge_fwd_tcw_kernel<2, ACT, HT>;
    CMCD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long ntiles = (a.N + TW_PB - 1) / TW_PB;
    long long grid = num_sms;
    if (grid > ntiles) grid = ntiles;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, TW_THREADS, smem, st>>>(a);
    CMCD_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_bridge_fwd_tcw(const BridgeArgs& a, int D, cudaStream_t st, int num_sms) {
    const int ht = tw_tile_width(a, D, num_sms);
    if (ht == 144) return launch_fwd_tcw_t<ACT_SOFTPLUS, 144>(a, st, num_sms);
    if (ht == 64) return a.net.arch == CMCD_ARCH_DDS ? launch_fwd_tcw_t<ACT_GELU, 64>(a, st, num_sms) : launch_fwd_tcw_t<ACT_SOFTPLUS, 64>(a, st, num_sms);
    set_error("bridge_fwd_tcw: no instantiation for dim=%d hidden_pad=%d", D, a.net.HP);
    return 2;
}

}  // namespace cmcd
