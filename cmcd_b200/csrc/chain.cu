// The O(K) "host chain" around the bridge kernels, fused: everything compute_log_elbo does per call that depends only on the
// parameters and the step index -- not on the particles -- as ONE prologue launch and TWO epilogue launches instead of the
// ~100 small framework kernels (and their autograd mirror images) per train iteration.
//
// Replaces, in /root/reference/src:
//   betas from mgridref_y:  cumsum / sum, concatenate 0, jnp.interp(target_x, gridref_x, gridref_y)   mcdboundingmachine.py:146-149
//   per-step step sizes:    constant / linear / cos^2 schedule                                         mcd_cais.py:34-44,54-59
//   dds (PISNet) step part: e = [sin, cos](coeff * t + phase); t_emb = Lin(gelu(Lin(e)));               nn_dds.py:130-143,156-158
//                           first state layer applied to t_emb  ->  c1[t] = t_emb W1[d:] + b1           nn_dds.py:159-161
//   geffner step part:      emb[min(t, K-1)] pushed through the emb rows of the three Dense layers      nn.py:62-70 (clamped gather)
// and, in reverse mode, the transposes of all of the above (what jax.grad derives for these O(K) pieces), writing the flat
// parameter gradient in the layout of ravel_pytree((params_train, params_notrain)) (mcdboundingmachine.py:122).
//
// Prologue  chain_fwd_kernel:   blocks [0, T): one table row each (T = K + 1); block T: betas + eps; further blocks: zero-padded
//                               copies of the geffner weights (hidden -> hidden_pad).
// Epilogue  chain_bwd_rows_kernel (per row: recompute the row's forward values, pull the table cotangents back to the row's
//           inputs, leave the per-row factors of the weight gradients in scratch; block T: betas / eps / q transposes) and
//           chain_bwd_weights_kernel (one thread per weight element: the sums over the T rows, in ascending row order ->
//           deterministic; plus the plain copies U1/W2/W3 cotangent -> weight rows).
#include "common.cuh"

namespace cmcd {

constexpr int CH_C = 64;   // PISNet channels (nn_dds.py:95)

struct ChainView {
    cmcd_chain c;
    const float* p;   // params_flat
};

__device__ __forceinline__ float ch_gelu(float x) { return x * 0.5f * (1.0f + erff(x * 0.70710678118654752440f)); }   // nn_dds.py:167-176
__device__ __forceinline__ float ch_gelu_grad(float x) {
    return 0.5f * (1.0f + erff(x * 0.70710678118654752440f)) + x * 0.3989422804014327f * expf(-0.5f * x * x);
}

// eps_i = eps0 * decay_i; returns decay_i = d eps_i / d eps0
__device__ __forceinline__ float ch_eps_decay(int schedule, int i, int K) {
    if (schedule == CMCD_EPS_COS_SQ) {     // mcd_cais.py:38-44: eps * cos^2(((i/K) + 0.008) / 1.008 * pi/2), i/K true division in float32
        const float phase = (float)i / (float)K;
        const float cs = cosf((phase + 0.008f) / 1.008f * 0.5f * 3.14159265358979323846f);
        return cs * cs;
    }
    return 1.0f;
}

// one row of the dds time coder: code[128], pre[64], h[64], tnet[64] in shared memory (blockDim.x == 128)
__device__ __forceinline__ void ch_dds_row_forward(const cmcd_chain& c, const float* __restrict__ p, int t, float* code, float* pre, float* h,
                                                   float* tnet) {
    const int tid = threadIdx.x;
    {
        const int ch = tid & (CH_C - 1);
        const float arg = c.dds_coeff[ch] * (float)t + p[c.off[CMCD_LEAF_DDS_PHASE] + ch];
        code[tid] = tid < CH_C ? sinf(arg) : cosf(arg);
    }
    __syncthreads();
    if (tid < CH_C) {
        const float* w = p + c.off[CMCD_LEAF_DDS_TC1_W];
        float s = 0.f;
        for (int i = 0; i < 2 * CH_C; ++i) s = fmaf(code[i], w[i * CH_C + tid], s);
        s += p[c.off[CMCD_LEAF_DDS_TC1_B] + tid];
        pre[tid] = s;
        h[tid] = ch_gelu(s);
    }
    __syncthreads();
    if (tid < CH_C) {
        const float* w = p + c.off[CMCD_LEAF_DDS_TC2_W];
        float s = 0.f;
        for (int i = 0; i < CH_C; ++i) s = fmaf(h[i], w[i * CH_C + tid], s);
        tnet[tid] = s + p[c.off[CMCD_LEAF_DDS_TC2_B] + tid];
    }
    __syncthreads();
}

__global__ void __launch_bounds__(128) chain_fwd_kernel(const ChainView v, float* __restrict__ betas, float* __restrict__ eps, float* __restrict__ c1,
                                                        float* __restrict__ c2, float* __restrict__ c3, float* __restrict__ U1p, float* __restrict__ U2p,
                                                        float* __restrict__ W2p, float* __restrict__ W3p) {
    const cmcd_chain& c = v.c;
    const float* __restrict__ p = v.p;
    const int K = c.nbridges, T = K + 1, d = c.in_dim, dout = c.dim, H = c.hidden, HP = c.hidden_pad, tid = threadIdx.x;
    __shared__ float sh[2 * CH_C + 3 * CH_C + 40];
    const int b = blockIdx.x;
    if (b < T && c.arch == CMCD_ARCH_DDS) {
        float *code = sh, *pre = sh + 2 * CH_C, *h = pre + CH_C, *tnet = h + CH_C;
        ch_dds_row_forward(c, p, b, code, pre, h, tnet);
        if (tid < CH_C) {
            const float* w = p + c.off[CMCD_LEAF_DDS_ST1_W] + (size_t)d * CH_C;   // rows of the first state layer that see t_emb
            float s = 0.f;
            for (int i = 0; i < CH_C; ++i) s = fmaf(tnet[i], w[i * CH_C + tid], s);
            c1[(size_t)b * HP + tid] = s + p[c.off[CMCD_LEAF_DDS_ST1_B] + tid];
            c2[(size_t)b * HP + tid] = p[c.off[CMCD_LEAF_DDS_ST2_B] + tid];
        }
        if (tid < dout) c3[(size_t)b * dout + tid] = p[c.off[CMCD_LEAF_DDS_OUT_B] + tid];
        return;
    }
    if (b < T && c.arch == CMCD_ARCH_GEFFNER) {
        // e = emb[min(t, K-1)] (JAX clamps the out-of-range gather of step K, nn.py:68); c_l[t] = e W_l[d:] + b_l
        const int E = c.emb_dim, r = b < K ? b : K - 1;
        const float* e = p + c.off[CMCD_LEAF_GEF_EMB] + (size_t)r * E;
        const float* w1 = p + c.off[CMCD_LEAF_GEF_W1] + (size_t)d * H;
        const float* w2 = p + c.off[CMCD_LEAF_GEF_W2] + (size_t)d * H;
        const float* w3 = p + c.off[CMCD_LEAF_GEF_W3] + (size_t)d * dout;
        for (int j = tid; j < HP; j += blockDim.x) {
            float s1 = 0.f, s2 = 0.f;
            if (j < H) {
                for (int i = 0; i < E; ++i) { const float ei = e[i]; s1 = fmaf(ei, w1[(size_t)i * H + j], s1); s2 = fmaf(ei, w2[(size_t)i * H + j], s2); }
                s1 += p[c.off[CMCD_LEAF_GEF_B1] + j];
                s2 += p[c.off[CMCD_LEAF_GEF_B2] + j];
            }
            c1[(size_t)b * HP + j] = s1;
            c2[(size_t)b * HP + j] = s2;
        }
        for (int j = tid; j < dout; j += blockDim.x) {
            float s = 0.f;
            for (int i = 0; i < E; ++i) s = fmaf(e[i], w3[(size_t)i * dout + j], s);
            c3[(size_t)b * dout + j] = s + p[c.off[CMCD_LEAF_GEF_B3] + j];
        }
        return;
    }
    if (b == T) {
        // betas: gridref_y = [0, cumsum(m) / sum(m)]; betas = interp(target_x, gridref_x, gridref_y)   (mcdboundingmachine.py:146-149)
        float* gy = sh;   // ngrid + 1 <= 40 values
        const int G1 = c.ngrid;
        if (tid == 0) {
            const float* m = p + c.off[CMCD_LEAF_MGRID_Y];
            float tot = 0.f;
            for (int k = 0; k < G1; ++k) tot += m[k];
            float run = 0.f;
            gy[0] = 0.f;
            for (int k = 0; k < G1; ++k) { run += m[k]; gy[k + 1] = run / tot; }
        }
        __syncthreads();
        const float* xp = p + c.off[CMCD_LEAF_GRID_X];
        const float* tx = p + c.off[CMCD_LEAF_TARGET_X];
        const float eps0 = p[c.off[CMCD_LEAF_EPS]];
        for (int i = tid; i < K; i += blockDim.x) {
            const float x = tx[i];
            int s = 0;                                   // searchsorted(xp, x, right = True), clamped to [1, n - 1]
            while (s < G1 + 1 && xp[s] <= x) ++s;
            s = min(max(s, 1), G1);
            const float df = gy[s] - gy[s - 1], dx = xp[s] - xp[s - 1], delta = x - xp[s - 1];
            betas[i] = (dx == 0.f) ? gy[s] : gy[s - 1] + (delta / dx) * df;
            if (c.eps_schedule == CMCD_EPS_LINEAR) eps[i] = (0.0001f - eps0) / (float)(K - 1) * (float)i + eps0;   // mcd_cais.py:34-36
            else eps[i] = eps0 * ch_eps_decay(c.eps_schedule, i, K);
        }
        return;
    }
    // zero-padded copies of the geffner weights: U1p [d][HP] = W1[:d], U2p = W2[:d], W2p [HP][HP] = W2, W3p [HP][dout] = W3
    if (c.arch == CMCD_ARCH_GEFFNER && U1p) {
        // one row per block iteration: rows [0, d) of U1p, [d, 2d) of U2p, then HP rows of W2p, then HP rows of W3p
        const int nb = gridDim.x - (T + 1), me = b - (T + 1);
        const float* w1 = p + c.off[CMCD_LEAF_GEF_W1];
        const float* w2 = p + c.off[CMCD_LEAF_GEF_W2];
        const float* w3 = p + c.off[CMCD_LEAF_GEF_W3];
        const int nrows = 2 * d + 2 * HP;
        for (int rr = me; rr < nrows; rr += nb) {
            if (rr < d) { for (int j = tid; j < HP; j += blockDim.x) U1p[(size_t)rr * HP + j] = j < H ? w1[(size_t)rr * H + j] : 0.f; }
            else if (rr < 2 * d) { const int r = rr - d; for (int j = tid; j < HP; j += blockDim.x) U2p[(size_t)r * HP + j] = j < H ? w2[(size_t)r * H + j] : 0.f; }
            else if (rr < 2 * d + HP) { const int r = rr - 2 * d; for (int j = tid; j < HP; j += blockDim.x) W2p[(size_t)r * HP + j] = (r < H && j < H) ? w2[(size_t)r * H + j] : 0.f; }
            else { const int r = rr - 2 * d - HP; for (int j = tid; j < dout; j += blockDim.x) W3p[(size_t)r * dout + j] = r < H ? w3[(size_t)r * dout + j] : 0.f; }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------- reverse
struct ChainGrads {
    const float *g_betas, *g_eps, *g_mean, *g_logdiag, *c1, *c2, *c3, *U1, *U2, *U3, *W2, *W3, *os;
};
__device__ __forceinline__ bool ch_train(const cmcd_chain& c, int leaf) { return c.off[leaf] >= 0 && ((c.train_mask >> leaf) & 1u); }

// y[i] = sum_j v[j] W[i][j] for a row-major [rows][64] matrix: one warp per row (lanes over j, coalesced), shuffle reduction in a
// fixed order.  v in shared memory; result to y (shared).  blockDim.x == 128 (4 warps).
__device__ __forceinline__ void ch_rowdot64(const float* __restrict__ W, int rows, const float* v, float* y) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float v0 = v[lane], v1 = v[lane + 32];
    for (int i = warp; i < rows; i += 4) {
        float s = fmaf(v0, W[(size_t)i * CH_C + lane], v1 * W[(size_t)i * CH_C + lane + 32]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) y[i] = s;
    }
}

// scratch layout (dds): per row t: code[128] | h[64] | tnet[64] | g_tnet[64] | g_pre[64] | g_phase[64]
constexpr int CH_ROW = 2 * CH_C + 5 * CH_C;

__global__ void __launch_bounds__(128) chain_bwd_rows_kernel(const ChainView v, const ChainGrads g, float* __restrict__ scratch, float* __restrict__ out) {
    const cmcd_chain& c = v.c;
    const float* __restrict__ p = v.p;
    const int K = c.nbridges, T = K + 1, d = c.in_dim, dout = c.dim, H = c.hidden, HP = c.hidden_pad, tid = threadIdx.x;
    __shared__ float sh[2 * CH_C + 3 * CH_C + 3 * CH_C + 48];
    const int b = blockIdx.x;
    if (b < T && c.arch == CMCD_ARCH_DDS) {
        float *code = sh, *pre = sh + 2 * CH_C, *h = pre + CH_C, *tnet = h + CH_C, *gt = tnet + CH_C, *gp = gt + CH_C, *gc1 = gp + CH_C;
        ch_dds_row_forward(c, p, b, code, pre, h, tnet);
        if (tid < CH_C) gc1[tid] = g.c1[(size_t)b * HP + tid];
        __syncthreads();
        // g_tnet[i] = sum_j g_c1[j] W1[d + i][j]
        ch_rowdot64(p + c.off[CMCD_LEAF_DDS_ST1_W] + (size_t)d * CH_C, CH_C, gc1, gt);
        __syncthreads();
        // g_h[i] = sum_j g_tnet[j] tc2.w[i][j]; g_pre = g_h gelu'(pre)
        ch_rowdot64(p + c.off[CMCD_LEAF_DDS_TC2_W], CH_C, gt, gp);
        __syncthreads();
        if (tid < CH_C) gp[tid] *= ch_gelu_grad(pre[tid]);
        __syncthreads();
        float* row = scratch + (size_t)b * CH_ROW;
        row[tid] = code[tid];
        __syncthreads();
        // g_code[cidx] = sum_j g_pre[j] tc1.w[cidx][j]  (128 rows); overwrites code in shared memory (its copy is in `row`)
        ch_rowdot64(p + c.off[CMCD_LEAF_DDS_TC1_W], 2 * CH_C, gp, code);
        __syncthreads();
        if (tid < CH_C) {
            // d sin(arg)/d phase = cos(arg) = (saved) code[64 + ch]; d cos(arg)/d phase = -sin(arg)
            const float sn = row[tid], cs = row[CH_C + tid];
            row[2 * CH_C + tid] = h[tid];
            row[3 * CH_C + tid] = tnet[tid];
            row[4 * CH_C + tid] = gt[tid];
            row[5 * CH_C + tid] = gp[tid];
            row[6 * CH_C + tid] = code[tid] * cs - code[CH_C + tid] * sn;
        }
        return;
    }
    if (b < T && c.arch == CMCD_ARCH_GEFFNER) {
        // g_emb[r] = sum over the rows t that read emb[r] (t = r, and t = K for r = K - 1) of sum_l g_cl[t] W_l[d:]^T
        if (b == K || !ch_train(c, CMCD_LEAF_GEF_EMB)) return;   // row K is folded into block K - 1
        const int E = c.emb_dim;
        const float* w1 = p + c.off[CMCD_LEAF_GEF_W1] + (size_t)d * H;
        const float* w2 = p + c.off[CMCD_LEAF_GEF_W2] + (size_t)d * H;
        const float* w3 = p + c.off[CMCD_LEAF_GEF_W3] + (size_t)d * dout;
        const int nrow = (b == K - 1) ? 2 : 1;
        const int warp = tid >> 5, lane = tid & 31;
        for (int i = warp; i < E; i += 4) {          // one warp per embedding index: lanes stride over the hidden units (coalesced)
            float s = 0.f;
            for (int rr = 0; rr < nrow; ++rr) {
                const int t = b + rr;
                const float* g1 = g.c1 + (size_t)t * HP;
                const float* g2 = g.c2 + (size_t)t * HP;
                const float* g3 = g.c3 + (size_t)t * dout;
                for (int j = lane; j < H; j += 32) s = fmaf(g1[j], w1[(size_t)i * H + j], fmaf(g2[j], w2[(size_t)i * H + j], s));
                for (int j = lane; j < dout; j += 32) s = fmaf(g3[j], w3[(size_t)i * dout + j], s);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) out[c.off[CMCD_LEAF_GEF_EMB] + (size_t)b * E + i] = s;
        }
        return;
    }
    if (b == T) {
        // transposes of the betas / eps / q prologue
        const int G1 = c.ngrid;
        float* gy = sh;              // forward values
        float* ggy = sh + 48;        // cotangents of gridref_y
        __shared__ float tot_s;
        if (tid == 0) {
            const float* m = p + c.off[CMCD_LEAF_MGRID_Y];
            float tot = 0.f;
            for (int k = 0; k < G1; ++k) tot += m[k];
            float run = 0.f;
            gy[0] = 0.f;
            for (int k = 0; k < G1; ++k) { run += m[k]; gy[k + 1] = run / tot; }
            tot_s = tot;
            for (int k = 0; k <= G1; ++k) ggy[k] = 0.f;
        }
        __syncthreads();
        // per step (in parallel, 128 at a time): segment, interpolation weight, eps-schedule factor; then one thread accumulates
        // the chunk in ascending step order (deterministic)
        __shared__ int seg_s[128];
        __shared__ float w_s[128], gb_s[128], ge_s[128];
        __shared__ float ge_tot;
        if (tid == 0) ge_tot = 0.f;
        for (int i0 = 0; i0 < K; i0 += 128) {
            const int i = i0 + tid;
            if (i < K) {
                const float* xp = p + c.off[CMCD_LEAF_GRID_X];
                const float x = p[c.off[CMCD_LEAF_TARGET_X] + i];
                int sidx = 0;
                while (sidx < G1 + 1 && xp[sidx] <= x) ++sidx;
                sidx = min(max(sidx, 1), G1);
                const float dx = xp[sidx] - xp[sidx - 1], delta = x - xp[sidx - 1];
                seg_s[tid] = sidx;
                w_s[tid] = (dx == 0.f) ? 2.0f : delta / dx;          // 2 = marker for the degenerate segment (whole weight on the right node)
                gb_s[tid] = g.g_betas[i];
                const float dec = (c.eps_schedule == CMCD_EPS_LINEAR) ? 1.0f - (float)i / (float)(K - 1) : ch_eps_decay(c.eps_schedule, i, K);
                ge_s[tid] = g.g_eps[i] * dec;
            }
            __syncthreads();
            if (tid == 0) {
                float ge = ge_tot;
                const int n = min(128, K - i0);
                for (int q = 0; q < n; ++q) {
                    const int sidx = seg_s[q];
                    const float w = w_s[q], gb = gb_s[q];
                    if (w == 2.0f) ggy[sidx] += gb;
                    else { ggy[sidx] += w * gb; ggy[sidx - 1] += (1.0f - w) * gb; }
                    ge += ge_s[q];
                }
                ge_tot = ge;
            }
            __syncthreads();
        }
        if (tid == 0 && K >= 1) {
            const float ge = ge_tot;
            if (ch_train(c, CMCD_LEAF_EPS)) out[c.off[CMCD_LEAF_EPS]] = ge;
            if (ch_train(c, CMCD_LEAF_MGRID_Y)) {
                // gy[k] = C[k] / tot (k >= 1): g_C[k] = ggy[k] / tot, g_tot = -sum_k ggy[k] gy[k] / tot; m[l] feeds C[k] for k > l and tot
                const float tot = tot_s;
                float gtot = 0.f;
                for (int k = 1; k <= G1; ++k) gtot -= ggy[k] * gy[k] / tot;
                float suffix = 0.f;
                for (int l = G1 - 1; l >= 0; --l) { suffix += ggy[l + 1] / tot; out[c.off[CMCD_LEAF_MGRID_Y] + l] = suffix + gtot; }
            }
        }
        if (tid < dout) {
            if (ch_train(c, CMCD_LEAF_VD_MEAN)) out[c.off[CMCD_LEAF_VD_MEAN] + tid] = g.g_mean[tid];
            if (ch_train(c, CMCD_LEAF_VD_LOGDIAG)) out[c.off[CMCD_LEAF_VD_LOGDIAG] + tid] = g.g_logdiag[tid];
        }
        for (int j = tid + blockDim.x; j < dout; j += blockDim.x) {   // d > 128 (lgcp)
            if (ch_train(c, CMCD_LEAF_VD_MEAN)) out[c.off[CMCD_LEAF_VD_MEAN] + j] = g.g_mean[j];
            if (ch_train(c, CMCD_LEAF_VD_LOGDIAG)) out[c.off[CMCD_LEAF_VD_LOGDIAG] + j] = g.g_logdiag[j];
        }
        if (tid == 0 && c.arch == CMCD_ARCH_GEFFNER && ch_train(c, CMCD_LEAF_GEF_FACTOR) && g.os) out[c.off[CMCD_LEAF_GEF_FACTOR]] = g.os[0];
    }
}

// one thread per weight element; sums over the T rows in ascending order
#define CH_PUT(leaf, k, val) do { if ((c.train_mask >> (leaf)) & 1u) out[c.off[leaf] + (k)] = (val); } while (0)
__global__ void __launch_bounds__(256) chain_bwd_weights_kernel(const ChainView v, const ChainGrads g, const float* __restrict__ scratch,
                                                                float* __restrict__ out, long long total) {
    const cmcd_chain& c = v.c;
    const float* __restrict__ p = v.p;
    const int K = c.nbridges, T = K + 1, d = c.in_dim, dout = c.dim, H = c.hidden, HP = c.hidden_pad;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        int k = (int)idx;   // total < 2^31 (checked by the launcher): 32-bit divisions
        if (c.arch == CMCD_ARCH_DDS) {
            const int C = CH_C;
            // segments: tc1.w [128][64] | tc1.b | tc2.w [64][64] | tc2.b | st1.w [(d+64)][64] | st1.b | st2.w | st2.b | out.w [64][dout] | out.b | phase
            if (k < 2 * C * C) {            // tc1.w[i][j] = sum_t code[t][i] g_pre[t][j]
                const int i = (int)(k / C), j = (int)(k % C);
                float s = 0.f;
                for (int t = 0; t < T; ++t) s = fmaf(scratch[(size_t)t * CH_ROW + i], scratch[(size_t)t * CH_ROW + 5 * C + j], s);
                CH_PUT(CMCD_LEAF_DDS_TC1_W, k, s);
                continue;
            }
            k -= 2 * C * C;
            if (k < C) { float s = 0.f; for (int t = 0; t < T; ++t) s += scratch[(size_t)t * CH_ROW + 5 * C + k]; CH_PUT(CMCD_LEAF_DDS_TC1_B, k, s); continue; }
            k -= C;
            if (k < C * C) {                // tc2.w[i][j] = sum_t h[t][i] g_tnet[t][j]
                const int i = (int)(k / C), j = (int)(k % C);
                float s = 0.f;
                for (int t = 0; t < T; ++t) s = fmaf(scratch[(size_t)t * CH_ROW + 2 * C + i], scratch[(size_t)t * CH_ROW + 4 * C + j], s);
                CH_PUT(CMCD_LEAF_DDS_TC2_W, k, s);
                continue;
            }
            k -= C * C;
            if (k < C) { float s = 0.f; for (int t = 0; t < T; ++t) s += scratch[(size_t)t * CH_ROW + 4 * C + k]; CH_PUT(CMCD_LEAF_DDS_TC2_B, k, s); continue; }
            k -= C;
            if (k < (d + C) * C) {   // st1.w: rows < d = U1 cotangent; rows >= d: sum_t tnet[t][i] g_c1[t][j]
                const int r = (int)(k / C), j = (int)(k % C);
                float s;
                if (r < d) s = g.U1[(size_t)r * HP + j];
                else {
                    s = 0.f;
                    for (int t = 0; t < T; ++t) s = fmaf(scratch[(size_t)t * CH_ROW + 3 * C + (r - d)], g.c1[(size_t)t * HP + j], s);
                }
                CH_PUT(CMCD_LEAF_DDS_ST1_W, k, s);
                continue;
            }
            k -= (d + C) * C;
            if (k < C) { float s = 0.f; for (int t = 0; t < T; ++t) s += g.c1[(size_t)t * HP + k]; CH_PUT(CMCD_LEAF_DDS_ST1_B, k, s); continue; }
            k -= C;
            if (k < C * C) { CH_PUT(CMCD_LEAF_DDS_ST2_W, k, g.W2[k]); continue; }
            k -= C * C;
            if (k < C) { float s = 0.f; for (int t = 0; t < T; ++t) s += g.c2[(size_t)t * HP + k]; CH_PUT(CMCD_LEAF_DDS_ST2_B, k, s); continue; }
            k -= C;
            if (k < C * dout) { CH_PUT(CMCD_LEAF_DDS_OUT_W, k, g.W3[k]); continue; }
            k -= C * dout;
            if (k < dout) { float s = 0.f; for (int t = 0; t < T; ++t) s += g.c3[(size_t)t * dout + k]; CH_PUT(CMCD_LEAF_DDS_OUT_B, k, s); continue; }
            k -= dout;
            if (k < C) { float s = 0.f; for (int t = 0; t < T; ++t) s += scratch[(size_t)t * CH_ROW + 6 * C + k]; CH_PUT(CMCD_LEAF_DDS_PHASE, k, s); continue; }
        } else if (c.arch == CMCD_ARCH_GEFFNER) {
            const int E = c.emb_dim, in = d + E;   // = H
            // segments: W1 [in][H] | b1 [H] | W2 [in][H] | b2 | W3 [in][dout] | b3
            auto emb_row = [&](int t) { return p + c.off[CMCD_LEAF_GEF_EMB] + (size_t)(t < K ? t : K - 1) * E; };
            if (k < in * H) {     // W1: rows < d = U1 cotangent; rows >= d: sum_t e[t][i] g_c1[t][j]
                const int r = (int)(k / H), j = (int)(k % H);
                float s;
                if (r < d) s = g.U1[(size_t)r * HP + j];
                else { s = 0.f; for (int t = 0; t < T; ++t) s = fmaf(emb_row(t)[r - d], g.c1[(size_t)t * HP + j], s); }
                CH_PUT(CMCD_LEAF_GEF_W1, k, s);
                continue;
            }
            k -= in * H;
            if (k < H) { float s = 0.f; for (int t = 0; t < T; ++t) s += g.c1[(size_t)t * HP + k]; CH_PUT(CMCD_LEAF_GEF_B1, k, s); continue; }
            k -= H;
            if (k < in * H) {     // W2: acts on a1 (all rows: W2 cotangent) + on x through U2 (rows < d) + on emb through c2 (rows >= d)
                const int r = (int)(k / H), j = (int)(k % H);
                float s = g.W2[(size_t)r * HP + j];
                if (r < d) s += g.U2 ? g.U2[(size_t)r * HP + j] : 0.f;
                else { float a = 0.f; for (int t = 0; t < T; ++t) a = fmaf(emb_row(t)[r - d], g.c2[(size_t)t * HP + j], a); s += a; }
                CH_PUT(CMCD_LEAF_GEF_W2, k, s);
                continue;
            }
            k -= in * H;
            if (k < H) { float s = 0.f; for (int t = 0; t < T; ++t) s += g.c2[(size_t)t * HP + k]; CH_PUT(CMCD_LEAF_GEF_B2, k, s); continue; }
            k -= H;
            if (k < in * dout) {
                const int r = (int)(k / dout), j = (int)(k % dout);
                float s = g.W3[(size_t)r * dout + j];
                if (r < d) s += g.U3 ? g.U3[(size_t)r * dout + j] : 0.f;
                else { float a = 0.f; for (int t = 0; t < T; ++t) a = fmaf(emb_row(t)[r - d], g.c3[(size_t)t * dout + j], a); s += a; }
                CH_PUT(CMCD_LEAF_GEF_W3, k, s);
                continue;
            }
            k -= in * dout;
            if (k < dout) { float s = 0.f; for (int t = 0; t < T; ++t) s += g.c3[(size_t)t * dout + k]; CH_PUT(CMCD_LEAF_GEF_B3, k, s); continue; }
        }
    }
}

static long long chain_weight_elements(const cmcd_chain& c) {
    const long long d = c.in_dim, dout = c.dim, H = c.hidden, C = CH_C;
    if (c.arch == CMCD_ARCH_DDS) return 2 * C * C + C + C * C + C + (d + C) * C + C + C * C + C + C * dout + dout + C;
    if (c.arch == CMCD_ARCH_GEFFNER) return 2 * (H * H + H) + H * dout + dout;
    return 0;
}

static int chain_check(const cmcd_chain* c) {
    if (!c) { set_error("chain: null descriptor"); return 2; }
    if (c->arch != CMCD_ARCH_NONE && c->arch != CMCD_ARCH_GEFFNER && c->arch != CMCD_ARCH_DDS) { set_error("chain: nn_arch %d not implemented", c->arch); return 2; }
    if (c->nbridges < 1 || c->ngrid < 1 || c->ngrid > 39) { set_error("chain: needs nbridges >= 1 and 1 <= len(mgridref_y) <= 39 (got %d, %d)", c->nbridges, c->ngrid); return 2; }
    if (c->arch == CMCD_ARCH_DDS && (c->hidden != CH_C || c->hidden_pad != CH_C || !c->dds_coeff)) { set_error("chain: dds needs hidden = 64 and the timestep coefficients"); return 2; }
    if (c->arch == CMCD_ARCH_GEFFNER && c->hidden != c->in_dim + c->emb_dim) { set_error("chain: geffner hidden must equal in_dim + emb_dim"); return 2; }
    for (int l : {CMCD_LEAF_EPS, CMCD_LEAF_MGRID_Y, CMCD_LEAF_GRID_X, CMCD_LEAF_TARGET_X})
        if (c->off[l] < 0) { set_error("chain: leaf %d missing from params_flat", l); return 2; }
    return 0;
}

int launch_chain_fwd(const cmcd_chain* c, cudaStream_t st, const float* params_flat, float* betas, float* eps, float* c1, float* c2, float* c3,
                     float* U1p, float* U2p, float* W2p, float* W3p) {
    if (int rc = chain_check(c)) return rc;
    if (!params_flat || !betas || !eps) { set_error("chain_fwd: null buffer"); return 2; }
    if (c->arch != CMCD_ARCH_NONE && (!c1 || !c2 || !c3)) { set_error("chain_fwd: table buffers missing"); return 2; }
    ChainView v; v.c = *c; v.p = params_flat;
    const int T = c->nbridges + 1;
    int grid = T + 1;
    if (c->arch == CMCD_ARCH_GEFFNER && U1p) {
        int nb = 2 * c->in_dim + 2 * c->hidden_pad;   // one padded row per block (capped: the blocks stride over the rows)
        if (nb > 4096) nb = 4096;
        grid += nb;
    }
    chain_fwd_kernel<<<grid, 128, 0, st>>>(v, betas, eps, c1, c2, c3, U1p, U2p, W2p, W3p);
    CMCD_CUDA_OK(cudaGetLastError());
    return 0;
}

size_t chain_bwd_scratch_floats(const cmcd_chain* c) {
    if (!c || c->arch != CMCD_ARCH_DDS) return 16;
    return (size_t)(c->nbridges + 1) * CH_ROW;
}

int launch_chain_bwd(const cmcd_chain* c, cudaStream_t st, const float* params_flat, const float* g_betas, const float* g_eps,
                     const float* g_vd_mean, const float* g_vd_logdiag, const cmcd_net_grad* gn, float* scratch, size_t scratch_floats,
                     float* grad_flat) {
    if (int rc = chain_check(c)) return rc;
    if (!params_flat || !g_betas || !g_eps || !g_vd_mean || !g_vd_logdiag || !grad_flat) { set_error("chain_bwd: null buffer"); return 2; }
    if (c->arch != CMCD_ARCH_NONE && (!gn || !gn->c1 || !gn->c2 || !gn->c3 || !gn->U1 || !gn->W2 || !gn->W3)) { set_error("chain_bwd: network cotangents missing"); return 2; }
    if (scratch_floats < chain_bwd_scratch_floats(c) || !scratch) { set_error("chain_bwd: scratch too small"); return 2; }
    ChainView v; v.c = *c; v.p = params_flat;
    ChainGrads g{};
    g.g_betas = g_betas; g.g_eps = g_eps; g.g_mean = g_vd_mean; g.g_logdiag = g_vd_logdiag;
    if (gn) { g.c1 = gn->c1; g.c2 = gn->c2; g.c3 = gn->c3; g.U1 = gn->U1; g.U2 = gn->U2; g.U3 = gn->U3; g.W2 = gn->W2; g.W3 = gn->W3; g.os = gn->out_scale; }
    CMCD_CUDA_OK(cudaMemsetAsync(grad_flat, 0, (size_t)c->n_params * sizeof(float), st));
    const int T = c->nbridges + 1;
    chain_bwd_rows_kernel<<<T + 1, 128, 0, st>>>(v, g, scratch, grad_flat);
    CMCD_CUDA_OK(cudaGetLastError());
    const long long total = chain_weight_elements(*c);
    if (total >= (1LL << 31)) { set_error("chain_bwd: network too large (%lld weight elements)", total); return 2; }
    if (total > 0) {
        long long nb = (total + 255) / 256;
        if (nb > 4096) nb = 4096;
        chain_bwd_weights_kernel<<<(unsigned)nb, 256, 0, st>>>(v, g, scratch, grad_flat, total);
        CMCD_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

}  // namespace cmcd
