// Network recompute + block-cooperative pull-back shared by the FP32-FMA adjoint kernels (bridge_bwd.cu: overdamped
// modes; bridge_ud.cu: underdamped lp_a family).  DI = network input width (D, or 2D for the (z, rho) networks).
#pragma once
#include "net.cuh"

// A/B on B200 (tools/ab.sh): inlining the two network functions into the adjoint kernel is 21% faster than
// calling them (the 255-register kernel pays for the ABI spills); the forward kernel prefers the call.
#ifdef CMCD_NOINLINE_NET_BWD
#define CMCD_NETB_INL __noinline__
#else
#define CMCD_NETB_INL __forceinline__
#endif

namespace cmcd {

struct BwdLayout {  // offsets (floats) into one block's partial-gradient slice
    int W2, U1, U2, U3, W3, c1, c2, c3, os, beta, eps, mu, ls, P;
};

static BwdLayout make_layout(int D, int K, int HP, int arch, int DI = 0) {
    BwdLayout l;
    int o = 0;
    if (DI == 0) DI = D;
    const int T = K + 1;
    const bool net = arch != CMCD_ARCH_NONE;
    l.W2 = o; o += net ? HP * HP : 0;
    l.U1 = o; o += net ? DI * HP : 0;
    l.U2 = o; o += net ? DI * HP : 0;
    l.U3 = o; o += net ? DI * D : 0;
    l.W3 = o; o += net ? HP * D : 0;
    l.c1 = o; o += net ? T * HP : 0;
    l.c2 = o; o += net ? T * HP : 0;
    l.c3 = o; o += net ? T * D : 0;
    l.os = o; o += 1;
    l.beta = o; o += K > 0 ? K : 1;
    l.eps = o; o += K > 0 ? K : 1;
    l.mu = o; o += D;
    l.ls = o; o += D;
    l.P = (o + 3) & ~3;
    return l;
}

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- network forward with stored activations ------------------------------------------------
// S1 <- a1, S2 <- a2, S3 <- act'(pre2).  Returns raw o (before clamp / out_scale).
template <int D, int ACT, int HPT, int JC, int RS, int DI = D>
__device__ CMCD_NETB_INL void net_fwd_store(const NetView& nv, const NetSmem& s, int t, const float* __restrict__ xin,
                                           float* __restrict__ oout, float* __restrict__ S1c, float* __restrict__ S2c,
                                           float* __restrict__ S3c) {
    const int HP = HPT ? HPT : nv.HP;
    float x[DI], o[D];
#pragma unroll
    for (int a = 0; a < DI; ++a) x[a] = xin[a];
    const float* __restrict__ c1 = nv.c1 + (size_t)t * HP;
    const float* __restrict__ c2 = nv.c2 + (size_t)t * HP;
    const float* __restrict__ c3 = nv.c3 + (size_t)t * D;
    constexpr bool has_u2 = (ACT == ACT_SOFTPLUS), has_u3 = (ACT == ACT_SOFTPLUS);
    constexpr float skip = (ACT == ACT_SOFTPLUS) ? 1.f : 0.f;
#pragma unroll 4
    for (int j = 0; j < HP; ++j) {
        float p = __ldg(c1 + j);
#pragma unroll
        for (int a = 0; a < DI; ++a) p = fmaf(x[a], s.U1[a * HP + j], p);
        S1c[j * RS] = act_fwd<ACT>(p);
    }
#pragma unroll
    for (int m = 0; m < D; ++m) {
        float p = __ldg(c3 + m);
        if (has_u3) {
#pragma unroll
            for (int a = 0; a < DI; ++a) p = fmaf(x[a], s.U3[a * D + m], p);
        }
        o[m] = p;
    }
#pragma unroll 1
    for (int j0 = 0; j0 < HP; j0 += JC) {
        float acc[JC];
#pragma unroll
        for (int jj = 0; jj < JC; ++jj) {
            float p = __ldg(c2 + j0 + jj);
            if (has_u2) {
#pragma unroll
                for (int a = 0; a < DI; ++a) p = fmaf(x[a], s.U2[a * HP + j0 + jj], p);
            }
            acc[jj] = p;
        }
#pragma unroll (inner_unroll(JC))
        for (int i = 0; i < HP; ++i) {
            const float h = S1c[i * RS];
            const float4* __restrict__ w = reinterpret_cast<const float4*>(s.W2 + (size_t)i * HP + j0);
#pragma unroll
            for (int q = 0; q < JC / 4; ++q) {
                const float4 ww = w[q];
                acc[4 * q + 0] = fmaf(h, ww.x, acc[4 * q + 0]);
                acc[4 * q + 1] = fmaf(h, ww.y, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(h, ww.z, acc[4 * q + 2]);
                acc[4 * q + 3] = fmaf(h, ww.w, acc[4 * q + 3]);
            }
        }
        // park the pre-activations, then a rolled activation / layer-3 loop (compact code)
#pragma unroll
        for (int jj = 0; jj < JC; ++jj) S2c[(j0 + jj) * RS] = acc[jj];
#pragma unroll 4
        for (int jj = 0; jj < JC; ++jj) {
            float a2, da2;
            act_fwd_grad<ACT>(S2c[(j0 + jj) * RS], a2, da2);
            S2c[(j0 + jj) * RS] = a2;
            S3c[(j0 + jj) * RS] = da2;
            const float hs = a2 + skip * S1c[(j0 + jj) * RS];
#pragma unroll
            for (int m = 0; m < D; ++m) o[m] = fmaf(hs, s.W3[(j0 + jj) * D + m], o[m]);
        }
    }
#pragma unroll
    for (int m = 0; m < D; ++m) oout[m] = o[m];
}

// ---- network backward (block-cooperative) ----------------------------------------------------
// v: cotangent on the network output (per particle).  Returns dx = J_x^T v and accumulates the
// parameter cotangents into this block's partial slice.  Contains __syncthreads(): every thread of
// the block must call it (inactive particles pass v = 0).
template <int D, int ACT, int HPT, int JC, int BPB, int DI = D>
__device__ CMCD_NETB_INL void net_bwd(const NetView& nv, const NetSmem& s, int t, const float* __restrict__ xin,
                                     const float* __restrict__ oin, const float* __restrict__ vin, float* __restrict__ dxout,
                                     float* __restrict__ S1, float* __restrict__ S2, float* __restrict__ S3,
                                     float* __restrict__ sX, float* __restrict__ sVo,
                                     float* __restrict__ part, const BwdLayout& L) {
    constexpr int RS = BPB + 4;
    float x[DI], o[D], v[D], dx[DI];
#pragma unroll
    for (int a = 0; a < DI; ++a) x[a] = xin[a];
#pragma unroll
    for (int a = 0; a < D; ++a) { o[a] = oin[a]; v[a] = vin[a]; }
    const int HP = HPT ? HPT : nv.HP;
    const int tid = threadIdx.x;
    constexpr float skip = (ACT == ACT_SOFTPLUS) ? 1.f : 0.f;
    constexpr bool has_u2 = (ACT == ACT_SOFTPLUS), has_u3 = (ACT == ACT_SOFTPLUS);
    float* S2c = S2 + tid; float* S3c = S3 + tid;

    // (1) private: output layer cotangent, dp2 = W3 vo * act'(pre2) -> S3
    float vo[D];
    float gos = 0.f;
#pragma unroll
    for (int m = 0; m < D; ++m) {
        const float oc = fminf(fmaxf(o[m], -nv.out_clip), nv.out_clip);
        gos = fmaf(v[m], oc, gos);
        vo[m] = (fabsf(o[m]) <= nv.out_clip) ? v[m] * net_out_scale(nv) : 0.f;
        sVo[m * RS + tid] = vo[m];
    }
#pragma unroll
    for (int a = 0; a < DI; ++a) sX[a * RS + tid] = x[a];
#pragma unroll 4
    for (int j = 0; j < HP; ++j) {
        float d2 = 0.f;
#pragma unroll
        for (int m = 0; m < D; ++m) d2 = fmaf(s.W3[j * D + m], vo[m], d2);
        S3c[j * RS] = d2 * S3c[j * RS];
    }
    gos = warp_sum_f(gos);
    if ((tid & 31) == 0 && gos != 0.f) atomicAdd(part + L.os, gos);
    __syncthreads();

    // (2) cooperative: gW2 += a1^T dp2 (4x4 register tiles over interleaved rows), then the
    //     skinny products gc2, gU2, gW3, gc3, gU3.
    {
        const int G = HP / 4;  // tile grid edge; rows of tile (ti,.) are ti + G*r
        for (int tile = tid; tile < G * G; tile += BPB) {
            const int ti = tile / G, tj = tile % G;
            float acc[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[r][q] = 0.f;
            for (int p = 0; p < BPB; p += 4) {
                float4 A[4], B[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    A[r] = *reinterpret_cast<const float4*>(S1 + (ti + G * r) * RS + p);
                    B[r] = *reinterpret_cast<const float4*>(S3 + (tj + G * r) * RS + p);
                }
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        acc[r][q] = fmaf(A[r].x, B[q].x, acc[r][q]);
                        acc[r][q] = fmaf(A[r].y, B[q].y, acc[r][q]);
                        acc[r][q] = fmaf(A[r].z, B[q].z, acc[r][q]);
                        acc[r][q] = fmaf(A[r].w, B[q].w, acc[r][q]);
                    }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    atomicAdd(part + L.W2 + (ti + G * r) * HP + (tj + G * q), acc[r][q]);
        }
        // skinny: one (row j, particle-slice) per thread
        int nparts = 1;                       // power-of-two particle slices per row (slices stay float4 aligned)
        while (nparts * 2 * HP <= BPB) nparts *= 2;
        const int plen = BPB / nparts;
        for (int job = tid; job < HP * nparts; job += BPB) {
            const int j = job % HP, pp = job / HP;
            float s2 = 0.f, gu2[DI], gw3[D];
#pragma unroll
            for (int a = 0; a < DI; ++a) gu2[a] = 0.f;
#pragma unroll
            for (int a = 0; a < D; ++a) gw3[a] = 0.f;
            for (int p = pp * plen; p < (pp + 1) * plen; p += 4) {
                const float4 d2 = *reinterpret_cast<const float4*>(S3 + j * RS + p);
                const float4 a2 = *reinterpret_cast<const float4*>(S2 + j * RS + p);
                const float4 a1 = *reinterpret_cast<const float4*>(S1 + j * RS + p);
                s2 += (d2.x + d2.y) + (d2.z + d2.w);
                const float h0 = a2.x + skip * a1.x, h1 = a2.y + skip * a1.y, h2 = a2.z + skip * a1.z, h3 = a2.w + skip * a1.w;
                if (has_u2) {
#pragma unroll
                    for (int a = 0; a < DI; ++a) {
                        const float4 xx = *reinterpret_cast<const float4*>(sX + a * RS + p);
                        gu2[a] += d2.x * xx.x + d2.y * xx.y + d2.z * xx.z + d2.w * xx.w;
                    }
                }
#pragma unroll
                for (int a = 0; a < D; ++a) {
                    const float4 vv = *reinterpret_cast<const float4*>(sVo + a * RS + p);
                    gw3[a] += h0 * vv.x + h1 * vv.y + h2 * vv.z + h3 * vv.w;
                }
            }
            atomicAdd(part + L.c2 + (size_t)t * HP + j, s2);
            if (has_u2) {
#pragma unroll
                for (int a = 0; a < DI; ++a) atomicAdd(part + L.U2 + a * HP + j, gu2[a]);
            }
#pragma unroll
            for (int a = 0; a < D; ++a) atomicAdd(part + L.W3 + j * D + a, gw3[a]);
        }
        // gc3[t][m] = sum_p vo[m][p];  gU3[a][m] = sum_p x[a][p] vo[m][p]
        for (int job = tid; job < D + (has_u3 ? DI * D : 0); job += BPB) {
            float sacc = 0.f;
            if (job < D) {
                for (int p = 0; p < BPB; ++p) sacc += sVo[job * RS + p];
                atomicAdd(part + L.c3 + (size_t)t * D + job, sacc);
            } else {
                const int a = (job - D) / D, m = (job - D) % D;
                for (int p = 0; p < BPB; ++p) sacc = fmaf(sX[a * RS + p], sVo[m * RS + p], sacc);
                atomicAdd(part + L.U3 + a * D + m, sacc);
            }
        }
    }
    __syncthreads();

    // (3) private: da1 = W2 dp2 + skip * W3 vo ; dp1 = da1 * act'(pre1) -> S2
    {
        const float* __restrict__ c1 = nv.c1 + (size_t)t * HP;
        for (int j0 = 0; j0 < HP; j0 += JC) {
            float dreg[JC];
#pragma unroll
            for (int jj = 0; jj < JC; ++jj) dreg[jj] = S3c[(j0 + jj) * RS];
#pragma unroll (inner_unroll(JC))
            for (int i = 0; i < HP; ++i) {
                const float4* __restrict__ w = reinterpret_cast<const float4*>(s.W2 + (size_t)i * HP + j0);
                float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
#pragma unroll
                for (int q = 0; q < JC / 4; ++q) {
                    const float4 ww = w[q];
                    p0 = fmaf(ww.x, dreg[4 * q + 0], p0);
                    p1 = fmaf(ww.y, dreg[4 * q + 1], p1);
                    p2 = fmaf(ww.z, dreg[4 * q + 2], p2);
                    p3 = fmaf(ww.w, dreg[4 * q + 3], p3);
                }
                const float part_sum = (p0 + p1) + (p2 + p3);
                if (j0 == 0) S2c[i * RS] = part_sum; else S2c[i * RS] += part_sum;
            }
        }
#pragma unroll 2
        for (int i = 0; i < HP; ++i) {
            float da1 = S2c[i * RS];
            if (skip != 0.f) {
#pragma unroll
                for (int m = 0; m < D; ++m) da1 = fmaf(s.W3[i * D + m], vo[m], da1);
            }
            float p = __ldg(c1 + i);
#pragma unroll
            for (int a = 0; a < DI; ++a) p = fmaf(x[a], s.U1[a * HP + i], p);
            float a1, g1;
            act_fwd_grad<ACT>(p, a1, g1);
            S2c[i * RS] = da1 * g1;
        }
    }
    __syncthreads();

    // (4) cooperative: gc1[t] += colsum(dp1), gU1 += x^T dp1 ; private: dx
    {
        int nparts = 1;                       // power-of-two particle slices per row (slices stay float4 aligned)
        while (nparts * 2 * HP <= BPB) nparts *= 2;
        const int plen = BPB / nparts;
        for (int job = tid; job < HP * nparts; job += BPB) {
            const int j = job % HP, pp = job / HP;
            float s1 = 0.f, gu1[DI];
#pragma unroll
            for (int a = 0; a < DI; ++a) gu1[a] = 0.f;
            for (int p = pp * plen; p < (pp + 1) * plen; p += 4) {
                const float4 d1 = *reinterpret_cast<const float4*>(S2 + j * RS + p);
                s1 += (d1.x + d1.y) + (d1.z + d1.w);
#pragma unroll
                for (int a = 0; a < DI; ++a) {
                    const float4 xx = *reinterpret_cast<const float4*>(sX + a * RS + p);
                    gu1[a] += d1.x * xx.x + d1.y * xx.y + d1.z * xx.z + d1.w * xx.w;
                }
            }
            atomicAdd(part + L.c1 + (size_t)t * HP + j, s1);
#pragma unroll
            for (int a = 0; a < DI; ++a) atomicAdd(part + L.U1 + a * HP + j, gu1[a]);
        }
#pragma unroll
        for (int a = 0; a < DI; ++a) {
            float acc = 0.f;
            if (has_u3) {
#pragma unroll
                for (int m = 0; m < D; ++m) acc = fmaf(s.U3[a * D + m], vo[m], acc);
            }
            dx[a] = acc;
        }
#pragma unroll 4
        for (int j = 0; j < HP; ++j) {
            const float d1 = S2c[j * RS], d2 = S3c[j * RS];
#pragma unroll
            for (int a = 0; a < DI; ++a) {
                dx[a] = fmaf(s.U1[a * HP + j], d1, dx[a]);
                if (has_u2) dx[a] = fmaf(s.U2[a * HP + j], d2, dx[a]);
            }
        }
    }
#pragma unroll
    for (int a = 0; a < DI; ++a) dxout[a] = dx[a];
    __syncthreads();
}

// ---- underdamped operators (bridge_ud.cu, bridge_blk.cu): the eps slot of the layout holds the coefficient rows
constexpr int UD_ROWS = 7;   // rows of the eps table: eps, a_f, s_f, a_b, c_n, s_b, c_f
static inline int ud_net_in(int mode, int D) {
    return (mode == CMCD_MODE_UD_NET_ZRHO || mode == CMCD_MODE_UD_CAIS) ? 2 * D : (mode == CMCD_MODE_UD_NET_Z ? D : 0);
}
static inline BwdLayout ud_layout(int D, int K, int HP, int arch, int din) {
    // the eps slot holds the UD_ROWS coefficient rows
    BwdLayout l = make_layout(D, K, HP, arch, din ? din : D);
    const int extra = (UD_ROWS - 1) * (K > 0 ? K : 1);
    l.mu += extra; l.ls += extra; l.P = (l.ls + D + 3) & ~3;
    return l;
}

// out[k] = sum_b partials[b][k], scattered into the caller's cotangent buffers (bridge_bwd.cu)
struct BwdOut {
    float *W2, *U1, *U2, *U3, *W3, *c1, *c2, *c3, *os, *beta, *eps, *mu, *ls;
};
int launch_bwd_reduce(const float* partials, int nblocks, const BwdLayout& L, const BwdOut& out, int HP, int D, int K, cudaStream_t st);

}  // namespace cmcd
