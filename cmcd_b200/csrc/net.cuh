// Per-particle drift-network evaluation ("table form", see include/cmcd_b200.h cmcd_net).
//
// Replaces apply_fun_sn(params["sn"], x, t): src/nn.py:66-70 (geffner) and
// src/nn_dds.py:145-164 (PISNet).  One thread owns one particle; its layer-1 activations live
// in a private shared-memory column (stride = particles per block), the weights are read as
// warp-uniform (broadcast) float4 from shared memory, layer 2 is accumulated in registers in
// chunks of JC output units and layer 3 is folded into the chunk epilogue.
#pragma once
#include "common.cuh"

// Unroll depth of the layer-2 inner loops (one private LDS + JC/4 broadcast LDS.128 + JC FMAs per iteration): the narrow
// register chunks (JC = 8) need several iterations in flight to cover the shared-memory latency when few warps are resident.
#ifndef CMCD_UNROLL_NARROW
#define CMCD_UNROLL_NARROW 2
#endif

namespace cmcd {

__host__ __device__ constexpr int inner_unroll(int jc) { return jc >= 32 ? 2 : CMCD_UNROLL_NARROW; }

// Shared-memory carve-up of the network weights (all HP-padded).
struct NetSmem {
    const float *W2, *U1, *U2, *W3, *U3;  // shared memory copies
};

// DI = input width of the network (D for the overdamped modes; 2D for the (z, rho) networks of the underdamped modes,
// mcdboundingmachine.py:84-102), D = output width.
__host__ __device__ inline size_t net_smem_floats(int D, int HP, int DI = 0) {
    if (DI == 0) DI = D;
    return (size_t)HP * HP + 2 * (size_t)DI * HP + (size_t)HP * D + (size_t)DI * D;
}

// cooperative copy global -> shared; call from all threads, followed by __syncthreads()
__device__ inline NetSmem net_stage_smem(const NetView& nv, int D, float* sm, int DI = 0) {
    if (DI == 0) DI = D;
    const int HP = nv.HP;
    float* sW2 = sm;
    float* sU1 = sW2 + (size_t)HP * HP;
    float* sU2 = sU1 + DI * HP;
    float* sW3 = sU2 + DI * HP;
    float* sU3 = sW3 + HP * D;
    if (nv.arch != CMCD_ARCH_NONE) {
        for (int i = threadIdx.x; i < HP * HP; i += blockDim.x) sW2[i] = nv.W2[i];
        for (int i = threadIdx.x; i < DI * HP; i += blockDim.x) {
            sU1[i] = nv.U1[i];
            sU2[i] = nv.U2 ? nv.U2[i] : 0.f;
        }
        for (int i = threadIdx.x; i < HP * D; i += blockDim.x) sW3[i] = nv.W3[i];
        for (int i = threadIdx.x; i < DI * D; i += blockDim.x) sU3[i] = nv.U3 ? nv.U3[i] : 0.f;
    }
    NetSmem s;
    s.W2 = sW2; s.U1 = sU1; s.U2 = sU2; s.W3 = sW3; s.U3 = sU3;
    return s;
}

// out = NN(x, t).  a1col: this thread's private activation column (element j at a1col[j*PBS]).
// __noinline__: the two evaluations per bridge step share one copy of the code (the fully inlined kernel was
// 146 KB of SASS and stalled on instruction fetch, profiles/r1_ncu_summary.md).
template <int D, int ACT, int HPT, int JC, int PBS, int DI = D>
__device__ __noinline__ void net_fwd(const NetView& nv, const NetSmem& s, int t, const float* __restrict__ x,
                                     float* __restrict__ out, float* __restrict__ a1col) {
    const int HP = HPT ? HPT : nv.HP;
    const float* __restrict__ c1 = nv.c1 + (size_t)t * HP;
    const float* __restrict__ c2 = nv.c2 + (size_t)t * HP;
    const float* __restrict__ c3 = nv.c3 + (size_t)t * D;
    // geffner <=> softplus, residual skip, U2/U3 present; dds <=> gelu, no skip, U2 = U3 = 0 (compile-time)
    constexpr bool has_u2 = (ACT == ACT_SOFTPLUS), has_u3 = (ACT == ACT_SOFTPLUS);
    constexpr float skip = (ACT == ACT_SOFTPLUS) ? 1.f : 0.f;
    float xr[DI];
#pragma unroll
    for (int a = 0; a < DI; ++a) xr[a] = x[a];
    // layer 1
#pragma unroll 4
    for (int j = 0; j < HP; ++j) {
        float p = __ldg(c1 + j);
#pragma unroll
        for (int a = 0; a < DI; ++a) p = fmaf(xr[a], s.U1[a * HP + j], p);
        a1col[j * PBS] = act_fwd<ACT>(p);
    }
    float o[D];
#pragma unroll
    for (int m = 0; m < D; ++m) {
        float p = __ldg(c3 + m);
        if (has_u3) {
#pragma unroll
            for (int a = 0; a < DI; ++a) p = fmaf(xr[a], s.U3[a * D + m], p);
        }
        o[m] = p;
    }
    // layer 2 in chunks of JC output units, layer 3 folded in
#pragma unroll 1
    for (int j0 = 0; j0 < HP; j0 += JC) {
        float acc[JC];
#pragma unroll
        for (int jj = 0; jj < JC; ++jj) {
            float p = __ldg(c2 + j0 + jj);
            if (has_u2) {
#pragma unroll
                for (int a = 0; a < DI; ++a) p = fmaf(xr[a], s.U2[a * HP + j0 + jj], p);
            }
            acc[jj] = p;
        }
#pragma unroll (inner_unroll(JC))
        for (int i = 0; i < HP; ++i) {
            const float h = a1col[i * PBS];
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
        if constexpr (HPT != 0 && JC == HPT) {
            // whole layer in registers: park the pre-activations in the (now dead) a1 column and run the
            // activation + layer 3 as a rolled loop (compact code instead of 64 inlined GELUs)
            if (skip != 0.f) {
#pragma unroll
                for (int jj = 0; jj < JC; ++jj) {
                    const float a1 = a1col[jj * PBS];
#pragma unroll
                    for (int m = 0; m < D; ++m) o[m] = fmaf(a1, s.W3[jj * D + m], o[m]);
                }
            }
#pragma unroll
            for (int jj = 0; jj < JC; ++jj) a1col[jj * PBS] = acc[jj];
#pragma unroll 4
            for (int j = 0; j < HP; ++j) {
                const float a2 = act_fwd<ACT>(a1col[j * PBS]);
#pragma unroll
                for (int m = 0; m < D; ++m) o[m] = fmaf(a2, s.W3[j * D + m], o[m]);
            }
        } else {
#pragma unroll
            for (int jj = 0; jj < JC; ++jj) {
                const float a2 = act_fwd<ACT>(acc[jj]);
                const float hs = a2 + skip * a1col[(j0 + jj) * PBS];
#pragma unroll
                for (int m = 0; m < D; ++m) o[m] = fmaf(hs, s.W3[(j0 + jj) * D + m], o[m]);
            }
        }
    }
    const float out_scale = net_out_scale(nv);
#pragma unroll
    for (int m = 0; m < D; ++m) out[m] = out_scale * fminf(fmaxf(o[m], -nv.out_clip), nv.out_clip);
}

}  // namespace cmcd
