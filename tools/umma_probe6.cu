// tcgen05 probe 6: kind::f16 (bf16, fp32 accumulate) with the A operand K-MAJOR, SWIZZLE_128B, read from SHARED memory --
// the layout the adjoint kernel's per-particle staging tiles already have (row p = particle, 64 bf16 = 128 B per row, 16 B chunks
// XOR-ed with p % 8): D[p][n] = sum_k A[p][k] B(k, n), M = 128 particles, K = 64, N = 64, 4 MMAs of K = 16 (descriptor start
// advanced by 32 B per K block inside the swizzle atom).
//   test 0: B MN-major SW128: tile stored [k][n] row-major (k rows of 128 B)      -> GEMM1 with B = W2[i][j] as stored
//   test 1: B K-major  SW128: tile stored [n][k] row-major (n rows of 128 B)      -> GEMM2 with B = W2[i][j] as stored
// Each test is run for several LBO encodings of the A descriptor (the field is documented as unused for swizzled K-major tiles).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I cmcd_b200/csrc -o tools/build/umma_probe6 tools/umma_probe6.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_bf16.h>

#include "umma.cuh"

using namespace cmcd::umma;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
    uint32_t d = 0;
    d |= 1u << 4; d |= 1u << 7; d |= 1u << 10;
    d |= (uint32_t)a_mn << 15;
    d |= (uint32_t)b_mn << 16;
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}
__host__ __device__ inline int sw_off(int r, int c) { return r * 128 + ((((c / 8) ^ (r % 8)) * 16) + (c % 8) * 2); }

// A [128 p][64 k] fp32, Bm [64 k][64 n] fp32 (global) -> bf16 tiles
__global__ void __launch_bounds__(128) probe6_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ Dout, int test,
                                                      int lbo) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;              // 16 KB
    uint8_t* sB = smem + 16384;      // 8 KB
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t mbar;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int idx = tid; idx < 128 * 64; idx += 128) {
        const int p = idx / 64, k = idx % 64;
        *(__nv_bfloat16*)(sA + sw_off(p, k)) = __float2bfloat16(A[idx]);
    }
    for (int idx = tid; idx < 64 * 64; idx += 128) {
        const int k = idx / 64, n = idx % 64;
        const __nv_bfloat16 v = __float2bfloat16(Bm[idx]);
        if (test == 0) *(__nv_bfloat16*)(sB + sw_off(k, n)) = v;     // [k][n]
        else *(__nv_bfloat16*)(sB + sw_off(n, k)) = v;               // [n][k]
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, 64);
    if (tid == 0) mbar_init(&mbar, 1);
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = make_idesc(128, 64, 0, test == 0 ? 1 : 0);
        for (int k = 0; k < 4; ++k) {
            const uint64_t ad = make_desc_sw128(smem_u32(sA) + k * 32, lbo, 1024);
            const uint64_t bd = (test == 0) ? make_desc_sw128(smem_u32(sB) + k * 2048, 8192, 1024)     // 16 k-rows further
                                            : make_desc_sw128(smem_u32(sB) + k * 32, lbo, 1024);        // 32 B further inside the rows
            mma_f16_ss(tmem, ad, bd, idesc, k > 0);
        }
        commit(&mbar);
    }
    mbar_wait(&mbar, 0);
    fence_after();
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 4; ++c) {
        uint32_t v[16];
        tmem_ld16(lane_base + c * 16, v);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) Dout[tid * 64 + c * 16 + j] = __uint_as_float(v[j]);
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}

static float bf16r(float x) {
    uint32_t u; memcpy(&u, &x, 4);
    u += 0x7FFFu + ((u >> 16) & 1u); u &= 0xFFFF0000u;
    memcpy(&x, &u, 4); return x;
}

int main() {
    std::vector<float> A(128 * 64), B(64 * 64), D(128 * 64);
    srand(3);
    for (auto& x : A) x = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    for (auto& x : B) x = (rand() / (float)RAND_MAX - 0.5f);
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(probe6_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    std::vector<double> E(128 * 64);
    for (int p = 0; p < 128; ++p)
        for (int n = 0; n < 64; ++n) {
            double r = 0;
            for (int k = 0; k < 64; ++k) r += (double)bf16r(A[p * 64 + k]) * bf16r(B[k * 64 + n]);
            E[p * 64 + n] = r;
        }
    const int lbos[4] = {16, 0, 1024, 16384};
    for (int test = 0; test < 2; ++test)
        for (int li = 0; li < 4; ++li) {
            CK(cudaMemset(dD, 0xFF, D.size() * 4));
            probe6_kernel<<<1, 128, 24576 + 1024, 0>>>(dA, dB, dD, test, lbos[li]);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
            double maxerr = 0;
            for (int i = 0; i < 128 * 64; ++i) maxerr = fmax(maxerr, fabs(D[i] - E[i]));
            printf("test %d lbo %5d: max |D - E| = %.3e  %s   D[0][0..2]=%.4f %.4f %.4f E=%.4f %.4f %.4f\n", test, lbos[li], maxerr,
                   maxerr < 1e-4 ? "OK" : "MISMATCH", D[0], D[1], D[2], E[0], E[1], E[2]);
        }
    return 0;
}
