// tcgen05 probe 5: kind::f16 with the A operand in TMEM as packed bf16 pairs (two K elements per 32-bit column) and a
// bf16 K-major core-matrix B tile in shared memory: D[128x64] = A[128x64] B[64x64]^T over K = 64 (4 MMAs of K = 16).
//   order 0: column c of lane m holds (A[m][2c] in the low half, A[m][2c+1] in the high half); order 1: swapped.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I cmcd_b200/csrc -o tools/build/umma_probe5 tools/umma_probe5.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_bf16.h>
#include "umma.cuh"

using namespace cmcd::umma;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t make_idesc_bf16_k(int M, int N) {
    uint32_t d = 0;
    d |= 1u << 4; d |= 1u << 7; d |= 1u << 10;   // f32 accumulate, bf16 x bf16, both K-major
    d |= (uint32_t)(N >> 3) << 17; d |= (uint32_t)(M >> 4) << 24;
    return d;
}
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__host__ __device__ inline int core_off16(int n, int k, int kdim) { return (n / 8) * (16 * kdim) + (k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2; }

__global__ void __launch_bounds__(128) probe5_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ Dout, int order) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t slot;
    __shared__ __align__(8) uint64_t mbar;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 64 * 64; i += 128) *(__nv_bfloat16*)(smem + core_off16(i / 64, i % 64, 64)) = __float2bfloat16(B[i]);
    if (warp == 0) tmem_alloc(&slot, 128);
    if (tid == 0) mbar_init(&mbar, 1);
    fence_async_smem(); fence_before(); __syncthreads(); fence_after();
    const uint32_t tmem = slot, lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 2; ++c) {   // 32 packed columns = 64 K elements
        uint32_t w[16];
        for (int j = 0; j < 16; ++j) {
            const int k = (c * 16 + j) * 2;
            const __nv_bfloat162 v = order == 0 ? __floats2bfloat162_rn(A[tid * 64 + k], A[tid * 64 + k + 1])
                                                : __floats2bfloat162_rn(A[tid * 64 + k + 1], A[tid * 64 + k]);
            w[j] = *reinterpret_cast<const uint32_t*>(&v);
        }
        tmem_st16(lane_base + 64 + c * 16, w);
    }
    tmem_st_wait(); fence_before(); __syncthreads();
    if (tid == 0) {
        fence_after();
        const uint32_t idesc = make_idesc_bf16_k(128, 64);
        for (int k = 0; k < 4; ++k)   // K = 16 per MMA: 8 packed TMEM columns, 2 core matrices (256 B) of B
            mma_f16_ts(tmem, tmem + 64 + k * 8, make_desc(smem_u32(smem) + k * 256, 128, 16 * 64), idesc, k > 0);
        commit(&mbar);
    }
    mbar_wait(&mbar, 0); fence_after();
    for (int c = 0; c < 4; ++c) {
        uint32_t v[16];
        tmem_ld16(lane_base + c * 16, v); tmem_ld_wait();
        for (int j = 0; j < 16; ++j) Dout[tid * 64 + c * 16 + j] = __uint_as_float(v[j]);
    }
    fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

static float bf16r(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x7FFFu + ((u >> 16) & 1u); u &= 0xFFFF0000u; memcpy(&x, &u, 4); return x; }

int main() {
    std::vector<float> A(128 * 64), B(64 * 64), D(128 * 64);
    srand(3);
    for (auto& x : A) x = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    for (auto& x : B) x = (rand() / (float)RAND_MAX - 0.5f);
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    for (int order = 0; order < 2; ++order) {
        probe5_kernel<<<1, 128, 8192 + 1024>>>(dA, dB, dD, order);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 64; ++n) {
                double r = 0;
                for (int k = 0; k < 64; ++k) r += (double)bf16r(A[m * 64 + k]) * bf16r(B[n * 64 + k]);
                maxerr = fmax(maxerr, fabs(D[m * 64 + n] - r));
            }
        printf("order %d: max abs err %.3e  D[0][0..2]=%.4f %.4f %.4f\n", order, maxerr, D[0], D[1], D[2]);
    }
    return 0;
}
