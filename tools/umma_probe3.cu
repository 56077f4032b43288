// tcgen05 probe 3: kind::f16 (bf16 operands, fp32 accumulate) layouts for the weight-gradient tiles of the adjoint
// kernel: G[M x N] = sum_p L[p][m] * R[p][n] with K = particles.
//   test 0: A and B MN-major, SWIZZLE_128B: both operands are plain row-major [p][64] bf16 tiles (128 B per
//           particle row, 16 B chunks XOR-ed with p%8) -- the layout a particle-owning thread can write with STS.128.
//           M = 128 = two stacked 64-wide tiles (LBO = tile stride), N = 64, K = 128 (8 MMAs of K = 16).
//   test 1: A and B K-major, no swizzle, core-matrix layout (8 rows x 16 B): operands stored transposed [m][p].
//   test 2: as test 0 but M = 64 (single tile) -- which TMEM lanes hold the rows.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I cmcd_b200/csrc -o tools/build/umma_probe3 tools/umma_probe3.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_bf16.h>

#include "umma.cuh"

using namespace cmcd::umma;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;  // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ uint32_t make_idesc_bf16(int M, int N, int a_mn, int b_mn) {
    uint32_t d = 0;
    d |= 1u << 4;    // c_format = F32
    d |= 1u << 7;    // a_format = BF16
    d |= 1u << 10;   // b_format = BF16
    d |= (uint32_t)a_mn << 15;
    d |= (uint32_t)b_mn << 16;
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}

// row-major [p][64] bf16 tile with the 128B swizzle: byte offset of element (p, i)
__host__ __device__ inline int sw_off(int p, int i) { return p * 128 + ((((i / 8) ^ (p % 8)) * 16) + (i % 8) * 2); }
// K-major core-matrix tile [m][p] bf16, kdim = 128: 8 rows x 16 B (8 bf16)
__host__ __device__ inline int core_off16(int m, int p) { return (m / 8) * (16 * 128) + (p / 8) * 128 + (m % 8) * 16 + (p % 8) * 2; }

// L: [128 p][128 m] (two 64-wide tiles), R: [128 p][64 n], both fp32 in global, rounded to bf16 on load
__global__ void __launch_bounds__(128) probe3_kernel(const float* __restrict__ L, const float* __restrict__ R, float* __restrict__ Dout, int test) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;              // 2 x 16 KB
    uint8_t* sB = smem + 32768;      // 16 KB
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t mbar;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int idx = tid; idx < 128 * 128; idx += 128) {
        const int p = idx / 128, m = idx % 128;
        const __nv_bfloat16 v = __float2bfloat16(L[idx]);
        if (test == 1) *(__nv_bfloat16*)(sA + core_off16(m, p)) = v;
        else *(__nv_bfloat16*)(sA + (m / 64) * 16384 + sw_off(p, m % 64)) = v;
    }
    for (int idx = tid; idx < 128 * 64; idx += 128) {
        const int p = idx / 64, n = idx % 64;
        const __nv_bfloat16 v = __float2bfloat16(R[idx]);
        if (test == 1) *(__nv_bfloat16*)(sB + core_off16(n, p)) = v;
        else *(__nv_bfloat16*)(sB + sw_off(p, n)) = v;
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, 64);
    if (tid == 0) mbar_init(&mbar, 1);
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        const int M = (test == 2) ? 64 : 128;
        if (test == 1) {
            const uint32_t idesc = make_idesc_bf16(M, 64, 0, 0);
            for (int k = 0; k < 8; ++k) {  // K = 16 particles = 2 core matrices = 256 B
                const uint64_t ad = make_desc(smem_u32(sA) + k * 256, 128, 16 * 128);
                const uint64_t bd = make_desc(smem_u32(sB) + k * 256, 128, 16 * 128);
                mma_f16_ss(tmem, ad, bd, idesc, k > 0);
            }
        } else {
            const uint32_t idesc = make_idesc_bf16(M, 64, 1, 1);
            for (int k = 0; k < 8; ++k) {  // K = 16 particles = 2 groups of 8 rows = 2048 B
                const uint64_t ad = make_desc_sw128(smem_u32(sA) + k * 2048, 16384, 1024);
                const uint64_t bd = make_desc_sw128(smem_u32(sB) + k * 2048, 16384, 1024);
                mma_f16_ss(tmem, ad, bd, idesc, k > 0);
            }
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

static float bf16r(float x) {  // round to nearest even bf16
    uint32_t u; memcpy(&u, &x, 4);
    u += 0x7FFFu + ((u >> 16) & 1u); u &= 0xFFFF0000u;
    memcpy(&x, &u, 4); return x;
}

int main() {
    std::vector<float> L(128 * 128), R(128 * 64), D(128 * 64);
    srand(2);
    for (auto& x : L) x = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    for (auto& x : R) x = (rand() / (float)RAND_MAX - 0.5f);
    float *dL, *dR, *dD;
    CK(cudaMalloc(&dL, L.size() * 4)); CK(cudaMalloc(&dR, R.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dL, L.data(), L.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dR, R.data(), R.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(probe3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152 + 1024));
    std::vector<double> E(128 * 64);
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
            double r = 0;
            for (int p = 0; p < 128; ++p) r += (double)bf16r(L[p * 128 + m]) * bf16r(R[p * 64 + n]);
            E[m * 64 + n] = r;
        }
    for (int test = 0; test < 3; ++test) {
        CK(cudaMemset(dD, 0xFF, D.size() * 4));
        probe3_kernel<<<1, 128, 49152, 0>>>(dL, dR, dD, test);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        int mapped = 0; double maxerr = 0;
        printf("test %d: lane->row:", test);
        for (int lane = 0; lane < 128; ++lane) {
            int best = -1; double be = 1e30;
            for (int i = 0; i < 128; ++i) { double e = 0; for (int j = 0; j < 64; ++j) e = fmax(e, fabs(D[lane * 64 + j] - E[i * 64 + j])); if (e < be) { be = e; best = i; } }
            if (be < 1e-3) { mapped++; maxerr = fmax(maxerr, be); if (lane % 16 == 0) printf(" %d->%d", lane, best); }
        }
        printf("  | %d lanes valid, max err %.2e, D[0][0..2]=%.4f %.4f %.4f E[0][0..2]=%.4f %.4f %.4f\n", mapped, maxerr,
               D[0], D[1], D[2], E[0], E[1], E[2]);
    }
    return 0;
}
