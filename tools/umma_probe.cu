// tcgen05 probe: validates the shared-memory operand layout / descriptors used by the bridge kernels' tensor-core
// tiles against a host GEMM, and dumps which TMEM lanes hold which accumulator rows for M=128 and M=64.
//   case 0: D[128x64] = A[128x64(K)] * B[64(N)x64(K)]^T, both K-major, no swizzle, kind::tf32
//   case 1: D[64x64]  = A^T * B with A=[128(K)x64(M)], B=[128(K)x64(N)] stored as the same row tiles (MN-major)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/umma_probe tools/umma_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// tile element (row r, col c) of a [rows x 64] fp32 tile in "core-matrix" form:
//   offset = (r/8)*2048 + (c/4)*128 + (r%8)*16 + (c%4)*4   bytes     (8 rows x 16 B core matrices, K chunks adjacent)
__host__ __device__ inline int tile_off(int r, int c) { return (r / 8) * 2048 + (c / 4) * 128 + (r % 8) * 16 + (c % 4) * 4; }

// 128B-swizzled tile: 8-row x 128 B atoms (16 B chunks XOR-ed with row%8); column halves 1024 B apart, row groups 2048 B
__host__ __device__ inline int tile_off_sw(int r, int c) {
    return (r / 8) * 2048 + (c / 32) * 1024 + (r % 8) * 128 + ((((c % 32) / 4) ^ (r % 8)) * 16) + (c % 4) * 4;
}
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;  // SWIZZLE_128B
    return d;
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
    return d;                // layout_type = 0 (no swizzle), base_offset = 0
}

__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
    uint32_t d = 0;
    d |= 1u << 4;                 // c_format = F32
    d |= 2u << 7;                 // a_format = TF32
    d |= 2u << 10;                // b_format = TF32
    d |= (uint32_t)a_mn_major << 15;
    d |= (uint32_t)b_mn_major << 16;
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}

__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ Dout, int mode) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;               // 128 rows x 256 B = 32 KB
    uint8_t* sB = smem + 32768;       // up to 128 rows x 256 B
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t mbar;
    const int tid = threadIdx.x, warp = tid >> 5;
    // stage: A is [128][64] row-major in global; B is [rowsB][64]
    const int rowsB = (mode == 0 || mode == 3 || mode == 5) ? 64 : 128;
    const bool sw = mode >= 5;
    for (int i = tid; i < 128 * 64; i += 128) { int r = i / 64, c = i % 64; *(float*)(sA + (sw ? tile_off_sw(r, c) : tile_off(r, c))) = A[i]; }
    for (int i = tid; i < rowsB * 64; i += 128) { int r = i / 64, c = i % 64; *(float*)(sB + (sw ? tile_off_sw(r, c) : tile_off(r, c))) = B[i]; }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&mbar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        if (mode == 0 || mode == 3) {
            const uint32_t idesc = make_idesc_tf32(mode == 0 ? 128 : 64, 64, 0, 0);
            for (int k = 0; k < 8; ++k) {  // K = 64 in steps of 8 tf32 = 2 chunks of 16 B = 256 B in this layout
                const uint64_t ad = make_desc(smem_u32(sA) + k * 256, 128, 2048);
                const uint64_t bd = make_desc(smem_u32(sB) + k * 256, 128, 2048);
                umma_tf32(tmem, ad, bd, idesc, k > 0);
            }
        } else if (mode == 5 || mode == 6) {
            // 128B swizzle: A K-major; B K-major (5) or MN-major (6, B(n,k) = Bt[k][n])
            const uint32_t idesc = make_idesc_tf32(128, 64, 0, mode == 6 ? 1 : 0);
            for (int k = 0; k < 8; ++k) {
                const uint32_t koff = (k / 4) * 1024 + (k % 4) * 32;
                const uint64_t ad = make_desc_sw128(smem_u32(sA) + koff, 16, 2048);
                const uint64_t bd = (mode == 5) ? make_desc_sw128(smem_u32(sB) + koff, 16, 2048)
                                                : make_desc_sw128(smem_u32(sB) + k * 2048, 1024, 2048);
                umma_tf32(tmem, ad, bd, idesc, k > 0);
            }
        } else if (mode == 7) {
            // 128B swizzle, both MN-major: D[64 x 64] = A^T B over K = 128 rows
            const uint32_t idesc = make_idesc_tf32(64, 64, 1, 1);
            for (int k = 0; k < 16; ++k) {
                const uint64_t ad = make_desc_sw128(smem_u32(sA) + k * 2048, 1024, 2048);
                const uint64_t bd = make_desc_sw128(smem_u32(sB) + k * 2048, 1024, 2048);
                umma_tf32(tmem, ad, bd, idesc, k > 0);
            }
        } else if (mode == 2 || mode == 4) {
            // A K-major (M=128,K=64); B MN-major: B(n,k) = Bt[k][n], Bt = rows 0..63 of the B tile
            const uint32_t idesc = make_idesc_tf32(128, 64, 0, 1);
            for (int k = 0; k < 8; ++k) {
                const uint64_t ad = make_desc(smem_u32(sA) + k * 256, 128, 2048);
                const uint64_t bd = (mode == 2) ? make_desc(smem_u32(sB) + k * 2048, 2048, 128)
                                                : make_desc(smem_u32(sB) + k * 2048, 128, 2048);
                umma_tf32(tmem, ad, bd, idesc, k > 0);
            }
        } else {
            const uint32_t idesc = make_idesc_tf32(64, 64, 1, 1);
            for (int k = 0; k < 16; ++k) {  // K = 128 particles in steps of 8 rows = 2048 B
                const uint64_t ad = make_desc(smem_u32(sA) + k * 2048, 2048, 128);
                const uint64_t bd = make_desc(smem_u32(sB) + k * 2048, 2048, 128);
                umma_tf32(tmem, ad, bd, idesc, k > 0);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
    }
    // wait for the MMA to complete
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // every warp dumps its 32 lanes x 64 columns
    uint32_t v[64];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]),
          "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]),
          "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]),
          "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]),
          "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 64; ++j) Dout[tid * 64 + j] = __uint_as_float(v[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(64));
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

int main() {
    std::vector<float> A(128 * 64), B(128 * 64), D(128 * 64);
    srand(1);
    for (auto& x : A) x = (rand() / (float)RAND_MAX - 0.5f);
    for (auto& x : B) x = (rand() / (float)RAND_MAX - 0.5f);
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024));
    for (int mode = 0; mode < 8; ++mode) {
        CK(cudaMemset(dD, 0xFF, D.size() * 4));
        probe_kernel<<<1, 128, 65536, 0>>>(dA, dB, dD, mode);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        const int M = (mode == 1 || mode == 3 || mode == 7) ? 64 : 128;
        std::vector<double> E((size_t)M * 64);
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < 64; ++n) {
                double r = 0;
                if (mode == 0 || mode == 3 || mode == 5) for (int k = 0; k < 64; ++k) r += (double)tf32_trunc(A[m * 64 + k]) * tf32_trunc(B[n * 64 + k]);
                else if (mode == 2 || mode == 4 || mode == 6) for (int k = 0; k < 64; ++k) r += (double)tf32_trunc(A[m * 64 + k]) * tf32_trunc(B[k * 64 + n]);
                else for (int p = 0; p < 128; ++p) r += (double)tf32_trunc(A[p * 64 + m]) * tf32_trunc(B[p * 64 + n]);
                E[(size_t)m * 64 + n] = r;
            }
        int mapped = 0; double maxerr = 0;
        printf("mode %d (M=%d): lane->row:", mode, M);
        for (int lane = 0; lane < 128; ++lane) {
            int best = -1; double be = 1e30;
            for (int i = 0; i < M; ++i) { double e = 0; for (int j = 0; j < 64; ++j) e = fmax(e, fabs(D[lane * 64 + j] - E[(size_t)i * 64 + j])); if (e < be) { be = e; best = i; } }
            if (be < 1e-4) { mapped++; maxerr = fmax(maxerr, be); if (lane % 16 == 0) printf(" %d->%d", lane, best); }
        }
        printf("  | %d lanes valid, max err %.2e, D[0][0..2]=%.4f %.4f %.4f E[0][0..2]=%.4f %.4f %.4f\n", mapped, maxerr,
               D[0], D[1], D[2], E[0], E[1], E[2]);
    }
    return 0;
}
