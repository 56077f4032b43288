// FP32 pipe microbenchmarks for the roofline denominators ("measure an FFMA
// microbenchmark and quote 'of measured'").  Standalone: nvcc -gencode arch=compute_100a,code=sm_100a.
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;

__global__ void __launch_bounds__(256) k_ffma(float* out, float a, float b) {
    float c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-9f + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fmaf(c[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 3 distinct register operands per FFMA (acc += h * w), like the MLP inner loop
__global__ void __launch_bounds__(256) k_ffma3(float* out, const float* in) {
    float c[16], w[4];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-9f + i;
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = in[i];
    float h = in[5];
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fmaf(h, w[i & 3], c[i]);
        h += 1e-9f;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

__global__ void __launch_bounds__(256) k_ffma2(float* out, const float* in) {
    unsigned long long c[16], w[4];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = (unsigned long long)(threadIdx.x + i);
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = ((const unsigned long long*)in)[i];
    unsigned long long h = ((const unsigned long long*)in)[5];
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = ffma2(h, w[i & 3], c[i]);
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s ^= c[i];
    ((unsigned long long*)out)[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// inner loop of net_fwd: 1 private LDS + 16 broadcast LDS.128 + 64 FFMA per i
__global__ void __launch_bounds__(128) k_mlp_loop(float* out, const float* in, int reps) {
    __shared__ float4 sW[64 * 16];
    __shared__ float sA[64 * 128];
    for (int i = threadIdx.x; i < 64 * 16; i += 128) sW[i] = make_float4(in[i & 7], 1e-3f, 2e-3f, 3e-3f);
    for (int i = threadIdx.x; i < 64 * 128; i += 128) sA[i] = in[i & 7];
    __syncthreads();
    float acc[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) acc[j] = 0.f;
    for (int r = 0; r < reps; ++r) {
#pragma unroll 2
        for (int i = 0; i < 64; ++i) {
            const float h = sA[i * 128 + threadIdx.x];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const float4 w = sW[i * 16 + q];
                acc[4 * q + 0] = fmaf(h, w.x, acc[4 * q + 0]);
                acc[4 * q + 1] = fmaf(h, w.y, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(h, w.z, acc[4 * q + 2]);
                acc[4 * q + 3] = fmaf(h, w.w, acc[4 * q + 3]);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < 64; ++j) s += acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// same loop with packed f32x2 FMAs over output pairs
__global__ void __launch_bounds__(128) k_mlp_loop2(float* out, const float* in, int reps) {
    __shared__ float4 sW[64 * 16];
    __shared__ float sA[64 * 128];
    for (int i = threadIdx.x; i < 64 * 16; i += 128) sW[i] = make_float4(in[i & 7], 1e-3f, 2e-3f, 3e-3f);
    for (int i = threadIdx.x; i < 64 * 128; i += 128) sA[i] = in[i & 7];
    __syncthreads();
    unsigned long long acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = 0ull;
    for (int r = 0; r < reps; ++r) {
#pragma unroll 2
        for (int i = 0; i < 64; ++i) {
            const float h = sA[i * 128 + threadIdx.x];
            unsigned long long hh;
            asm("mov.b64 %0, {%1, %1};" : "=l"(hh) : "f"(h));
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const ulonglong2 w = *reinterpret_cast<const ulonglong2*>(&sW[i * 16 + q]);
                acc[2 * q + 0] = ffma2(hh, w.x, acc[2 * q + 0]);
                acc[2 * q + 1] = ffma2(hh, w.y, acc[2 * q + 1]);
            }
        }
    }
    unsigned long long s = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) s ^= acc[j];
    ((unsigned long long*)out)[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_erf(float* out, float a) {
    float c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = threadIdx.x * 1e-3f + i * 0.1f - 0.5f;
    for (int it = 0; it < 512; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) c[i] = erff(c[i]) * a;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_expf(float* out, float a) {
    float c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = threadIdx.x * 1e-3f + i * 0.1f - 0.5f;
    for (int it = 0; it < 512; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) c[i] = expf(c[i]) * a - 1.0f;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static float time_ms(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f(); f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    float *out, *in;
    CK(cudaMalloc(&out, 148 * 64 * 1024 * 8));
    CK(cudaMalloc(&in, 4096));
    float h[1024];
    for (int i = 0; i < 1024; ++i) h[i] = 1e-3f * (i + 1);
    CK(cudaMemcpy(in, h, 4096, cudaMemcpyHostToDevice));
    const int grid = sms * 8, thr = 256;
    const double flop = 2.0 * ITERS * 16.0 * grid * thr;
    float ms;
    ms = time_ms([&] { k_ffma<<<grid, thr>>>(out, 1.0001f, 1e-7f); });
    printf("{\"bench\":\"ffma_imm2reg\",\"tflops\":%.2f,\"ms\":%.4f}\n", flop / ms * 1e-9, ms);
    ms = time_ms([&] { k_ffma3<<<grid, thr>>>(out, in); });
    printf("{\"bench\":\"ffma_3reg\",\"tflops\":%.2f,\"ms\":%.4f}\n", flop / ms * 1e-9, ms);
    ms = time_ms([&] { k_ffma2<<<grid, thr>>>(out, in); });
    printf("{\"bench\":\"ffma2_f32x2\",\"tflops\":%.2f,\"ms\":%.4f}\n", 2 * flop / ms * 1e-9, ms);
    for (int occ = 1; occ <= 4; ++occ) {
        const int g2 = sms * occ, reps = 64;
        const double f2 = 2.0 * 64 * 64 * reps * (double)g2 * 128;
        ms = time_ms([&] { k_mlp_loop<<<g2, 128>>>(out, in, reps); });
        printf("{\"bench\":\"mlp_loop_ffma\",\"blocks_per_sm\":%d,\"tflops\":%.2f,\"ms\":%.4f}\n", occ, f2 / ms * 1e-9, ms);
        ms = time_ms([&] { k_mlp_loop2<<<g2, 128>>>(out, in, reps); });
        printf("{\"bench\":\"mlp_loop_ffma2\",\"blocks_per_sm\":%d,\"tflops\":%.2f,\"ms\":%.4f}\n", occ, f2 / ms * 1e-9, ms);
    }
    const double nerf = 512.0 * 8 * grid * thr;
    ms = time_ms([&] { k_erf<<<grid, thr>>>(out, 0.999f); });
    printf("{\"bench\":\"erff\",\"gops\":%.1f,\"ms\":%.4f}\n", nerf / ms * 1e-6, ms);
    ms = time_ms([&] { k_expf<<<grid, thr>>>(out, 0.999f); });
    printf("{\"bench\":\"expf\",\"gops\":%.1f,\"ms\":%.4f}\n", nerf / ms * 1e-6, ms);
    return 0;
}
