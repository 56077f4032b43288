// tcgen05 probe 4: cross-lane data exchange through TMEM.  Every thread writes 16 values into its own lane
// (tcgen05.st.32x32b.x16), the warp reads the same block back with tcgen05.ld.16x256b.x2 (twice: lanes 0-15 and
// 16-31 of the warp's quarter).  Prints which (lane, column) each register of each thread received -- the layout that
// a TMEM-based column reduction over the 32 particles of a warp would rely on.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I cmcd_b200/csrc -o tools/build/umma_probe4 tools/umma_probe4.cu
#include <cstdio>
#include <vector>

#include "umma.cuh"

using namespace cmcd::umma;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
}

__global__ void __launch_bounds__(128) probe4_kernel(float* out) {
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(&slot, 32);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t base = slot + ((uint32_t)(warp * 32) << 16);
    uint32_t w[16];
    for (int c = 0; c < 16; ++c) w[c] = __float_as_uint((float)(tid * 100 + c));   // value encodes (TMEM lane, column)
    tmem_st16(base, w);
    tmem_st_wait();
    fence_before();
    __syncwarp();
    fence_after();
    uint32_t a[8], b[8];
    tmem_ld_16x256b_x2(base, a);                                  // lanes 0..15 of this warp's quarter, 16 columns
    tmem_ld_16x256b_x2(base + ((uint32_t)16 << 16), b);           // lanes 16..31
    tmem_ld_wait();
    for (int r = 0; r < 8; ++r) { out[tid * 16 + r] = __uint_as_float(a[r]); out[tid * 16 + 8 + r] = __uint_as_float(b[r]); }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(slot, 32);
}

int main() {
    float* d; CK(cudaMalloc(&d, 128 * 16 * 4));
    probe4_kernel<<<1, 128>>>(d);
    CK(cudaDeviceSynchronize());
    std::vector<float> h(128 * 16);
    CK(cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost));
    for (int t : {0, 1, 2, 3, 4, 5, 31, 32, 37, 127}) {
        printf("thread %3d:", t);
        for (int r = 0; r < 16; ++r) { int v = (int)h[t * 16 + r]; printf(" (%d,%d)", v / 100, v % 100); }
        printf("\n");
    }
    // check the hypothesis: reg r of load L: lane = 32*warp + 16*L + (t%32)/4 + 8*((r/2)%2), col = 8*(r/4) + 2*(t%4) + r%2
    int bad = 0;
    for (int t = 0; t < 128; ++t)
        for (int L = 0; L < 2; ++L)
            for (int r = 0; r < 8; ++r) {
                const int v = (int)h[t * 16 + L * 8 + r];
                const int lane = 32 * (t / 32) + 16 * L + (t % 32) / 4 + 8 * ((r / 2) % 2), col = 8 * (r / 4) + 2 * (t % 4) + r % 2;
                if (v != lane * 100 + col) ++bad;
            }
    printf("hypothesis mismatches: %d of %d\n", bad, 128 * 16);
    return 0;
}
