// tcgen05 / TMEM / mbarrier primitives (inline PTX, sm_100a) used by the tensor-core bridge kernels.
//
// Operand conventions of this library (validated on B200 by tools/umma_probe.cu / tools/umma_probe2.cu):
//   * accumulator D[128 x N] fp32 in TMEM: row m = TMEM lane m, column n = TMEM column d_col + n;
//   * A[128 x K] kind::tf32 in TMEM: row m = lane m, K element k = column a_col + k (one 32-bit word per element);
//   * B[N x K] kind::tf32 in shared memory, K-major, no swizzle, "core-matrix" form: 8 rows x 16 B core matrices,
//     element (n, k) at byte (n/8)*SBO + (k/4)*128 + (n%8)*16 + (k%4)*4 with SBO = 32*K bytes; one MMA consumes K = 8
//     (two core matrices, LBO = 128 B apart).
// fp32 operands are read by the tensor core as tf32 (low 13 mantissa bits ignored); fp32-level accuracy comes from
// the 3-pass split  a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi  with a_hi = RN_tf32(a), a_lo = a - a_hi (exact).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace cmcd {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (row r, k) in a K-major core-matrix tile with `kdim` K elements per row
__host__ __device__ inline int core_off(int r, int k, int kdim) { return (r / 8) * (32 * kdim) + (k / 4) * 128 + (r % 8) * 16 + (k % 4) * 4; }

// shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (Blackwell)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// instruction descriptor: kind::tf32, fp32 accumulate, A and B K-major, dense
__host__ __device__ inline uint32_t make_idesc_tf32(int M, int N) {
    uint32_t d = 0;
    d |= 1u << 4;                  // c_format = F32
    d |= 2u << 7;                  // a_format = TF32
    d |= 2u << 10;                 // b_format = TF32
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}

// D[tmem] (+)= A[smem] * B[smem]^T
__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}

// ---- kind::f16 with bf16 operands, both MN-major, SWIZZLE_128B (validated by tools/umma_probe3.cu) ----------------
// An MN-major operand tile is a plain row-major [k][64] bf16 array (128 B per k row, the eight 16 B chunks of a row
// XOR-ed with k % 8): exactly what a thread that owns row k can write with two STS.128 per 16 elements.  Groups of
// 8 k-rows are SBO = 1024 B apart; further 64-wide groups of the M/N dimension are LBO bytes apart (stacked tiles).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;  // SWIZZLE_128B
    return d;
}
__host__ __device__ inline uint32_t make_idesc_bf16_mn(int M, int N) {
    uint32_t d = 0;
    d |= 1u << 4;                  // c_format = F32
    d |= 1u << 7;                  // a_format = BF16
    d |= 1u << 10;                 // b_format = BF16
    d |= 1u << 15;                 // A MN-major
    d |= 1u << 16;                 // B MN-major
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
// kind::f16, bf16 x bf16 -> fp32, both operands K-major; A may come from TMEM as packed pairs: column c of lane m holds
// (A[m][2c] in the low half, A[m][2c+1] in the high half) -- validated by tools/umma_probe5.cu.  One MMA consumes K = 16
// = 8 packed TMEM columns and two 8x16 B core matrices of the B tile (element (n, k) at (n/8)*16*K + (k/8)*128 + (n%8)*16 + (k%8)*2).
__host__ __device__ inline uint32_t make_idesc_bf16_k(int M, int N) {
    uint32_t d = 0;
    d |= 1u << 4; d |= 1u << 7; d |= 1u << 10;
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__host__ __device__ inline int core_off16(int n, int k, int kdim) { return (n / 8) * (16 * kdim) + (k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2; }
// byte offset of element (k, i) in an MN-major SW128 bf16 tile
__host__ __device__ inline int sw128_off(int k, int i) { return k * 128 + ((((i / 8) ^ (k % 8)) * 16) + (i % 8) * 2); }

// make all previously issued tcgen05.mma of this thread arrive on `mbar` when they complete
__device__ __forceinline__ void commit(uint64_t* mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(mbar)) : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(mbar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;");
}
// plain arrive (release.cta): counts one of the `count` arrivals of the current phase
__device__ __forceinline__ void mbar_arrive(uint64_t* mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
    uint32_t done = 0;
    const uint32_t addr = smem_u32(mbar);
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    }
}

// TMA bulk copy (cp.async.bulk, non-tensor form): `bytes` (multiple of 16, 16-byte aligned on both sides) global -> shared,
// completion signalled on `mbar` as transaction bytes; the issuing thread first arms the barrier with expect_tx.
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(mbar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(mbar)) : "memory");
}

// true in exactly one lane of a fully converged warp (the lane that issues tcgen05.mma / commit for the CTA)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// one warp allocates `ncols` (power of two >= 32) TMEM columns; base address lands in *slot (shared memory)
// (last = false: more allocations follow -- the permit is relinquished only after the CTA's final allocation)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols, bool last = true) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "r"(ncols) : "memory");
    if (last) asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(base), "r"(ncols) : "memory");
}

// warp-collective: thread (lane l of warp w) reads / writes 16 consecutive columns of TMEM lane 32*(w%4)+l.
// taddr = base + (lane_quarter << 21) + column  (lane index lives in bits 31..16)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
// warp-collective 16x256b.x2 load: thread t of the warp receives, from the 16 TMEM lanes starting at the lane in taddr and
// the 16 columns starting at its column: v[r] = element (lane (t % 32) / 4 + 8 ((r / 2) % 2), column 8 (r / 4) + 2 (t % 4) + r % 2)
// (validated by tools/umma_probe4.cu).  Together with tcgen05.st.32x32b this moves data across the lanes of a warp.
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
           "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 16-byte asynchronous global -> shared copy (LDGSTS) and its group bookkeeping
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// fp32 -> (hi, lo) with hi = RN to 11 significant bits (tf32 grid), lo = x - hi exact in fp32
__host__ __device__ inline void split_tf32(float x, float& hi, float& lo) {
#ifdef __CUDA_ARCH__
    const uint32_t u = (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
    hi = __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.f = x; c.u = (c.u + 0x1000u) & 0xFFFFE000u; hi = c.f;
#endif
    lo = x - hi;
}

}  // namespace umma
}  // namespace cmcd
