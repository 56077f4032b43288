// Host-training-step kernels (SURVEY.md section 8f "next" rows): the optimizer update and the per-iteration seed
// draw of the reference's training loop, moved onto the device so that iterations can pipeline without a host sync.
//
// Replaces, in src/opt.py: optimizer.update + optax.apply_updates + project (:126-128, :14-24, optimizer =
// optax.chain(optax.clip(5.0), optax.adam(lr)) :26-35), optax.incremental_update for the EMA copy (:129-132), and
// seeds = jax.random.randint(rng_key, (N,), 1, 1e6) (:93-94, :182-184).
#include "common.cuh"

namespace cmcd {

// p <- clamp(p - lr * mhat / (sqrt(vhat) + eps), lo, hi) with Adam moments on the clipped gradient
__global__ void __launch_bounds__(256) adam_project_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                           float* __restrict__ v, const float* __restrict__ lo,
                                                           const float* __restrict__ hi, long long n, float lr, float b1, float b2,
                                                           float eps, float clip, float bc1, float bc2, float* __restrict__ ema,
                                                           float ema_step, const int32_t* __restrict__ skip_flag) {
    if (skip_flag && *skip_flag) return;   // diverged (opt.py:122-124): parameters stay as they are
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gi = fminf(fmaxf(g[i], -clip), clip);          // optax.clip(5.0): elementwise
    const float mi = b1 * m[i] + (1.0f - b1) * gi;             // optax.scale_by_adam
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float mhat = mi / bc1, vhat = vi / bc2;              // bias correction, bc = 1 - b^count
    float x = p[i] - lr * (mhat / (sqrtf(vhat) + eps));
    if (lo) x = fmaxf(x, lo[i]);                               // project: jnp.clip / relu(x - 0.001) + 0.001 == max(x, 0.001)
    if (hi) x = fminf(x, hi[i]);
    p[i] = x;
    if (ema) ema[i] = ema_step * x + (1.0f - ema_step) * ema[i];
}

// jax.random.randint(key, (n,), minval, maxval), int32 (jax/_src/random.py::_randint): two 32-bit draws per element
// from the two halves of split(key), combined as ((hi % span) * mult + lo % span) % span in uint32 arithmetic, mult = ((2^16 % span)^2 mod 2^32) % span.
__global__ void __launch_bounds__(256) randint_kernel(uint32_t key0, uint32_t key1, long long n, int32_t minval, uint32_t span,
                                                      uint32_t mult, int32_t* __restrict__ out) {
    __shared__ Key k12[2];
    if (threadIdx.x == 0) {
        Key k; k.k0 = key0; k.k1 = key1;
        split(k, k12[0], k12[1]);
    }
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // random_bits(key, (n,)): threefry over iota(n) split in halves, element i comes from block (i mod h)
    const long long h = (n + 1) / 2;
    const long long blk = i < h ? i : i - h;
    uint32_t bits[2];
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        uint32_t x0 = (uint32_t)blk, x1 = (blk + h < n) ? (uint32_t)(blk + h) : 0u;
        threefry2x32(k12[w], x0, x1);
        bits[w] = i < h ? x0 : x1;
    }
    const uint32_t off = ((bits[0] % span) * mult + (bits[1] % span)) % span;
    out[i] = minval + (int32_t)off;
}

int launch_adam_project(cudaStream_t st, float* p, const float* g, float* m, float* v, const float* lo, const float* hi, long long n,
                        float lr, float b1, float b2, float eps, float clip, int step, float* ema, float ema_step,
                        const int32_t* skip_flag) {
    const float bc1 = 1.0f - powf(b1, (float)step), bc2 = 1.0f - powf(b2, (float)step);
    adam_project_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, g, m, v, lo, hi, n, lr, b1, b2, eps, clip, bc1, bc2, ema,
                                                                     ema_step, skip_flag);
    CMCD_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_randint(cudaStream_t st, uint32_t key0, uint32_t key1, long long n, int32_t minval, int32_t maxval, int32_t* out) {
    uint32_t span = maxval > minval ? (uint32_t)((long long)maxval - (long long)minval) : 1u;
    uint32_t mult = 65536u % span;
    // lax.rem(lax.mul(multiplier, multiplier), span) is evaluated in uint32 by JAX: for span > 65536 the square of
    // 2^16 wraps to 0, so the multiplier is 0 and the offset reduces to lower_bits % span (e.g. span = 999999 for
    // randint(key, (N,), 1, 1e6), opt.py:94).  The wrap is part of the reference's stream and is reproduced here.
    mult = (uint32_t)(mult * mult) % span;
    randint_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(key0, key1, n, minval, span, mult, out);
    CMCD_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace cmcd
