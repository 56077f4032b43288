"""Bit-exact numpy restatement of the JAX threefry PRNG calls on the CMCD hot path.

ORACLE / TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Reference call sites (all under /root/reference/src):
  jax.random.PRNGKey / split / normal   mcdboundingmachine.py:151-162, mcd_cais.py:66,87,94,
                                        mcd_utils.py:14-16, vardist/diag_gauss.py:49-62
  jax.random.uniform                    model_handler.py:256-261 (many_gmm means)
The algorithm lives in jax (un-pinned third-party dependency, jax ~0.4.14-0.4.24 era,
``jax_threefry_partitionable=False``); restated from its published definition:
Threefry-2x32, 20 rounds (Salmon et al., Random123).
"""
from __future__ import annotations

import numpy as np

_U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl(x, r):
    return (x << _U32(r)) | (x >> _U32(32 - r))


def threefry2x32(k0, k1, x0, x1):
    """20-round Threefry-2x32 block function, vectorised (all args uint32 arrays/scalars)."""
    with np.errstate(over="ignore"):
        k0 = np.asarray(k0, _U32)
        k1 = np.asarray(k1, _U32)
        x0 = np.asarray(x0, _U32).copy()
        x1 = np.asarray(x1, _U32).copy()
        ks = (k0, k1, k0 ^ k1 ^ _U32(0x1BD11BDA))
        x0 = x0 + ks[0]
        x1 = x1 + ks[1]
        for g in range(5):
            for r in _ROT[g % 2]:
                x0 = x0 + x1
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = x0 + ks[(g + 1) % 3]
            x1 = x1 + ks[(g + 2) % 3] + _U32(g + 1)
        return x0, x1


def prng_key(seed):
    """jax.random.PRNGKey(seed) for 32-bit seeds: key = (0, seed).  Returns [...,2] uint32."""
    seed = np.asarray(seed)
    lo = seed.astype(np.int64).astype(np.uint64) & np.uint64(0xFFFFFFFF)
    return np.stack([np.zeros_like(lo, dtype=_U32), lo.astype(_U32)], axis=-1)


def threefry_counts(key, n):
    """jax's threefry_2x32(key, iota(n)): pad odd n, split counts in halves.  key [...,2] -> [...,n]."""
    key = np.asarray(key, _U32)
    m = (n + 1) // 2
    c = np.arange(2 * m, dtype=_U32)
    c[n:] = 0
    y0, y1 = threefry2x32(key[..., 0:1], key[..., 1:2], c[:m], c[m:])
    return np.concatenate([y0, y1], axis=-1)[..., :n]


def split(key):
    """jax.random.split(key) (num=2).  key [...,2] -> (key_a, key_b) each [...,2]."""
    bits = threefry_counts(key, 4)
    return bits[..., 0:2], bits[..., 2:4]


def random_bits(key, d):
    return threefry_counts(key, d)


# ---------------------------------------------------------------------------------------
# float32 helpers -- every operation below is a single correctly rounded fp32 op (no FMA),
# so the CUDA kernel (which uses __fadd_rn/__fmul_rn/__fdiv_rn/__fsqrt_rn in the same
# order, cmcd_b200/csrc/prng.cuh) reproduces the results bit for bit.
# ---------------------------------------------------------------------------------------
_f = np.float32

_LN2_HI = _f(6.9313812256e-01)
_LN2_LO = _f(9.0580006145e-06)
_LG1 = _f(0.66666662693)
_LG2 = _f(0.40000972152)
_LG3 = _f(0.28498786688)
_LG4 = _f(0.24279078841)
_SQRT_HALF = _f(0.70710678118654752440)


def log_f32(t):
    """Deterministic fp32 natural log for normal positive t (fdlibm-style, basic ops only)."""
    t = np.asarray(t, _f)
    m, e = np.frexp(t)  # m in [0.5,1)
    m = m.astype(_f)
    e = e.astype(np.int32)
    small = m < _SQRT_HALF
    m = np.where(small, m * _f(2.0), m).astype(_f)
    e = np.where(small, e - 1, e)
    f = m - _f(1.0)
    s = f / (_f(2.0) + f)
    z = s * s
    w = z * z
    t1 = w * (_LG2 + w * _LG4)
    t2 = z * (_LG1 + w * _LG3)
    r = t2 + t1
    hfsq = (_f(0.5) * f) * f
    dk = e.astype(_f)
    return dk * _LN2_HI - ((hfsq - (s * (hfsq + r) + dk * _LN2_LO)) - f)


def log1p_f32(u):
    """Deterministic fp32 log1p for u in (-1, 0]: log(t) - ((t-1)-u)/t with t = fl(1+u)."""
    u = np.asarray(u, _f)
    t = _f(1.0) + u
    c = (t - _f(1.0)) - u
    return log_f32(t) - c / t


_ERFINV_LT = [_f(v) for v in (2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06,
                              0.00021858087, -0.00125372503, -0.00417768164, 0.246640727, 1.50140941)]
_ERFINV_GT = [_f(v) for v in (-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844,
                              0.00573950773, -0.0076224613, 0.00943887047, 1.00167406, 2.83297682)]


def erf_inv_f32(x, log1p=log1p_f32):
    """XLA's float32 erf_inv (Giles' single-precision approximation), op order preserved."""
    x = np.asarray(x, _f)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = -log1p(-(x * x))
        lt = w < _f(5.0)
        w = np.where(lt, w - _f(2.5), np.sqrt(np.maximum(w, _f(0))) - _f(3.0)).astype(_f)
        p = np.where(lt, _ERFINV_LT[0], _ERFINV_GT[0]).astype(_f)
        for a, b in zip(_ERFINV_LT[1:], _ERFINV_GT[1:]):
            p = (np.where(lt, a, b).astype(_f) + p * w).astype(_f)
        r = p * x
        return np.where(np.abs(x) == _f(1.0), x * _f(np.inf), r).astype(_f)


def bits_to_unit_float(bits):
    """mantissa trick: bitcast((bits >> 9) | 0x3F800000) - 1.0 in [0,1)."""
    bits = np.asarray(bits, _U32)
    return ((bits >> _U32(9)) | _U32(0x3F800000)).view(_f) - _f(1.0)


def uniform(key, d, minval=0.0, maxval=1.0):
    """jax.random.uniform(key, (d,), float32, minval, maxval)."""
    f = bits_to_unit_float(random_bits(key, d))
    lo, hi = _f(minval), _f(maxval)
    return np.maximum(lo, f * (hi - lo) + lo).astype(_f)


_NORMAL_LO = np.nextafter(_f(-1.0), _f(0.0))
_SQRT2 = _f(np.sqrt(2))


def normal(key, d, log1p=log1p_f32):
    """jax.random.normal(key, (d,)) float32: sqrt(2) * erf_inv(uniform(nextafter(-1,0), 1))."""
    u = uniform(key, d, _NORMAL_LO, 1.0)
    return (_SQRT2 * erf_inv_f32(u, log1p)).astype(_f)


# ---------------------------------------------------------------------------------------
# per-particle key chain of compute_log_elbo + evolve (mcdboundingmachine.py:151-162,
# mcd_cais.py:66,87,94 -- identical in mcd_cais_var.py and mcd_over_orig.py)
# ---------------------------------------------------------------------------------------
def particle_noise(seeds, dim, nbridges):
    """Returns (xi0 [N,d], xi [K,N,d]) -- every Gaussian a particle consumes, in order."""
    k0 = prng_key(np.asarray(seeds))
    a, k = split(k0)                 # mcdboundingmachine.py:153
    xi0 = normal(a, dim)             # :156 -> diag_gauss.py:55
    xi = np.zeros((nbridges,) + xi0.shape, _f)
    if nbridges >= 1:
        a, _ = split(k)              # :162  (rng_key passed to evolve as rng_key_gen)
        _, k = split(a)              # mcd_cais.py:94 (first half discarded)
        for i in range(nbridges):
            a, k = split(k)          # mcd_cais.py:66
            xi[i] = normal(a, dim)   # mcd_utils.py:15
            _, k = split(k)          # mcd_cais.py:87
    return xi0, xi


def evolve_noise(keys, dim, nbridges):
    """Gaussians consumed by mcd_utils.evolve(z, betas, params, rng_key_gen, ...) started from ``keys`` [N,2] = rng_key_gen
    (mcd_cais.py:94 then :66,87 per step; identical in mcd_cais_var.py / mcd_over_orig.py).  Returns xi [K,N,d]."""
    _, k = split(np.asarray(keys, _U32))     # mcd_cais.py:94
    xi = np.zeros((nbridges, k.shape[0], dim), _f)
    for i in range(nbridges):
        a, k = split(k)                      # :66
        xi[i] = normal(a, dim)
        _, k = split(k)                      # :87
    return xi


def particle_noise_ud(seeds, dim, nbridges):
    """Key chain of the underdamped lp_a operator (mcdboundingmachine.py:151-162 + mcd_under_lp_a.py:62-73,31,59):
    returns (xi0 [N,d], rho0 [N,d], xi [K,N,d])."""
    k0 = prng_key(np.asarray(seeds))
    a, k = split(k0)                 # mcdboundingmachine.py:153
    xi0 = normal(a, dim)             # :156
    rho0 = np.zeros_like(xi0)
    xi = np.zeros((nbridges,) + xi0.shape, _f)
    if nbridges >= 1:
        g, _ = split(k)              # :162 (rng_key handed to evolve as rng_key_gen)
        a, g = split(g)              # mcd_under_lp_a.py:62
        rho0 = normal(a, dim)        # :63
        _, g = split(g)              # :70
        for i in range(nbridges):
            a, g = split(g)          # :31
            xi[i] = normal(a, dim)   # :32 -> mcd_utils.py:15
            _, g = split(g)          # :59
    return xi0, rho0, xi
