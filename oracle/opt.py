"""Training-loop pieces restated on the CPU -- ORACLE / TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows /root/reference/src/opt.py: project :14-24, optimizer = optax.chain(optax.clip(5.0), optax.adam) :26-35
(optax 0.1.3 ``scale_by_adam``: bias-corrected moments, eps outside the square root, eps_root = 0), the update /
apply / project sequence :126-128, and seeds = jax.random.randint(key, (N,), 1, 1e6) :93-94.  ``randint`` restates
jax/_src/random.py::_randint (two 32-bit draws from split(key), combined modulo the span in uint32 arithmetic); like
everything downstream of threefry it is **unpinned** against a running JAX."""
from __future__ import annotations

import numpy as np

from . import prng

_U32 = np.uint32


def randint(key, n, minval, maxval):
    """jax.random.randint(key, (n,), minval, maxval) -> int32[n]."""
    k1, k2 = prng.split(np.asarray(key, _U32))
    hi = prng.random_bits(k1, n).astype(np.uint64)
    lo = prng.random_bits(k2, n).astype(np.uint64)
    span = np.uint64(maxval - minval if maxval > minval else 1)
    mult = (np.uint64(65536) % span)
    mult = ((mult * mult) & np.uint64(0xFFFFFFFF)) % span          # lax.mul(multiplier, multiplier) in uint32 wraps: 0 for span > 65536
    off = (((hi % span) * mult) & np.uint64(0xFFFFFFFF))          # lax.mul in uint32 wraps
    off = ((off + (lo % span)) & np.uint64(0xFFFFFFFF)) % span      # lax.add in uint32 wraps
    return (np.int64(minval) + off.astype(np.int64)).astype(np.int32)


def project_bounds(n, index_tree_train, trainable):
    lo, hi = np.full(n, -np.inf, np.float32), np.full(n, np.inf, np.float32)
    for name, (a, b) in {"eps": (1e-7, 0.5), "eta": (0.0, 0.99), "gamma": (0.001, np.inf), "mgridref_y": (0.001, np.inf)}.items():
        if name in trainable and name in index_tree_train:
            ix = np.asarray(index_tree_train[name]).reshape(-1).astype(np.int64)
            lo[ix], hi[ix] = a, b
    return lo, hi


def adam_project_step(p, g, m, v, count, lr, lo=None, hi=None, b1=0.9, b2=0.999, eps=1e-8, clip=5.0):
    """One optimizer.update + apply_updates + project in float32.  Returns (p, m, v)."""
    f = np.float32
    g = np.clip(g.astype(f), -f(clip), f(clip))
    m = (f(b1) * m + (f(1) - f(b1)) * g).astype(f)
    v = (f(b2) * v + (f(1) - f(b2)) * g * g).astype(f)
    mhat = m / f(1.0 - b1 ** count)
    vhat = v / f(1.0 - b2 ** count)
    p = (p - f(lr) * (mhat / (np.sqrt(vhat) + f(eps)))).astype(f)
    if lo is not None:
        p = np.maximum(p, lo)
    if hi is not None:
        p = np.minimum(p, hi)
    return p.astype(f), m, v
