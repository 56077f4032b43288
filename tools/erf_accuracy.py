"""Accuracy of the erfc-based GELU used by the kernels (csrc/common.cuh erfc_pos / norm_cdf) vs the exact GELU and
vs the reference's own fp32 formula x*0.5*(1+erf(x/sqrt 2)) (nn_dds.py:167-176).  Pure numpy float32 emulation."""
import numpy as np
from scipy.special import erf, erfc

f = np.float32


def erfc_pos32(z):
    z = z.astype(f)
    t = f(1) / (f(1) + f(0.5) * z)
    co = [0.17087277, -0.82215223, 1.48851587, -1.13520398, 0.27886807, -0.18628806, 0.09678418, 0.37409196, 1.00002368]
    p = f(co[0])
    for c in co[1:]:
        p = (p * t + f(c)).astype(f)
    arg = (-(z * z) - f(1.26551223) + t * p).astype(f)
    return (t * np.exp(arg.astype(np.float64)).astype(f)).astype(f)


if __name__ == "__main__":
    x = np.linspace(-8, 8, 4_000_001)
    exact = x * 0.5 * erfc(-x / np.sqrt(2))
    e = erfc_pos32(np.abs(x) * 0.70710678)
    fast = (x.astype(f) * np.where(x < 0, f(0.5) * e, f(1) - f(0.5) * e).astype(f)).astype(f)
    ref32 = (x.astype(f) * f(0.5) * (f(1) + erf((x.astype(f) * f(0.70710678)).astype(np.float64)).astype(f))).astype(f)
    print("kernel GELU   : max abs err %.3e" % np.abs(fast - exact).max())
    print("reference fp32: max abs err %.3e" % np.abs(ref32 - exact).max())
    print("kernel vs reference-fp32: max abs diff %.3e" % np.abs(fast - ref32).max())
