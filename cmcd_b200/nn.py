"""Drift networks -- host side of the fused kernel (parameter containers + per-step tables).

Mirrors /root/reference/src/nn.py:21-72 (``initialize_network``, "geffner" net) and
src/nn_dds.py:55-70,91-192 (PISNet, "dds").  The per-particle evaluation
``apply_fun(params, x, i)`` happens inside the CUDA bridge kernels; this module only does the
O(K) work that depends on the step index alone -- the embedding row / sinusoidal time code
pushed through the time-coder MLP and the first-layer weights -- and packs it into the
"table form" consumed by the C ABI (include/cmcd_b200.h, ``cmcd_net``).  It is written with
differentiable torch ops so the kernel's table cotangents chain back into the raw parameters.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

DDS_CHANNELS = 64  # nn_dds.py:95 hard-codes fully_connected_units = [64, 64]


def _pad8(n):
    return (n + 7) // 8 * 8


class ApplyFun:
    """Tag returned in ``params_fixed[3]`` (the reference stores the apply function there)."""

    def __init__(self, arch, x_dim, emb_dim, nbridges, rho_dim=0):
        self.arch, self.x_dim, self.emb_dim, self.nbridges, self.rho_dim = arch, x_dim, emb_dim, nbridges, rho_dim
        self.in_dim = x_dim + rho_dim   # per-particle input: z, or (z, rho) for the underdamped networks (nn.py:43, nn_dds.py:56)
        self.hidden = DDS_CHANNELS if arch == "dds" else self.in_dim + emb_dim
        self.hidden_pad = _pad8(self.hidden)

    def __call__(self, params, inputs, i, **kwargs):
        raise RuntimeError("apply_fun_sn is evaluated inside the fused CUDA bridge kernel; "
                           "call mcd_utils.evolve / mcdboundingmachine.compute_bound instead")

    def __hash__(self):
        return hash((self.arch, self.x_dim, self.emb_dim, self.nbridges, self.rho_dim))

    def __eq__(self, o):
        return isinstance(o, ApplyFun) and (self.arch, self.x_dim, self.emb_dim, self.nbridges, self.rho_dim) == \
            (o.arch, o.x_dim, o.emb_dim, o.nbridges, o.rho_dim)


# ------------------------------------------------------------------ init (same distributions as stax / haiku)
def init_geffner(x_dim, emb_dim, nbridges, gen, device=None, dtype=torch.float32, rho_dim=0):
    """nn.py:42-64: Dense = glorot-normal W [in,out] + 1e-2 N(0,1) b; emb = 0.05 N(0,1); factor_sn = 0."""
    in_dim = x_dim + rho_dim + emb_dim

    def dense(i, o):
        std = math.sqrt(2.0 / (i + o))
        return {"w": (torch.randn(i, o, generator=gen, dtype=dtype) * std).to(device),
                "b": (torch.randn(o, generator=gen, dtype=dtype) * 1e-2).to(device)}

    return {"nn": [dense(in_dim, in_dim), dense(in_dim, in_dim), dense(in_dim, x_dim)],
            "emb": (torch.randn(nbridges, emb_dim, generator=gen, dtype=dtype) * 0.05).to(device),
            "factor_sn": torch.tensor(0.0, dtype=dtype, device=device)}


def init_dds(x_dim, gen, device=None, dtype=torch.float32, rho_dim=0):
    """nn_dds.py:91-127,179-192: haiku Linear (trunc-normal 1/sqrt(fan_in), zero bias); zero head; zero phase."""
    c = DDS_CHANNELS

    def lin(i, o):
        w = torch.empty(i, o, dtype=dtype)
        s = 1.0 / math.sqrt(i)
        torch.nn.init.trunc_normal_(w, std=s, a=-2 * s, b=2 * s, generator=gen)
        return {"w": w.to(device), "b": torch.zeros(o, dtype=dtype, device=device)}

    return {"timestep_phase": torch.zeros(1, c, dtype=dtype, device=device),
            "tc1": lin(2 * c, c), "tc2": lin(c, c), "st1": lin(x_dim + rho_dim + c, c), "st2": lin(c, c),
            "out": {"w": torch.zeros(c, x_dim, dtype=dtype, device=device),
                    "b": torch.zeros(x_dim, dtype=dtype, device=device)}}


def initialize_network(x_dim, emb_dim, nbridges, rho_dim=0, nlayers=4, nn_arch="geffner",
                       fully_connected_units=None):
    """nn.py:21-39 -> (init_fun(rng, input_shape) -> (None, params), apply_fun)."""
    if nn_arch not in ("geffner", "dds"):
        raise NotImplementedError(f"nn_arch {nn_arch!r} not implemented (dds_grad is broken in the reference)")
    apply_fun = ApplyFun(nn_arch, x_dim, emb_dim, nbridges, rho_dim)

    def init_fun(rng, input_shape=None, device=None):
        gen = rng if isinstance(rng, torch.Generator) else torch.Generator().manual_seed(int(rng))
        if nn_arch == "geffner":
            return None, init_geffner(x_dim, emb_dim, nbridges, gen, device, rho_dim=rho_dim)
        return None, init_dds(x_dim, gen, device, rho_dim=rho_dim)

    return init_fun, apply_fun


# ------------------------------------------------------------------ per-step tables
def _gelu(x):
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def _pad_cols(t, hp):
    return t if t.shape[-1] == hp else F.pad(t, (0, hp - t.shape[-1]))


_COEFF_CACHE = {}


def dds_timestep_coeff(device, dtype=torch.float32):
    """linspace(0.1, 100, 64) (nn_dds.py:108), cached per device so that build_tables does no host->device copy."""
    import numpy as np
    key = (str(device), dtype)
    if key not in _COEFF_CACHE:
        _COEFF_CACHE[key] = torch.tensor(np.linspace(0.1, 100.0, DDS_CHANNELS).astype(np.float32), dtype=dtype, device=device)
    return _COEFF_CACHE[key]


def build_tables(apply_fun: ApplyFun, sn):
    """params["sn"] -> dict(U1,U2,U3,W2,W3,c1,c2,c3,out_scale) in the layout of ``cmcd_net`` (differentiable)."""
    d, K, hp = apply_fun.in_dim, apply_fun.nbridges, apply_fun.hidden_pad   # d: rows of the weights that see the particle
    if apply_fun.arch == "geffner":
        (l1, l2, l3) = sn["nn"]
        rows = torch.clamp(torch.arange(K + 1, device=sn["emb"].device), max=K - 1)  # JAX clamps emb[K] (nn.py:68)
        e = sn["emb"][rows]  # [K+1, E]
        h = apply_fun.hidden
        W2 = F.pad(l2["w"], (0, hp - h, 0, hp - h))
        W3 = F.pad(l3["w"], (0, 0, 0, hp - h))
        return {"U1": _pad_cols(l1["w"][:d], hp), "U2": _pad_cols(l2["w"][:d], hp), "U3": l3["w"][:d],
                "W2": W2, "W3": W3,
                "c1": _pad_cols(e @ l1["w"][d:] + l1["b"], hp), "c2": _pad_cols(e @ l2["w"][d:] + l2["b"], hp),
                "c3": e @ l3["w"][d:] + l3["b"], "out_scale": sn["factor_sn"].reshape(1)}
    # dds: sinusoidal code -> time coder (nn_dds.py:130-143,156-158), t = integer step index
    dev = sn["timestep_phase"].device
    t = torch.arange(K + 1, device=dev, dtype=torch.float32)[:, None]
    arg = dds_timestep_coeff(dev)[None] * t + sn["timestep_phase"]
    code = torch.cat([torch.sin(arg), torch.cos(arg)], dim=-1)
    t_net = _gelu(code @ sn["tc1"]["w"] + sn["tc1"]["b"]) @ sn["tc2"]["w"] + sn["tc2"]["b"]  # [K+1, 64]
    w1 = sn["st1"]["w"]
    return {"U1": w1[:d], "U2": None, "U3": None, "W2": sn["st2"]["w"], "W3": sn["out"]["w"],
            "c1": t_net @ w1[d:] + sn["st1"]["b"], "c2": sn["st2"]["b"][None].expand(K + 1, -1),
            "c3": sn["out"]["b"][None].expand(K + 1, -1),
            "out_scale": torch.ones(1, device=dev, dtype=torch.float32)}
