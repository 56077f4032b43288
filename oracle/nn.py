"""Drift networks of the CMCD hot path restated in torch-CPU (batched over particles).

ORACLE / TEST INFRASTRUCTURE ONLY.  Follows /root/reference/src/nn.py:17-72 ("geffner"
residual-softplus net with a learned per-step embedding table) and
/root/reference/src/nn_dds.py:55-70,91-192 (PISNet / "dds": sin-cos time embedding ->
time coder MLP, state net with exact-erf GELU, zero-initialised head, output clip 1e4).
Parameter containers are plain nested dicts (haiku/stax pytrees are not reproduced).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- geffner (nn.py:42-72)
def init_geffner(x_dim, emb_dim, nbridges, gen, live=False, dtype=torch.float32, rho_dim=0):
    """stax Dense init: glorot-normal W [in,out], b ~ 1e-2*N(0,1); emb ~ 0.05 N(0,1) (nn.py:17-18);
    factor_sn = 0 (nn.py:63).  ``live=True`` sets factor_sn=0.1 so the drift path is numerically
    active (SURVEY section 8d synthetic-input convention)."""
    in_dim = x_dim + rho_dim + emb_dim   # nn.py:43 (rho_dim = dim for the (z, rho) networks of the underdamped modes)

    def dense(i, o):
        std = math.sqrt(2.0 / (i + o))
        return {"w": torch.randn(i, o, generator=gen, dtype=dtype) * std,
                "b": torch.randn(o, generator=gen, dtype=dtype) * 1e-2}

    return {"nn": [dense(in_dim, in_dim), dense(in_dim, in_dim), dense(in_dim, x_dim)],
            "emb": torch.randn(nbridges, emb_dim, generator=gen, dtype=dtype) * 0.05,
            "factor_sn": torch.tensor(0.1 if live else 0.0, dtype=dtype)}


def apply_geffner(params, x, i):
    """nn.py:66-70.  x [N,d]; i python int.  emb[i] with JAX's clamped gather (i = nbridges -> last row)."""
    emb = params["emb"]
    row = emb[min(max(int(i), 0), emb.shape[0] - 1)]
    h = torch.cat([x, row[None].expand(x.shape[0], -1)], dim=-1)
    l1, l2, l3 = params["nn"]
    h = h + F.softplus(h @ l1["w"] + l1["b"])
    h = h + F.softplus(h @ l2["w"] + l2["b"])
    return (h @ l3["w"] + l3["b"]) * params["factor_sn"]


# --------------------------------------------------------------------------- dds / PISNet (nn_dds.py:91-192)
DDS_CHANNELS = 64


def dds_timestep_coeff(dtype=torch.float32):
    """np.linspace(0.1, 100, 64)[None] in float32 (nn_dds.py:108)."""
    return torch.tensor(np.linspace(0.1, 100.0, DDS_CHANNELS).astype(np.float32), dtype=dtype)


def init_dds(x_dim, gen, live=False, dtype=torch.float32, rho_dim=0):
    """haiku Linear init: W ~ truncated-normal(1/sqrt(fan_in)), b = 0; LinearZero head zeros
    (nn_dds.py:179-192).  ``live=True``: head W ~ 0.01 N(0,1) so the drift is non-zero."""
    c = DDS_CHANNELS

    def lin(i, o):
        w = torch.empty(i, o, dtype=dtype)
        torch.nn.init.trunc_normal_(w, std=1.0 / math.sqrt(i), a=-2.0 / math.sqrt(i), b=2.0 / math.sqrt(i), generator=gen)
        return {"w": w, "b": torch.zeros(o, dtype=dtype)}

    head = {"w": torch.zeros(c, x_dim, dtype=dtype), "b": torch.zeros(x_dim, dtype=dtype)}
    if live:
        head["w"] = torch.randn(c, x_dim, generator=gen, dtype=dtype) * 0.01
        head["b"] = torch.randn(x_dim, generator=gen, dtype=dtype) * 0.01
    return {"timestep_phase": torch.zeros(1, c, dtype=dtype) if not live else torch.randn(1, c, generator=gen, dtype=dtype) * 0.1,
            "tc1": lin(2 * c, c), "tc2": lin(c, c),
            "st1": lin(x_dim + rho_dim + c, c), "st2": lin(c, c), "out": head}   # nn_dds.py:56,159


def gelu_exact(x):
    """nn_dds.py:167-176: x * 0.5 * (1 + erf(x / sqrt(2)))."""
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def dds_time_embedding(params, t):
    """nn_dds.py:130-143,156-158 for an integer step index t -> t_net_1 [64]."""
    coeff = dds_timestep_coeff(params["timestep_phase"].dtype)
    tf = torch.tensor(float(int(t)), dtype=coeff.dtype)
    arg = coeff[None] * tf + params["timestep_phase"]
    e = torch.cat([torch.sin(arg), torch.cos(arg)], dim=-1)  # [1,128]
    h = gelu_exact(e @ params["tc1"]["w"] + params["tc1"]["b"])
    return (h @ params["tc2"]["w"] + params["tc2"]["b"]).reshape(-1)


def apply_dds(params, x, t):
    """nn_dds.py:145-164.  x [N,d]; t python int (the reference passes the scan index, unbatched)."""
    t_net = dds_time_embedding(params, t)
    h = torch.cat([x, t_net[None].expand(x.shape[0], -1)], dim=-1)
    h = gelu_exact(h @ params["st1"]["w"] + params["st1"]["b"])
    h = gelu_exact(h @ params["st2"]["w"] + params["st2"]["b"])
    out = h @ params["out"]["w"] + params["out"]["b"]
    return torch.clamp(out, -1.0e4, 1.0e4)


def initialize_network(x_dim, emb_dim, nbridges, nn_arch="geffner", gen=None, live=False, dtype=torch.float32, rho_dim=0):
    """nn.py:21-39 -> (init_params, apply_fun(params, x, i)); x = [z] or [z, rho] (rho_dim > 0)."""
    gen = gen or torch.Generator().manual_seed(1)
    if nn_arch == "geffner":
        return init_geffner(x_dim, emb_dim, nbridges, gen, live, dtype, rho_dim), apply_geffner
    if nn_arch == "dds":
        return init_dds(x_dim, gen, live, dtype, rho_dim), apply_dds
    raise NotImplementedError(f"nn_arch {nn_arch!r}: dds_grad is broken in the reference (SURVEY section 2 row 7)")
