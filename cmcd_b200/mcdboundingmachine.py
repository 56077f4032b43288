"""The MCD bound -- host side.  Mirrors /root/reference/src/mcdboundingmachine.py.

``initialize`` (:11-123) builds the same (params_flat, unflatten, params_fixed) triple;
``compute_log_elbo`` (:126-179), ``compute_bound`` (:183-205) and ``compute_bound_var`` (:208-231)
keep the reference's argument order and return values.  ``grad_and_loss`` is the equivalent of
``jax.jit(jax.grad(compute_bound_fn, 1, has_aux=True))`` (main.py:174-176).
All O(N*K) work happens in the CUDA kernels behind ``mcd_utils.bridge``.
"""
from __future__ import annotations

import torch

from . import mcd_utils
from . import variationaldist as vd
from .nn import initialize_network
from .pytree import ravel_pytree, tree_map

_SN_MODES = ("MCD_ULA_sn", "MCD_CAIS_sn", "MCD_CAIS_var_sn", "MCD_U_a-lp-sna", "MCD_U_e-lp-sna")   # mcdboundingmachine.py:67-83
_SN_RHO_MODES = ("MCD_U_a-lp-sn", "MCD_U_ea-lp-sn", "MCD_CAIS_UHA_sn")                               # :84-102: network on (z, rho), rho_dim = dim


def initialize(dim, vdparams=None, nbridges=0, eps=0.01, gamma=10.0, eta=0.5, ngridb=32, mgridref_y=None,
               trainable=("eps",), use_score_nn=True, emb_dim=48, nlayers=3, seed=1, mode="MCD_CAIS_sn",
               nn_arch="dds", fully_connected_units=None, device="cuda"):
    """mcdboundingmachine.py:11-123."""
    if mode not in mcd_utils.SUPPORTED_MODES:
        raise NotImplementedError(f"Mode {mode} not implemented (hot-path scope: {mcd_utils.SUPPORTED_MODES}).")
    dev = torch.device(device)
    pt, pn = {}, {}
    f = lambda v: torch.tensor(float(v), dtype=torch.float32, device=dev)
    vdp = vdparams if vdparams is not None else vd.initialize(dim, device=dev)
    (pt if "vd" in trainable else pn)["vd"] = tree_map(lambda t: t.to(dev), vdp)
    for name, val in (("eps", eps), ("gamma", gamma), ("eta", eta)):
        (pt if name in trainable else pn)[name] = f(val)
    if mode in _SN_MODES:
        init_fun_sn, apply_fun_sn = initialize_network(dim, emb_dim, nbridges, nlayers=nlayers, nn_arch=nn_arch,
                                                       fully_connected_units=fully_connected_units)
        pt["sn"] = init_fun_sn(seed, None, device=dev)[1]
    elif mode in _SN_RHO_MODES:
        init_fun_sn, apply_fun_sn = initialize_network(dim, emb_dim, nbridges, rho_dim=dim, nlayers=nlayers, nn_arch=nn_arch,
                                                       fully_connected_units=fully_connected_units)
        pt["sn"] = init_fun_sn(seed, None, device=dev)[1]
    else:
        apply_fun_sn = None
    if mgridref_y is not None:
        ngridb = mgridref_y.shape[0] - 1
        mgridref_y = mgridref_y.to(dev)
    else:
        ngridb = min(ngridb, nbridges)
        mgridref_y = torch.ones(ngridb + 1, device=dev)
    pn["gridref_x"] = torch.linspace(0, 1, ngridb + 2, device=dev)
    pn["target_x"] = torch.linspace(0, 1, nbridges + 2, device=dev)[1:-1]
    (pt if "mgridref_y" in trainable else pn)["mgridref_y"] = mgridref_y
    params_fixed = (dim, nbridges, mode, apply_fun_sn)
    params_flat, unflatten = ravel_pytree((pt, pn), device=dev)
    return params_flat, unflatten, params_fixed


def _interp(x, xp, fp):
    """jnp.interp (differentiable in fp)."""
    i = torch.clamp(torch.searchsorted(xp, x.contiguous(), right=True), 1, xp.numel() - 1)
    df, dx, delta = fp[i] - fp[i - 1], xp[i] - xp[i - 1], x - xp[i - 1]
    return torch.where(dx == 0, fp[i], fp[i - 1] + (delta / dx) * df)


def make_betas(params):
    """mcdboundingmachine.py:146-149."""
    m = params["mgridref_y"]
    gridref_y = torch.cat([torch.zeros(1, device=m.device), torch.cumsum(m, 0) / torch.sum(m)])
    return _interp(params["target_x"], params["gridref_x"], gridref_y)


def compute_log_elbo(seed, params_flat, unflatten, params_fixed, log_prob, eps_schedule=None, grad_clipping=False):
    """mcdboundingmachine.py:126-179, batched: ``seed`` may be an int32 vector -> (-w[N], (z[N,d], None))."""
    scalar = (seed.dim() == 0) if isinstance(seed, torch.Tensor) else not hasattr(seed, "__len__")
    seeds = torch.as_tensor(seed, dtype=torch.int32).reshape(-1) if scalar else torch.as_tensor(seed, dtype=torch.int32)
    if mcd_utils.chain_supported(params_flat, unflatten, params_fixed, eps_schedule):
        # betas / eps schedule / per-step network tables and their transposes in the fused chain kernels (csrc/chain.cu);
        # stop_gradient(params_notrain) (:142) is the chain's train mask
        negw, z = mcd_utils.fused_bridge(seeds, params_flat, unflatten, params_fixed, log_prob, eps_schedule, grad_clipping)
        return (negw[0], (z[0], None)) if scalar else (negw, (z, None))
    pt, pn = unflatten(params_flat)
    pn = tree_map(lambda t: t.detach(), pn)  # jax.lax.stop_gradient(params_notrain) :142
    params = {**pt, **pn}
    nbridges = params_fixed[1]
    betas = make_betas(params) if nbridges >= 1 else None
    negw, z = mcd_utils.bridge(seeds, params, betas, params_fixed, log_prob, eps_schedule, grad_clipping)
    return (negw[0], (z[0], None)) if scalar else (negw, (z, None))


def compute_bound(seeds, params_flat, unflatten, params_fixed, log_prob, eps_schedule=None, grad_clipping=False):
    """mcdboundingmachine.py:183-205 -> (mean loss, (loss[N], z[N,d]))."""
    l, (z, _) = compute_log_elbo(seeds, params_flat, unflatten, params_fixed, log_prob, eps_schedule, grad_clipping)
    return l.mean(), (l, z)


def compute_bound_var(seeds, params_flat, unflatten, params_fixed, log_prob, eps_schedule=None,
                      grad_clipping=False, ln_Z_correction=False):
    """mcdboundingmachine.py:208-231 -> (clip(var(loss, ddof=0), +-1e7), (loss[N], z[N,d]))."""
    l, (z, _) = compute_log_elbo(seeds, params_flat, unflatten, params_fixed, log_prob, eps_schedule, grad_clipping)
    return torch.clamp(l.var(unbiased=False), -1e7, 1e7), (l, z)


def grad_and_loss(compute_bound_fn):
    """jax.grad(compute_bound_fn, 1, has_aux=True) (main.py:174-176):
    f(seeds, params_flat, unflatten, params_fixed, log_prob) -> (grad_flat, (loss[N], z[N,d]))."""

    def f(seeds, params_flat, unflatten, params_fixed, log_prob):
        p = params_flat.detach().requires_grad_(True)
        with torch.enable_grad():
            loss, (l, z) = compute_bound_fn(seeds, p, unflatten, params_fixed, log_prob)
            (g,) = torch.autograd.grad(loss, p, allow_unused=True)
        return (torch.zeros_like(p) if g is None else g), (l.detach(), z.detach())

    return f
