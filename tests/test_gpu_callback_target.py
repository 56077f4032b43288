"""Generic targets: a density outside the fused registry, handed over as a batched torch
log-density and evaluated through the C ABI's score callback (cmcd_target_fn) between the half-steps of the step-wise CUDA
path -- what the reference does with jax.grad of an arbitrary log_prob_model (numpyro logistic regression / inference-gym
models, model_handler.py:46-86).  Checked against the CPU oracle running the SAME torch density: per-particle loss / final
state within 1e-4 relative, gradients within max(1e-4, 2x the fp32 oracle's distance to fp64) per pytree leaf."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from cmcd_b200 import mcdboundingmachine as PM
from cmcd_b200 import model_handler as PH
from cmcd_b200 import variationaldist as PV
from cmcd_b200.pytree import tree_leaves
from oracle import mcdboundingmachine as OM
from helpers import rel_err, seeds_for

pytestmark = pytest.mark.gpu


def logistic_regression(dim, n_data=60, device="cpu", dtype=torch.float32):
    """Bayesian logistic regression with a N(0, I) prior on the weights (the structure of models/logistic_regression.py
    in the reference; synthetic design matrix -- the sonar / ionosphere files are not part of the hot path)."""
    g = torch.Generator().manual_seed(3)
    X = torch.randn(n_data, dim, generator=g)
    theta = torch.randn(dim, generator=g) * 0.5
    y = (torch.rand(n_data, generator=g) < torch.sigmoid(X @ theta)).float() * 2 - 1
    X, y = X.to(device=device, dtype=dtype), y.to(device=device, dtype=dtype)

    def log_prob(th):   # th [N, d] -> [N]
        return F.logsigmoid(y * (th @ X.T)).sum(-1) - 0.5 * (th ** 2).sum(-1) - 0.5 * dim * math.log(2 * math.pi)
    return log_prob


CASES = {   # name -> (dim, mode, K, N, eps, eps_schedule, clip, emb_dim)
    "logreg25_cais": (25, "MCD_CAIS_sn", 6, 150, 0.01, "cos_sq", True, 20),
    "logreg13_var": (13, "MCD_CAIS_var_sn", 4, 100, 0.02, None, True, 11),
    "logreg25_ulasn": (25, "MCD_ULA_sn", 5, 120, 0.01, None, False, 20),
    "logreg16_ula": (16, "MCD_ULA", 4, 100, 0.02, None, False, 20),
}
TRAINABLE = ("eta", "gamma", "eps", "vd", "mgridref_y")


def _problem(name, dtype):
    dim, mode, K, N, eps, sched, clip, emb = CASES[name]
    g = torch.Generator().manual_seed(7)
    vdp = OM.vd_initialize(dim, 0.7)
    vdp["mean"] = vdp["mean"] + 0.1 * torch.randn(dim, generator=g)
    mgrid = 1.0 + 0.3 * torch.rand(min(32, K) + 1, generator=g)
    pf, unf, fixed = OM.initialize(dim, vdparams=vdp, nbridges=K, eps=eps, trainable=TRAINABLE, emb_dim=emb, mode=mode,
                                   nn_arch="geffner", mgridref_y=mgrid, live=True)
    return pf.to(dtype), unf, fixed, logistic_regression(dim, dtype=dtype), dict(eps_schedule=sched, grad_clipping=clip)


def _leaf_errs(g, ref, unflatten):
    out = []
    for a, b in zip(tree_leaves(unflatten(g)), tree_leaves(unflatten(ref))):
        a, b = a.double().reshape(-1), b.double().reshape(-1)
        if b.numel() == 0:
            continue
        scale = b.abs().max().item()
        out.append((a - b).abs().max().item() / scale if scale > 0 else (a - b).abs().max().item())
    return np.array(out)


@pytest.mark.parametrize("name", sorted(CASES))
def test_callback_target_parity(name):
    dim, mode, K, N, eps, sched, clip, emb = CASES[name]
    seeds = seeds_for(N)
    var = "var" in mode
    out = {}
    for dtype in (torch.float32, torch.float64):
        pf, unf, fixed, lp, kw = _problem(name, dtype)
        out[dtype] = OM.grad_and_loss(OM.compute_bound_var if var else OM.compute_bound, seeds, pf, unf, fixed, lp, **kw)
    (g32, (l32, z32)), (g64, (l64, z64)) = out[torch.float32], out[torch.float64]
    pf, unf, fixed, _, kw = _problem(name, torch.float32)
    target, _ = PH.callback_target(logistic_regression(dim, device="cuda"), dim)
    pf_p, unf_p, fixed_p = PM.initialize(dim, vdparams=PV.initialize(dim, device="cuda"), nbridges=K, eps=eps, trainable=TRAINABLE,
                                         emb_dim=emb, mode=mode, nn_arch="geffner", mgridref_y=torch.ones(min(32, K) + 1), device="cuda")
    assert pf_p.numel() == pf.numel()
    fn = PM.compute_bound_var if var else PM.compute_bound
    gp, (l_p, z_p) = PM.grad_and_loss(lambda *a: fn(*a, **kw))(torch.from_numpy(seeds), pf.cuda(), unf_p, fixed_p, target)
    gp, l_p, z_p = gp.cpu(), l_p.cpu(), z_p.cpu()
    assert target.calls >= 2 * (K + 1)   # the callback really ran: K + 1 scores forward, K + 1 scores (+ HVPs) in reverse
    e_l, e_l32 = rel_err(l_p, l64).max(), rel_err(l32, l64).max()
    assert e_l < max(1e-4, 2 * e_l32), (name, e_l, e_l32)
    assert rel_err(z_p, z64).max() < max(1e-4, 2 * rel_err(z32, z64).max())
    e_k, e_o = _leaf_errs(gp, g64, unf), _leaf_errs(g32, g64, unf)
    print(f"{name}: loss {e_l:.2e} (fp32 oracle {e_l32:.2e}); grad kernel-vs-fp64 max {e_k.max():.2e}, fp32-oracle-vs-fp64 {e_o.max():.2e}")
    assert torch.isfinite(gp).all()
    assert (e_k <= np.maximum(1e-4, 2 * e_o)).all(), (name, e_k, e_o)


def test_callback_exception_surfaces():
    def bad(th):
        raise ValueError("density blew up")
    target, dim = PH.callback_target(bad, 8)
    pf, unf, fixed = PM.initialize(dim, nbridges=2, trainable=("vd",), mode="MCD_ULA", device="cuda")
    with pytest.raises(ValueError, match="blew up"):
        PM.compute_bound(torch.arange(1, 9), pf, unf, fixed, target)
