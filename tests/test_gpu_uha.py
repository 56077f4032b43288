"""UHA (config.boundmode "UHA": boundingmachine.py:73-111 over ais_utils.py:7-69): CUDA path through the C ABI vs the CPU
oracle on identical seeds and parameters -- per-particle loss / final state within 1e-4 relative, gradients (vd, eps, eta,
md, mgridref_y) within max(1e-4, 2x the fp32 oracle's own distance to fp64) per pytree leaf."""
import os

import numpy as np
import pytest
import torch

from cmcd_b200 import boundingmachine as PB
from cmcd_b200 import mcdboundingmachine as PM
from cmcd_b200.pytree import tree_leaves
from oracle import mcdboundingmachine as OM
from helpers import UHA_CONFIGS, rel_err, seeds_for, uha_oracle_problem, uha_product_problem

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _leaf_errs(g, ref, unflatten):
    out = []
    for a, b in zip(tree_leaves(unflatten(g)), tree_leaves(unflatten(ref))):
        a, b = a.double().reshape(-1), b.double().reshape(-1)
        if b.numel() == 0:
            continue
        scale = b.abs().max().item()
        out.append((a - b).abs().max().item() / scale if scale > 0 else (a - b).abs().max().item())
    return np.array(out)


def _both(name, K=None):
    c, lp, dim, pf, unf, fixed = uha_oracle_problem(name, torch.float32, K=K)
    _, lp64, _, pf64, unf64, fixed64 = uha_oracle_problem(name, torch.float64, K=K)
    seeds = seeds_for(c["N"])
    g32, (l32, z32) = OM.grad_and_loss(OM.uha_compute_bound, seeds, pf, unf, fixed, lp)
    g64, (l64, z64) = OM.grad_and_loss(OM.uha_compute_bound, seeds, pf64, unf64, fixed64, lp64)
    _, target, _, pf_p, unf_p, fixed_p = uha_product_problem(name, pf, K=K)
    gp, (l_p, z_p) = PM.grad_and_loss(PB.compute_bound)(torch.from_numpy(seeds), pf_p, unf_p, fixed_p, target)
    return unf, (g32, l32, z32), (g64, l64, z64), (gp.cpu(), l_p.cpu(), z_p.cpu())


@pytest.mark.parametrize("name,K", [(n, None) for n in UHA_CONFIGS] + [("UHA_gmm", 1), ("UHA_funnel_lf3", 2)])
def test_uha_parity(name, K):
    unf, (g32, l32, z32), (g64, l64, z64), (gp, l_p, z_p) = _both(name, K)
    fin = torch.isfinite(l64)
    assert (torch.isfinite(l_p) == fin).all()
    e_l, e_l32 = rel_err(l_p[fin], l64[fin]).max(), rel_err(l32[fin], l64[fin]).max()
    e_z = rel_err(z_p[fin], z64[fin]).max()
    assert e_l < max(1e-4, 2 * e_l32) and e_z < max(1e-4, 2 * rel_err(z32[fin], z64[fin]).max()), (name, e_l, e_l32, e_z)
    assert torch.isfinite(gp).all()
    e_k, e_o = _leaf_errs(gp, g64, unf), _leaf_errs(g32, g64, unf)
    print(f"{name}: loss {e_l:.2e} (fp32 oracle {e_l32:.2e}); grad kernel-vs-fp64 max {e_k.max():.2e}, fp32-oracle-vs-fp64 {e_o.max():.2e}")
    assert (e_k <= np.maximum(1e-4, 2 * e_o)).all(), (name, e_k, e_o)


@pytest.mark.parametrize("name", sorted(UHA_CONFIGS))
def test_uha_reproduces_fixture(name):
    f = np.load(os.path.join(GOLDEN, f"uha_{name}.npz"))
    n = len(f["seeds"])
    c, target, dim, pf_p, unf_p, fixed_p = uha_product_problem(name, torch.from_numpy(f["params_flat"]), N=n)
    g, (l, z) = PM.grad_and_loss(PB.compute_bound)(torch.from_numpy(f["seeds"]), pf_p, unf_p, fixed_p, target)
    assert rel_err(l.cpu().numpy(), f["l"]).max() < 1e-4 and rel_err(z.cpu().numpy(), f["z"]).max() < 1e-4
    errs = _leaf_errs(g.cpu(), torch.from_numpy(f["grad"]), unf_p)
    assert (errs <= np.maximum(1e-4, 2 * f["grad_fp32_floor"])).all(), (name, errs, f["grad_fp32_floor"])


def test_uha_forward_only_and_partition_invariance():
    c, lp, dim, pf, unf, fixed = uha_oracle_problem("UHA_manygmm_lf2")
    seeds = torch.from_numpy(seeds_for(777))
    _, target, _, pf_p, unf_p, fixed_p = uha_product_problem("UHA_manygmm_lf2", pf)
    with torch.no_grad():
        full = PB.compute_bound(seeds, pf_p, unf_p, fixed_p, target)[1]
        parts = [PB.compute_bound(seeds[a:b], pf_p, unf_p, fixed_p, target)[1] for a, b in ((0, 130), (130, 777))]
    assert torch.equal(full[0], torch.cat([p[0] for p in parts])) and torch.equal(full[1], torch.cat([p[1] for p in parts]))


def test_uha_several_tiles_per_cta():
    """N past one wave of CTAs: the persistent UHA kernels walk several particle tiles per CTA."""
    unf, (g32, l32, z32), (g64, l64, z64), (gp, l_p, z_p) = _both_n("UHA_gmm", 90000, 2)
    assert rel_err(l_p, l64).max() < max(1e-4, 2 * rel_err(l32, l64).max())
    e_k, e_o = _leaf_errs(gp, g64, unf), _leaf_errs(g32, g64, unf)
    assert (e_k <= np.maximum(1e-4, 2 * e_o)).all(), (e_k, e_o)


def _both_n(name, N, K):
    c, lp, dim, pf, unf, fixed = uha_oracle_problem(name, torch.float32, N=N, K=K)
    _, lp64, _, pf64, unf64, fixed64 = uha_oracle_problem(name, torch.float64, N=N, K=K)
    seeds = seeds_for(N)
    g32, (l32, z32) = OM.grad_and_loss(OM.uha_compute_bound, seeds, pf, unf, fixed, lp)
    g64, (l64, z64) = OM.grad_and_loss(OM.uha_compute_bound, seeds, pf64, unf64, fixed64, lp64)
    _, target, _, pf_p, unf_p, fixed_p = uha_product_problem(name, pf, N=N, K=K)
    gp, (l_p, z_p) = PM.grad_and_loss(PB.compute_bound)(torch.from_numpy(seeds), pf_p, unf_p, fixed_p, target)
    return unf, (g32, l32, z32), (g64, l64, z64), (gp.cpu(), l_p.cpu(), z_p.cpu())
