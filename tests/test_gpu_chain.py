"""The fused O(K) chain (csrc/chain.cu: betas, eps schedule, per-step network tables and their transposes in one prologue and two
epilogue kernels) against the framework chain it replaces (make_betas / eps_table / nn.build_tables under torch autograd,
CMCD_DISABLE_CHAIN=1) on identical inputs: same losses and the same flat gradient up to the summation order of the small
products.  (Every oracle parity test runs through the fused chain as well -- it is the default path.)"""
import pytest
import torch

from cmcd_b200 import mcd_utils as PU
from cmcd_b200 import mcdboundingmachine as PM
from cmcd_b200.pytree import tree_leaves
from helpers import oracle_problem, product_problem, seeds_for

pytestmark = pytest.mark.gpu

NAMES = ["A_gmm", "B_funnel", "C_manygmm_dds_small", "Cvar_manygmm", "Ckl_manygmm_geffner", "ULA_gmm", "ULAsn_funnel", "ULAsn_gmm_dds",
         "lin_funnel", "D_lgcp"]


def _run(name, monkeypatch, disable):
    if disable:
        monkeypatch.setenv("CMCD_DISABLE_CHAIN", "1")
    else:
        monkeypatch.delenv("CMCD_DISABLE_CHAIN", raising=False)
    c, lp, dim, pf, unf, fixed = oracle_problem(name, torch.float32)
    _, target, _, pf_p, unf_p, fixed_p = product_problem(name, pf)
    assert PU.chain_supported(pf_p, unf_p, fixed_p, c["eps_schedule"]) == (not disable)
    kw = dict(eps_schedule=c["eps_schedule"], grad_clipping=c["clip"])
    fn = PM.compute_bound_var if "var" in c["mode"] else PM.compute_bound
    g, (l, z) = PM.grad_and_loss(lambda *a: fn(*a, **kw))(torch.from_numpy(seeds_for(c["N"])), pf_p, unf_p, fixed_p, target)
    return unf_p, g.cpu(), l.cpu(), z.cpu()


@pytest.mark.parametrize("name", NAMES)
def test_fused_chain_matches_framework_chain(name, monkeypatch):
    unf, g0, l0, z0 = _run(name, monkeypatch, disable=True)
    _, g1, l1, z1 = _run(name, monkeypatch, disable=False)
    fin = torch.isfinite(l0)
    assert (torch.isfinite(l1) == fin).all()
    assert ((l1 - l0)[fin].abs() / l0[fin].abs().clamp(min=1)).max().item() < 1e-4
    assert ((z1 - z0)[fin].abs() / z0[fin].abs().clamp(min=1)).max().item() < 1e-4
    for a, b in zip(tree_leaves(unf(g1)), tree_leaves(unf(g0))):
        if b.numel() == 0:
            continue
        scale = b.abs().max().item()
        err = (a - b).abs().max().item()
        assert err <= 5e-5 * scale + 1e-12, (name, tuple(b.shape), err, scale)
    # frozen leaves: exactly zero, like stop_gradient(params_notrain) (mcdboundingmachine.py:142)
    assert all((leaf == 0).all() for leaf in tree_leaves(unf(g1)[1]))


def test_fused_chain_launch_count(monkeypatch):
    """One train iteration of the README gmm config: prologue + forward bridge + adjoint (+ its reduce) + two epilogue kernels of
    this library -- the ~100 framework launches of the table / betas / schedule chain are gone."""
    from cmcd_b200 import _lib
    monkeypatch.delenv("CMCD_DISABLE_CHAIN", raising=False)
    c, lp, dim, pf, unf, fixed = oracle_problem("A_gmm", torch.float32)
    _, target, _, pf_p, unf_p, fixed_p = product_problem("A_gmm", pf)
    gl = PM.grad_and_loss(lambda *a: PM.compute_bound(*a, eps_schedule=c["eps_schedule"], grad_clipping=c["clip"]))
    seeds = torch.from_numpy(seeds_for(c["N"]))
    gl(seeds, pf_p, unf_p, fixed_p, target)
    _lib.LAUNCHES["count"] = 0
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        gl(seeds, pf_p, unf_p, fixed_p, target)
        torch.cuda.synchronize()
    assert _lib.LAUNCHES["count"] == 6
    kernels = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "memset" not in e.name.lower() and "memcpy" not in e.name.lower()]
    assert len(kernels) <= 12, [e.name for e in kernels]
