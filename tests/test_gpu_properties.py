"""Size-independent properties of the CUDA path at sizes the oracle cannot reach (up to 2^17 particles, K = 64), plus
the edge cases of the particle axis (N = 1, ragged tiles, K = 1, empty batch).  All calls go through the C ABI."""
import numpy as np
import pytest
import torch

from helpers import oracle_problem, product_problem, rel_err, seeds_for
from oracle import mcdboundingmachine as OM

pytestmark = pytest.mark.gpu


def _big_problem(K=64):
    from cmcd_b200 import mcdboundingmachine as PM
    c, lp, dim, pf, unf, fixed = oracle_problem("C_manygmm_dds_small", K=K)
    _, target, _, pf_p, unf_p, fixed_p = product_problem("C_manygmm_dds_small", pf, K=K)
    kw = dict(eps_schedule=c["eps_schedule"], grad_clipping=c["clip"])
    return PM, target, pf_p, unf_p, fixed_p, kw


def test_forward_is_deterministic_and_partition_invariant_at_scale():
    PM, target, pf, unf, fixed, kw = _big_problem()
    n = (1 << 17) + 77                      # ragged last tile
    seeds = torch.from_numpy(seeds_for(n, seed=3))
    with torch.no_grad():
        l1, z1 = PM.compute_bound(seeds, pf, unf, fixed, target, **kw)[1]
        l2, z2 = PM.compute_bound(seeds, pf, unf, fixed, target, **kw)[1]
        cut = 50_001                        # odd split: different tiles, different CTAs
        la, za = PM.compute_bound(seeds[:cut], pf, unf, fixed, target, **kw)[1]
        lb, zb = PM.compute_bound(seeds[cut:], pf, unf, fixed, target, **kw)[1]
    assert torch.equal(l1, l2) and torch.equal(z1, z2)
    assert torch.equal(l1, torch.cat([la, lb])) and torch.equal(z1, torch.cat([za, zb]))
    # identical seeds give identical particles (the reference draws seeds with replacement, opt.py:94)
    dup = torch.cat([seeds[:1000], seeds[:1000]])
    with torch.no_grad():
        ld = PM.compute_bound(dup, pf, unf, fixed, target, **kw)[1][0]
    assert torch.equal(ld[:1000], ld[1000:])


def test_gradient_is_linear_in_the_loss_cotangent_at_scale():
    """grad(sum_n a_n l_n) + grad(sum_n b_n l_n) == grad(sum_n (a_n + b_n) l_n): checks the adjoint's accumulation
    (TMEM accumulator flushes, per-CTA partials, butterflies) at a size with many tiles per CTA."""
    PM, target, pf, unf, fixed, kw = _big_problem(K=32)
    n = 1 << 16
    seeds = torch.from_numpy(seeds_for(n, seed=5))
    g = torch.Generator().manual_seed(0)
    a = (torch.rand(n, generator=g) / n).cuda()
    b = (torch.rand(n, generator=g) / n).cuda()

    def grad(cot):
        p = pf.detach().requires_grad_(True)
        l = PM.compute_log_elbo(seeds, p, unf, fixed, target, **kw)[0]
        fin = torch.isfinite(l.detach())
        (gp,) = torch.autograd.grad(l, p, grad_outputs=torch.where(fin, cot, torch.zeros_like(cot)))
        return gp
    ga, gb, gab = grad(a), grad(b), grad(a + b)
    scale = gab.abs().max().item()
    assert torch.isfinite(gab).all() and scale > 0
    assert ((ga + gb) - gab).abs().max().item() / scale < 2e-5


@pytest.mark.parametrize("n,K", [(1, 8), (127, 3), (129, 1), (257, 1)])
def test_small_and_ragged_batches_match_oracle(n, K):
    from cmcd_b200 import mcdboundingmachine as PM
    c, lp, dim, pf, unf, fixed = oracle_problem("C_manygmm_dds_small", N=n, K=K)
    _, target, _, pf_p, unf_p, fixed_p = product_problem("C_manygmm_dds_small", pf, N=n, K=K)
    kw = dict(eps_schedule=c["eps_schedule"], grad_clipping=c["clip"])
    seeds = seeds_for(n, seed=11)
    g_o, (l_o, z_o) = OM.grad_and_loss(OM.compute_bound, seeds, pf, unf, fixed, lp, **kw)
    g_p, (l_p, z_p) = PM.grad_and_loss(lambda *a: PM.compute_bound(*a, **kw))(torch.from_numpy(seeds), pf_p, unf_p, fixed_p, target)
    fin = torch.isfinite(l_o)
    assert (torch.isfinite(l_p.cpu()) == fin).all()
    assert rel_err(l_p.cpu()[fin], l_o[fin]).max() < 1e-4
    assert rel_err(z_p.cpu()[fin], z_o[fin]).max() < 1e-4
    assert (g_p.cpu() - g_o).abs().max().item() <= 1e-4 * max(g_o.abs().max().item(), 1e-12) + 1e-7


def test_empty_batch_and_bad_arguments():
    from cmcd_b200 import mcdboundingmachine as PM
    PMod, target, pf, unf, fixed, kw = _big_problem(K=4)
    with torch.no_grad():
        l, (lv, z) = PM.compute_bound(torch.zeros(0, dtype=torch.int32), pf, unf, fixed, target, **kw)
    assert lv.numel() == 0 and z.shape == (0, 2)
    with pytest.raises(NotImplementedError):
        PM.compute_bound(torch.arange(1, 5), pf, unf, (fixed[0], fixed[1], "MCD_DNF", fixed[3]), target, **kw)


@pytest.mark.parametrize("name", ["C_manygmm_dds_small", "A_gmm", "B_funnel"])
def test_empty_and_single_particle_batches(name):
    """N = 0 (an empty shard of a sharded step) returns empty outputs without a launch; N = 1 runs a one-particle tile / CTA and
    equals the first particle of a larger batch bit for bit (per-particle results do not depend on the batch)."""
    from cmcd_b200 import mcdboundingmachine as PM
    c, lp, dim, pf, unf, fixed = oracle_problem(name)
    _, target, _, pf_p, unf_p, fixed_p = product_problem(name, pf)
    kw = dict(eps_schedule=c["eps_schedule"], grad_clipping=c["clip"])
    seeds = torch.from_numpy(seeds_for(64))
    with torch.no_grad():
        l0, (z0, _) = PM.compute_log_elbo(seeds[:0], pf_p, unf_p, fixed_p, target, **kw)
        assert l0.shape == (0,) and z0.shape == (0, dim)
        l1, (z1, _) = PM.compute_log_elbo(seeds[:1], pf_p, unf_p, fixed_p, target, **kw)
        lf, (zf, _) = PM.compute_log_elbo(seeds, pf_p, unf_p, fixed_p, target, **kw)
    assert torch.equal(l1, lf[:1]) and torch.equal(z1, zf[:1])
    g1, _ = PM.grad_and_loss(lambda *a: PM.compute_bound(*a, **kw))(seeds[:1], pf_p, unf_p, fixed_p, target)
    assert torch.isfinite(g1).all()
