"""Training-loop pieces (scope table "next" rows): seed draw, fused Adam + project, run / sample mirrors."""
import numpy as np
import pytest
import torch

from helpers import oracle_problem, product_problem, seeds_for
from oracle import mcdboundingmachine as OM
from oracle import opt as OO
from oracle import prng as P


def test_oracle_randint_properties():
    key = P.prng_key(7)
    a = OO.randint(key, 100_001, 1, 10**6)
    assert a.dtype == np.int32 and a.min() >= 1 and a.max() < 10**6
    np.testing.assert_array_equal(a, OO.randint(key, 100_001, 1, 10**6))
    assert abs(a.mean() / 5e5 - 1) < 0.02 and len(np.unique(a)) > 90_000
    # span = 1 (maxval <= minval) always returns minval, like jax (issue 222)
    assert (OO.randint(key, 10, 5, 5) == 5).all()


def test_oracle_randint_uint32_multiplier_wrap():
    """jax/_src/random.py::_randint squares 2^16 % span with lax.mul in uint32: for span > 65536 the product wraps to 0 and
    the draw is minval + lower_bits % span (the reference's randint(key, (N,), 1, 1e6), opt.py:94, is in that regime);
    for span <= 65536 the multiplier is (2^32 mod span) and both words contribute."""
    key = P.prng_key(3)
    k1, k2 = P.split(np.asarray(key, np.uint32))
    hi, lo = P.random_bits(k1, 1001).astype(np.uint64), P.random_bits(k2, 1001).astype(np.uint64)
    np.testing.assert_array_equal(OO.randint(key, 1001, 1, 10**6), (1 + lo % 999999).astype(np.int32))
    span = 1000
    want = ((hi % span) * ((1 << 32) % span) + lo % span) % span
    np.testing.assert_array_equal(OO.randint(key, 1001, 0, span), want.astype(np.int32))


def test_host_key_split_matches_oracle():
    from cmcd_b200 import opt as PO
    k = PO.prng_key(1)
    for _ in range(3):
        a, b = PO.split_key(k)
        ao, bo = P.split(np.asarray(k, np.uint32))
        assert a.tolist() == ao.tolist() and b.tolist() == bo.tolist()
        k = b


@pytest.mark.gpu
def test_randint_matches_oracle():
    """Kernel vs the oracle restatement of _randint (itself unpinned against a running JAX), both multiplier regimes."""
    from cmcd_b200 import opt as PO
    for seed, n in ((1, 300), (2, 2001), (3, 1 << 16)):
        key = PO.prng_key(seed)
        got = PO.randint_seeds(key, n).cpu().numpy()
        np.testing.assert_array_equal(got, OO.randint(key, n, 1, 10**6))
        got = PO.randint_seeds(key, n, 0, 1000).cpu().numpy()
        np.testing.assert_array_equal(got, OO.randint(key, n, 0, 1000))


@pytest.mark.gpu
def test_adam_project_matches_oracle():
    from cmcd_b200 import opt as PO
    g = np.random.default_rng(0)
    n = 5000
    p = g.normal(size=n).astype(np.float32)
    lo = np.where(g.random(n) < 0.1, -0.5, -np.inf).astype(np.float32)
    hi = np.where(g.random(n) < 0.1, 0.5, np.inf).astype(np.float32)
    m, v = np.zeros(n, np.float32), np.zeros(n, np.float32)
    opt = PO.Optimizer(1e-2)
    pd = torch.from_numpy(p.copy()).cuda()
    st = opt.init(pd)
    ema = pd.clone()
    ema_ref = p.copy()
    for step in range(1, 6):
        grad = (g.normal(size=n) * 10).astype(np.float32)    # exercises the +-5 clip
        p, m, v = OO.adam_project_step(p, grad, m, v, step, 1e-2, lo, hi)
        ema_ref = (np.float32(0.001) * p + np.float32(0.999) * ema_ref).astype(np.float32)
        opt.step(pd, torch.from_numpy(grad).cuda(), st, torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda(), ema, 0.001)
    np.testing.assert_allclose(pd.cpu().numpy(), p, rtol=1e-5, atol=1e-6)   # fp32, FMA contraction on the device
    np.testing.assert_allclose(ema.cpu().numpy(), ema_ref, rtol=1e-5, atol=1e-6)
    # divergence guard: with the flag set nothing moves
    before = pd.clone()
    opt.step(pd, torch.full((n,), float("nan")).cuda(), st, skip_flag=torch.ones((), dtype=torch.int32).cuda())
    assert torch.equal(pd, before)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["A_gmm", "LDVI_gmm", "CAISUHA_gmm"])
def test_run_matches_oracle_loop(name):
    """5 iterations of opt.run on the README gmm config (CMCD, LDVI and 2nd-order CMCD operators) vs the same loop driven by the
    oracle (grad + numpy Adam): eps / gamma / mgridref_y projections included."""
    from cmcd_b200 import mcdboundingmachine as PM
    from cmcd_b200 import opt as PO
    from cmcd_b200.pytree import tree_map

    class Info:
        N = 300
        run_cluster = 1
    c, lp, dim, pf, unf, fixed = oracle_problem(name, torch.float32)
    _, target, _, pf_p, unf_p, fixed_p = product_problem(name, pf)
    kw = dict(eps_schedule=c["eps_schedule"], grad_clipping=c["clip"])
    trainable = c["trainable"]
    iters, lr = 5, 1e-3
    gl = PM.grad_and_loss(lambda *a: PM.compute_bound(*a, **kw))
    losses, p_prod, _ = PO.run(Info, lr, iters, pf_p, unf_p, fixed_p, target, gl, trainable, PO.prng_key(1))
    # oracle-driven loop
    idx_train = tree_map(lambda t: t.numpy(), unf(torch.arange(pf.numel(), dtype=torch.float32))[0])
    lo, hi = OO.project_bounds(pf.numel(), idx_train, trainable)
    p = pf.numpy().astype(np.float32).copy()
    m, v = np.zeros_like(p), np.zeros_like(p)
    key = P.prng_key(1)
    ref_losses = []
    for i in range(iters):
        k, key = P.split(key)
        seeds = OO.randint(k, Info.N, 1, 10**6)
        g, (l, _) = OM.grad_and_loss(OM.compute_bound, seeds, torch.from_numpy(p), unf, fixed, lp, **kw)
        ref_losses.append(l.mean().item())
        p, m, v = OO.adam_project_step(p, g.numpy(), m, v, i + 1, lr, lo, hi)
    np.testing.assert_allclose(losses, ref_losses, rtol=1e-4)
    np.testing.assert_allclose(p_prod.cpu().numpy(), p, rtol=1e-4, atol=2e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["A_gmm", "B_funnel", "C_manygmm_dds_small", "LDVI_gmm", "Ckl_manygmm_geffner"])
def test_run_graph_mode_matches_eager(name):
    """opt.run(graph=True) replays one captured iteration (table chain + forward bridge + adjoint + loss mean + flag):
    same parameters and losses as the eager loop after 6 iterations (gradient atomics change the fp32 summation order)."""
    from cmcd_b200 import mcdboundingmachine as PM
    from cmcd_b200 import opt as PO

    class Info:
        pass
    c, lp, dim, pf, unf, fixed = oracle_problem(name, torch.float32)
    Info.N = c["N"]
    _, target, _, pf_p, unf_p, fixed_p = product_problem(name, pf)
    kw = dict(eps_schedule=c["eps_schedule"], grad_clipping=c["clip"])
    gl = PM.grad_and_loss(lambda *a: PM.compute_bound(*a, **kw))
    out = {}
    for mode in (False, True):
        losses, p, ema = PO.run(Info, 1e-3, 6, pf_p, unf_p, fixed_p, target, gl, c["trainable"], PO.prng_key(5), use_ema=True, graph=mode)
        out[mode] = (np.asarray(losses), p.cpu().numpy(), ema.cpu().numpy())
    np.testing.assert_allclose(out[True][0], out[False][0], rtol=1e-5)
    np.testing.assert_allclose(out[True][1], out[False][1], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(out[True][2], out[False][2], rtol=1e-4, atol=1e-6)


@pytest.mark.gpu
def test_sample_and_final_losses():
    from cmcd_b200 import mcdboundingmachine as PM
    from cmcd_b200 import opt as PO
    from cmcd_b200 import utils as PU
    c, lp, dim, pf, unf, fixed = oracle_problem("A_gmm", torch.float32)
    _, target, _, pf_p, unf_p, fixed_p = product_problem("A_gmm", pf)
    kw = dict(eps_schedule=c["eps_schedule"], grad_clipping=c["clip"])
    loss_fn = lambda *a: PM.compute_bound(*a, **kw)
    key = PO.prng_key(3)
    elbos, zs = PO.sample(None, 50, 4, pf_p, unf_p, fixed_p, target, loss_fn, key)
    assert elbos.shape == (4, 50) and zs.shape == (200, dim)
    seeds = OO.randint(key, 200, 1, 10**6)
    with torch.no_grad():
        l_o = OM.compute_bound(seeds, pf, unf, fixed, lp, **kw)[1][0]
    np.testing.assert_allclose(elbos.reshape(-1).cpu().numpy(), l_o.numpy(), rtol=1e-4, atol=1e-4)
    elbo, lnz = PU.log_final_losses(elbos)
    ref = OM.log_final_losses(l_o.reshape(4, 50).double())
    assert abs(elbo - ref["elbo"]) < 1e-3 and abs(lnz - ref["ln_Z"]) < 1e-3
