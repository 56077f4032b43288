"""Analytic pins for the oracle restatement: normalised targets, CAIS == ULA at
zero drift, K=0 == MFVI bound, autograd scores == closed forms, estimators."""
import math

import numpy as np
import pytest
import torch

from oracle import mcdboundingmachine as OM
from oracle import model_handler as OH
from helpers import oracle_problem, seeds_for


def test_targets_are_normalised_2d():
    # ln Z = 0: integrate exp(log p) on a grid (fp64)
    for model, lim, n in (("gmm", 12.0, 1201), ("many_gmm", 48.0, 2401)):
        lp, dim = OH.load_model(model, dtype=torch.float64)
        g = torch.linspace(-lim, lim, n, dtype=torch.float64)
        xx, yy = torch.meshgrid(g, g, indexing="ij")
        p = torch.exp(lp(torch.stack([xx.reshape(-1), yy.reshape(-1)], -1)))
        z = p.sum().item() * (2 * lim / (n - 1)) ** 2
        assert abs(z - 1.0) < 2e-3, (model, z)


def test_funnel_closed_form():
    lp, dim = OH.load_model("funnel", dtype=torch.float64)
    x = torch.randn(64, dim, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    v = x[:, 0]
    ref = (-v ** 2 / 18 - math.log(3 * math.sqrt(2 * math.pi)) - 0.5 * torch.exp(-v) * (x[:, 1:] ** 2).sum(-1)
           - (dim - 1) * (v + math.log(2 * math.pi)) / 2)
    torch.testing.assert_close(lp(x), ref, rtol=1e-12, atol=1e-12)


def test_lgcp_constants_golden():
    c = OH.lgcp_constants(OH.default_config().file_path)
    assert c["counts"].sum() == 127 and c["counts"].max() == 4 and (c["counts"] > 0).sum() == 106
    assert abs(c["half_log_det"] - 225.70548) < 1e-3
    assert abs(c["mu_zero"] - 3.8812819) < 1e-6
    ev = np.linalg.eigvalsh(c["gram"])
    assert abs(ev.max() / ev.min() - 27.6) < 0.2


def test_cais_equals_ula_at_zero_drift():
    # reference init: factor_sn = 0 -> network output exactly 0 -> CAIS step == ULA step
    lp, dim = OH.load_model("gmm")
    seeds = seeds_for(64)
    out = {}
    for mode in ("MCD_CAIS_sn", "MCD_ULA"):
        pf, unf, fixed = OM.initialize(dim, nbridges=6, eps=0.02, trainable=("vd",), emb_dim=8, mode=mode,
                                       nn_arch="geffner", live=False)
        out[mode] = OM.compute_bound(seeds, pf, unf, fixed, lp)[1][0]
    torch.testing.assert_close(out["MCD_CAIS_sn"], out["MCD_ULA"], rtol=0, atol=0)


def test_k0_ula_equals_mfvi():
    lp, dim = OH.load_model("funnel")
    seeds = seeds_for(50)
    pf, unf, fixed = OM.initialize(dim, nbridges=0, trainable=("vd",), mode="MCD_ULA")
    a = OM.compute_bound(seeds, pf, unf, fixed, lp)[1][0]
    pf2, unf2, fixed2 = OM.bm_initialize(dim)
    b = OM.bm_compute_bound(seeds, pf2, unf2, fixed2, lp)[1][0]
    torch.testing.assert_close(a, b, rtol=0, atol=0)


@pytest.mark.parametrize("name", ["A_gmm", "Cvar_manygmm", "ULAsn_funnel"])
def test_fp32_vs_fp64_oracle(name):
    c, lp, dim, pf, unf, fixed = oracle_problem(name, torch.float32, N=100)
    c64, lp64, _, pf64, unf64, fixed64 = oracle_problem(name, torch.float64, N=100)
    seeds = seeds_for(100)
    kw = dict(eps_schedule=c["eps_schedule"], grad_clipping=c["clip"])
    l32 = OM.compute_bound(seeds, pf, unf, fixed, lp, **kw)[1][0]
    l64 = OM.compute_bound(seeds, pf64, unf64, fixed64, lp64, **kw)[1][0]
    fin = torch.isfinite(l64)
    assert (torch.isfinite(l32) == fin).all()
    assert ((l32[fin].double() - l64[fin]).abs() / l64[fin].abs().clamp(min=1)).median() < 1e-5


def test_grad_matches_finite_difference_fp64():
    c, lp, dim, pf, unf, fixed = oracle_problem("A_gmm", torch.float64, N=20)
    seeds = seeds_for(20)
    g, _ = OM.grad_and_loss(OM.compute_bound, seeds, pf, unf, fixed, lp)
    rng = np.random.default_rng(0)
    for idx in rng.choice(np.nonzero(g.numpy())[0], 6, replace=False):
        e = torch.zeros_like(pf)
        e[idx] = 1e-6
        fd = (OM.compute_bound(seeds, pf + e, unf, fixed, lp)[0] - OM.compute_bound(seeds, pf - e, unf, fixed, lp)[0]) / 2e-6
        assert abs(fd.item() - g[idx].item()) < 1e-6 * max(1, abs(g[idx].item())), (idx, fd.item(), g[idx].item())


def test_log_final_losses():
    e = torch.randn(30, 500, dtype=torch.float64)
    r = OM.log_final_losses(e)
    assert abs(r["elbo"] + e.mean().item()) < 1e-12
    lnz = np.log(np.exp(-e.numpy()).mean(1))
    assert abs(r["ln_Z"] - lnz.mean()) < 1e-10 and abs(r["ln_Z_std"] - lnz.std()) < 1e-10


# ---------------------------------------------------------------- underdamped "LDVI" family (mcd_under_lp_a.py)
def test_ldvi_zero_drift_equals_no_network():
    # reference init: factor_sn = 0 -> the (z, rho) network contributes exactly 0 -> MCD_U_a-lp-sn == MCD_U_a-lp
    lp, dim = OH.load_model("gmm")
    seeds = seeds_for(64)
    out = {}
    for mode in ("MCD_U_a-lp-sn", "MCD_U_a-lp-sna", "MCD_U_a-lp"):
        pf, unf, fixed = OM.initialize(dim, nbridges=6, eps=0.05, gamma=4.0, trainable=("vd",), emb_dim=8, mode=mode,
                                       nn_arch="geffner", live=False)
        out[mode] = OM.compute_bound(seeds, pf, unf, fixed, lp)[1][0]
    torch.testing.assert_close(out["MCD_U_a-lp-sn"], out["MCD_U_a-lp"], rtol=0, atol=0)
    torch.testing.assert_close(out["MCD_U_a-lp-sna"], out["MCD_U_a-lp"], rtol=0, atol=0)


def test_underdamped_variants_agree_where_the_reference_formulas_coincide():
    # lp_ea differs from lp_a only in the forward kernel (exact OU refresh instead of its first-order expansion):
    # as gamma*eps -> 0 both refreshes coincide to O((gamma eps)^2) -- the two restatements must converge at that rate
    lp, dim = OH.load_model("gmm", dtype=torch.float64)
    seeds = seeds_for(200)
    gaps = []
    for gamma in (0.4, 0.2):
        out = {}
        for mode in ("MCD_U_a-lp-sn", "MCD_U_ea-lp-sn"):
            pf, unf, fixed = OM.initialize(dim, nbridges=4, eps=0.05, gamma=gamma, trainable=("vd",), emb_dim=8, mode=mode,
                                           nn_arch="geffner", live=True, dtype=torch.float64)
            out[mode] = OM.compute_bound(seeds, pf, unf, fixed, lp)[1][1]   # z_K depends on the refresh only
        gaps.append((out["MCD_U_a-lp-sn"] - out["MCD_U_ea-lp-sn"]).abs().max().item())
    assert gaps[1] < gaps[0] / 2.5 and gaps[0] < 1e-2, gaps


@pytest.mark.parametrize("mode", ["MCD_U_a-lp", "MCD_U_a-lp-sn", "MCD_U_e-lp-sna", "MCD_U_ea-lp-sn", "MCD_CAIS_UHA_sn"])
def test_ldvi_weights_are_unbiased(mode):
    """E[exp(w)] = Z = 1 for ANY parameters: the momentum refresh is a proper Markov kernel, the leapfrog step a
    volume-preserving bijection, the backward kernel a normalised density (Geffner & Domke 2021, eq. 9).  Pins the
    operator restatement (signs of the log-ratio, which momentum enters which kernel) without a running reference."""
    lp, dim = OH.load_model("gmm", dtype=torch.float64)
    n = 40000
    pf, unf, fixed = OM.initialize(dim, vdparams=OM.vd_initialize(dim, 2.0, torch.float64), nbridges=4, eps=0.2, gamma=3.0, eta=0.4,
                                   trainable=("vd",), emb_dim=8, mode=mode, nn_arch="geffner", live=True, dtype=torch.float64)
    with torch.no_grad():
        l = OM.compute_bound(np.arange(1, n + 1, dtype=np.int32), pf, unf, fixed, lp)[1][0]
    wts = torch.exp(-l)
    se = wts.std().item() / math.sqrt(n)
    assert abs(wts.mean().item() - 1.0) < 4 * se + 1e-3, (wts.mean().item(), se)


def test_ldvi_grad_matches_finite_difference_fp64():
    c, lp, dim, pf, unf, fixed = oracle_problem("LDVI_gmm", torch.float64, N=20)
    seeds = seeds_for(20)
    g, _ = OM.grad_and_loss(OM.compute_bound, seeds, pf, unf, fixed, lp)
    rng = np.random.default_rng(1)
    for idx in rng.choice(np.nonzero(g.numpy())[0], 6, replace=False):
        e = torch.zeros_like(pf)
        e[idx] = 1e-6
        fd = (OM.compute_bound(seeds, pf + e, unf, fixed, lp)[0] - OM.compute_bound(seeds, pf - e, unf, fixed, lp)[0]) / 2e-6
        assert abs(fd.item() - g[idx].item()) < 1e-6 * max(1, abs(g[idx].item())), (idx, fd.item(), g[idx].item())


# ---------------------------------------------------------------- UHA (boundingmachine.py + ais_utils.py)
@pytest.mark.parametrize("lfsteps", [1, 3])
def test_uha_weights_are_unbiased(lfsteps):
    """E[exp(w)] = 1 for any eps / eta / md: the refresh leaves N(0, sigma_m^2) invariant-in-law per step and the leapfrog
    is a volume-preserving bijection (Geffner & Domke 2021, UHA).  Pins the restatement of ais_utils.evolve."""
    lp, dim = OH.load_model("gmm", dtype=torch.float64)
    n = 40000
    pf, unf, fixed = OM.uha_initialize(dim, vdparams=OM.vd_initialize(dim, 2.0, torch.float64), nbridges=4, lfsteps=lfsteps,
                                       eps=0.2, eta=0.4, mdparams=torch.tensor([0.2, -0.1], dtype=torch.float64),
                                       trainable=("vd",), dtype=torch.float64)
    with torch.no_grad():
        l = OM.uha_compute_bound(np.arange(1, n + 1, dtype=np.int32), pf, unf, fixed, lp)[1][0]
    wts = torch.exp(-l)
    se = wts.std().item() / math.sqrt(n)
    assert abs(wts.mean().item() - 1.0) < 4 * se + 1e-3, (wts.mean().item(), se)


def test_uha_grad_matches_finite_difference_fp64():
    lp, dim = OH.load_model("funnel", dtype=torch.float64)
    pf, unf, fixed = OM.uha_initialize(dim, nbridges=3, lfsteps=2, eps=0.05, eta=0.4, mdparams=0.1 * torch.ones(dim, dtype=torch.float64),
                                       trainable=("eps", "eta", "vd", "md", "mgridref_y"), dtype=torch.float64)
    seeds = seeds_for(16)
    g, _ = OM.grad_and_loss(OM.uha_compute_bound, seeds, pf, unf, fixed, lp)
    rng = np.random.default_rng(2)
    for idx in rng.choice(np.nonzero(g.numpy())[0], 8, replace=False):
        e = torch.zeros_like(pf)
        e[idx] = 1e-6
        fd = (OM.uha_compute_bound(seeds, pf + e, unf, fixed, lp)[0] - OM.uha_compute_bound(seeds, pf - e, unf, fixed, lp)[0]) / 2e-6
        assert abs(fd.item() - g[idx].item()) < 2e-6 * max(1, abs(g[idx].item())), (idx, fd.item(), g[idx].item())


def test_uha_k0_equals_mfvi():
    lp, dim = OH.load_model("funnel")
    seeds = seeds_for(50)
    pf, unf, fixed = OM.uha_initialize(dim, nbridges=0, trainable=("vd",))
    a = OM.uha_compute_bound(seeds, pf, unf, fixed, lp)[1][0]
    pf2, unf2, fixed2 = OM.bm_initialize(dim)
    b = OM.bm_compute_bound(seeds, pf2, unf2, fixed2, lp)[1][0]
    torch.testing.assert_close(a, b, rtol=0, atol=0)


def test_analytic_scores_match_autograd():
    """bench.py's CPU-baseline leg may time the oracle with closed-form q / many_gmm scores (oracle.analytic_scores) instead of the
    create_graph autograd scores: same losses and gradients up to rounding."""
    from helpers import oracle_problem, seeds_for
    c, lp, dim, pf, unf, fixed = oracle_problem("C_manygmm_dds_small", torch.float64, N=64, K=6)
    seeds = seeds_for(64)
    kw = dict(eps_schedule=c["eps_schedule"], grad_clipping=c["clip"])
    g0, (l0, z0) = OM.grad_and_loss(OM.compute_bound, seeds, pf, unf, fixed, lp, **kw)
    with OM.analytic_scores():
        g1, (l1, z1) = OM.grad_and_loss(OM.compute_bound, seeds, pf, unf, fixed, lp, **kw)
    fin = torch.isfinite(l0)
    assert (torch.isfinite(l1) == fin).all()
    torch.testing.assert_close(l1[fin], l0[fin], rtol=1e-10, atol=1e-10)
    torch.testing.assert_close(g1, g0, rtol=1e-8, atol=1e-10)
