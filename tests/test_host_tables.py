"""Host-side tables of the momentum-augmented operators (CPU): `mcd_utils.ud_coeff_table` must reproduce, row by row, the
scalars the reference step bodies form in-line (mcd_under_lp_a.py:28-51, mcd_under_lp_e.py:27-43, mcd_under_lp_ea.py:28-57,
mcd_under_lp_a_cais.py:33-58,79-82), and chain gradients into eps / gamma / eta."""
import math

import pytest
import torch

from cmcd_b200 import mcd_utils


def _params(eps=0.07, gamma=3.0, eta=0.6):
    return {k: torch.tensor(v, dtype=torch.float32, requires_grad=True) for k, v in (("eps", eps), ("gamma", gamma), ("eta", eta))}


@pytest.mark.parametrize("mode", ["MCD_U_a-lp", "MCD_U_a-lp-sna", "MCD_U_a-lp-sn", "MCD_U_e-lp", "MCD_U_e-lp-sna", "MCD_U_ea-lp-sn",
                                  "MCD_CAIS_UHA_sn"])
def test_ud_coeff_rows_match_reference_formulas(mode):
    K = 5
    p = _params()
    rows = mcd_utils.ud_coeff_table(mode, p, K)
    assert rows.shape == (7, K)
    eps, gamma, eta = (p[k].detach() for k in ("eps", "gamma", "eta"))
    for i in range(K):
        e = eps
        if mode == "MCD_CAIS_UHA_sn":   # cosine schedule, mcd_under_lp_a_cais.py:33-40
            e = eps * torch.cos(torch.tensor((i / K + 0.008) / 1.008 * 0.5 * math.pi)) ** 2
        eta_aux = gamma * e
        if mode.startswith("MCD_U_a") or mode == "MCD_CAIS_UHA_sn":
            want = (e, 1.0 - eta_aux, torch.sqrt(2.0 * eta_aux), 1.0 - eta_aux, 2 * eta_aux, torch.sqrt(2.0 * eta_aux),
                    -2.0 * eta_aux if mode == "MCD_CAIS_UHA_sn" else torch.tensor(0.0))
        elif mode.startswith("MCD_U_e-"):
            want = (e, eta, torch.sqrt(1.0 - eta ** 2), eta, 2 * (1.0 - eta), torch.sqrt(1.0 - eta ** 2), torch.tensor(0.0))
        else:
            a_f = torch.exp(-gamma * e)
            want = (e, a_f, torch.sqrt(1.0 - a_f ** 2), 1.0 - eta_aux, 2 * eta_aux, torch.sqrt(2.0 * eta_aux), torch.tensor(0.0))
        torch.testing.assert_close(rows[:, i].detach(), torch.stack([torch.as_tensor(w, dtype=torch.float32) for w in want]),
                                   rtol=1e-6, atol=1e-7)
    # cotangents reach the scalars each operator depends on (and only those)
    g = torch.autograd.grad(rows.sum(), [p["eps"], p["gamma"], p["eta"]], allow_unused=True)
    uses_eta = mode.startswith("MCD_U_e-")
    assert (g[2] is not None and g[2].abs() > 0) == uses_eta
    assert (g[1] is not None and g[1].abs() > 0) == (not uses_eta)
    assert g[0] is not None and g[0].abs() > 0


def test_clip_settings_follow_the_operator():
    inf = float("inf")
    assert mcd_utils._clips("MCD_CAIS_UHA_sn", False) == (1e2, inf)      # stable=True is hard-coded, mcd_under_lp_a_cais.py:48
    assert mcd_utils._clips("MCD_U_a-lp-sn", True) == (inf, inf)        # the lp_a / lp_e / lp_ea operators never clip
    assert mcd_utils._clips("MCD_CAIS_sn", True) == (1e3, inf) and mcd_utils._clips("MCD_CAIS_var_sn", True) == (1e2, 1e2)


def test_mfvi_machine_pytree_matches_reference_layout():
    """boundingmachine.initialize(nbridges=0) (main.py:83-85) builds the same pytree as the reference: vd | then the frozen leaves in
    sorted key order eps, eta, gridref_x (2), md (dim), mgridref_y (1), target_x (0) -- boundingmachine.py:9-70."""
    import torch
    from cmcd_b200 import boundingmachine as PB
    from oracle import mcdboundingmachine as OM
    dim = 3
    pf, unf, fixed = PB.initialize(dim, trainable=("vd",), init_sigma=1.5, device="cpu")
    po, unfo, fixedo = OM.bm_initialize(dim, init_sigma=1.5)
    assert fixed == fixedo == (dim, 0, 1)
    assert pf.numel() == po.numel() == 2 * dim + 1 + 1 + 2 + dim + 1 + 0
    assert torch.equal(pf, po)
    pt, pn = unf(pf)
    assert sorted(pt) == ["vd"] and sorted(pn) == ["eps", "eta", "gridref_x", "md", "mgridref_y", "target_x"]
    assert pn["target_x"].numel() == 0 and pn["mgridref_y"].tolist() == [1.0] and pn["gridref_x"].tolist() == [0.0, 1.0]


def test_whitened_lgcp_closed_form_matches_autograd():
    """config.use_whitened = True (model_handler.py:348-351,373-384): the product serves the whitened LGCP density through the callback
    path with a closed-form score / Hessian-vector product; both must equal autograd over the oracle's restatement of
    whitened_posterior_log_density (fp64)."""
    import torch
    from cmcd_b200 import model_handler as PH
    from oracle import model_handler as OH
    cfg = OH.default_config(use_whitened=True)
    lp, d = OH.load_model("lgcp", cfg, dtype=torch.float64)
    target, dim = PH.load_model("lgcp", cfg, device="cpu")
    assert dim == d == 1600 and target.kind == "callback"
    g = torch.Generator().manual_seed(3)
    x = torch.randn(4, d, generator=g, dtype=torch.float64) * 0.3
    v = torch.randn(4, d, generator=g, dtype=torch.float64)
    xr = x.clone().requires_grad_(True)
    l = lp(xr)
    (s,) = torch.autograd.grad(l.sum(), xr, create_graph=True)
    (h,) = torch.autograd.grad((s * v).sum(), xr)
    l2, s2, h2 = target.closed_form(x.float(), v.float())
    rel = lambda a, b: ((a.double() - b).abs().max() / b.abs().max()).item()
    assert rel(l2, l.detach()) < 1e-6 and rel(s2, s.detach()) < 1e-5 and rel(h2, h) < 1e-5
    # the unwhitened density at latent = L e + mu0 differs from the whitened one at e by the log-determinant only
    lp_u, _ = OH.load_model("lgcp", OH.default_config(), dtype=torch.float64)
    c = OH.lgcp_constants(cfg.file_path)
    latent = x @ torch.tensor(c["chol"]).T + c["mu_zero"]
    assert torch.allclose(lp_u(latent) + c["half_log_det"], l.detach(), rtol=1e-10, atol=1e-8)
