"""Shared test plumbing: matched (oracle, product) problem instances on identical parameters."""
from __future__ import annotations

import numpy as np
import torch

from oracle import mcdboundingmachine as OM
from oracle import model_handler as OH

# name -> settings.  README configs A-E of SURVEY.md appendix A (sizes reduced where noted).
CONFIGS = {
    "A_gmm": dict(model="gmm", mode="MCD_CAIS_sn", N=300, K=8, nn_arch="geffner", emb_dim=20, eps=0.01, sigma=1.0,
                  eps_schedule=None, clip=False, trainable=("eta", "gamma", "vd", "mgridref_y")),
    "B_funnel": dict(model="funnel", mode="MCD_CAIS_sn", N=300, K=8, nn_arch="geffner", emb_dim=48, eps=0.1, sigma=1.0,
                     eps_schedule="cos_sq", clip=False, trainable=("eta", "gamma", "vd", "mgridref_y")),
    "C_manygmm_dds": dict(model="many_gmm", mode="MCD_CAIS_sn", N=2000, K=256, nn_arch="dds", emb_dim=20, eps=1.0,
                          sigma=60.0, eps_schedule="cos_sq", clip=True, trainable=("eta", "gamma", "mgridref_y")),
    "C_manygmm_dds_small": dict(model="many_gmm", mode="MCD_CAIS_sn", N=500, K=16, nn_arch="dds", emb_dim=20, eps=0.3,
                                sigma=20.0, eps_schedule="cos_sq", clip=True, trainable=("eta", "gamma", "mgridref_y")),
    "Cvar_manygmm": dict(model="many_gmm", mode="MCD_CAIS_var_sn", N=500, K=16, nn_arch="geffner", emb_dim=130, eps=0.65,
                         sigma=15.0, eps_schedule=None, clip=True, trainable=("eta", "gamma", "mgridref_y")),
    "Ckl_manygmm_geffner": dict(model="many_gmm", mode="MCD_CAIS_sn", N=300, K=16, nn_arch="geffner", emb_dim=130, eps=0.1,
                                sigma=15.0, eps_schedule=None, clip=True, trainable=("eta", "gamma", "mgridref_y")),
    # hidden_pad 144 (emb_dim 142 + d = 2): the 144-wide tensor-core tiles without padding columns; MCD_ULA_sn node form
    "ULAsn_manygmm_e142": dict(model="many_gmm", mode="MCD_ULA_sn", N=300, K=12, nn_arch="geffner", emb_dim=142, eps=0.1,
                               sigma=15.0, eps_schedule="cos_sq", clip=True, trainable=("eta", "gamma", "eps", "mgridref_y")),
    "ULA_gmm": dict(model="gmm", mode="MCD_ULA", N=300, K=8, nn_arch="geffner", emb_dim=20, eps=0.01, sigma=1.0,
                    eps_schedule=None, clip=False, trainable=("eta", "gamma", "eps", "vd", "mgridref_y")),
    "ULAsn_funnel": dict(model="funnel", mode="MCD_ULA_sn", N=300, K=8, nn_arch="geffner", emb_dim=20, eps=0.05, sigma=1.0,
                         eps_schedule="cos_sq", clip=True, trainable=("eta", "gamma", "eps", "vd", "mgridref_y")),
    "ULAsn_gmm_dds": dict(model="gmm", mode="MCD_ULA_sn", N=300, K=8, nn_arch="dds", emb_dim=20, eps=0.02, sigma=1.5,
                          eps_schedule=None, clip=False, trainable=("eta", "gamma", "eps", "vd", "mgridref_y")),
    "D_lgcp": dict(model="lgcp", mode="MCD_CAIS_sn", N=20, K=8, nn_arch="geffner", emb_dim=20, eps=1e-3, sigma=0.3,
                   vd_mean=3.5, eps_schedule=None, clip=False, trainable=("eta", "gamma", "eps", "vd", "mgridref_y")),
    "D_lgcp_ula": dict(model="lgcp", mode="MCD_ULA", N=11, K=4, nn_arch="geffner", emb_dim=20, eps=5e-4, sigma=0.3,
                       vd_mean=3.5, eps_schedule=None, clip=False, trainable=("eta", "gamma", "eps", "vd", "mgridref_y")),
    # config.use_whitened = True (model_handler.py:348-351,373-384): density of the whitened variable, served by the callback path
    "D_lgcp_white": dict(model="lgcp", model_cfg=dict(use_whitened=True), mode="MCD_CAIS_sn", N=11, K=4, nn_arch="geffner", emb_dim=20,
                         eps=1e-3, sigma=0.3, eps_schedule=None, clip=False, trainable=("eta", "gamma", "eps", "vd", "mgridref_y")),
    "lin_funnel": dict(model="funnel", mode="MCD_CAIS_sn", N=200, K=12, nn_arch="dds", emb_dim=20, eps=0.05, sigma=1.0,
                       eps_schedule="linear", clip=True, trainable=("eta", "gamma", "eps", "vd", "mgridref_y")),
    # underdamped "LDVI" family (mcd_under_lp_a.py): network on (z, rho), on z only, and none
    "LDVI_gmm": dict(model="gmm", mode="MCD_U_a-lp-sn", N=300, K=8, nn_arch="geffner", emb_dim=20, eps=0.05, sigma=1.0,
                     eps_schedule=None, clip=False, trainable=("eta", "gamma", "eps", "vd", "mgridref_y")),
    "LDVI_funnel_dds": dict(model="funnel", mode="MCD_U_a-lp-sn", N=300, K=8, nn_arch="dds", emb_dim=20, eps=0.04, sigma=1.0,
                            gamma=6.0, eps_schedule=None, clip=False, trainable=("eta", "gamma", "eps", "vd", "mgridref_y")),
    "LDVI_manygmm_dds": dict(model="many_gmm", mode="MCD_U_a-lp-sn", N=300, K=16, nn_arch="dds", emb_dim=20, eps=0.2, sigma=15.0,
                             gamma=2.0, eps_schedule=None, clip=False, trainable=("eta", "gamma", "eps", "mgridref_y")),
    "UDsna_funnel": dict(model="funnel", mode="MCD_U_a-lp-sna", N=300, K=8, nn_arch="geffner", emb_dim=20, eps=0.04, sigma=1.0,
                         eps_schedule=None, clip=False, trainable=("eta", "gamma", "eps", "vd", "mgridref_y")),
    "UDe_gmm": dict(model="gmm", mode="MCD_U_e-lp", N=300, K=8, nn_arch="geffner", emb_dim=20, eps=0.05, sigma=1.0, eta=0.6,
                    eps_schedule=None, clip=False, trainable=("eta", "gamma", "eps", "vd", "mgridref_y")),
    "UDesna_funnel_dds": dict(model="funnel", mode="MCD_U_e-lp-sna", N=300, K=8, nn_arch="dds", emb_dim=20, eps=0.04, sigma=1.0,
                              eta=0.5, eps_schedule=None, clip=False, trainable=("eta", "gamma", "eps", "vd", "mgridref_y")),
    "UDea_gmm": dict(model="gmm", mode="MCD_U_ea-lp-sn", N=300, K=8, nn_arch="geffner", emb_dim=20, eps=0.05, sigma=1.0,
                     gamma=5.0, eps_schedule=None, clip=False, trainable=("eta", "gamma", "eps", "vd", "mgridref_y")),
    "CAISUHA_gmm": dict(model="gmm", mode="MCD_CAIS_UHA_sn", N=300, K=8, nn_arch="geffner", emb_dim=20, eps=0.08, sigma=1.0,
                        gamma=4.0, eps_schedule=None, clip=False, trainable=("eta", "gamma", "eps", "vd", "mgridref_y")),
    # (eps = 0.3 makes the leapfrog unstable between mixture modes: one particle's cotangent recursion then amplifies fp32
    #  rounding to 1e-3 in BOTH fp32 implementations -- tools/dbg_cais.py; 0.1 keeps the chain well conditioned)
    "CAISUHA_manygmm_dds": dict(model="many_gmm", mode="MCD_CAIS_UHA_sn", N=300, K=16, nn_arch="dds", emb_dim=20, eps=0.1, sigma=15.0,
                                gamma=2.0, eps_schedule=None, clip=False, trainable=("eta", "gamma", "eps", "mgridref_y")),
    "UD_gmm": dict(model="gmm", mode="MCD_U_a-lp", N=300, K=8, nn_arch="geffner", emb_dim=20, eps=0.05, sigma=1.0,
                   eps_schedule=None, clip=False, trainable=("eta", "gamma", "eps", "vd", "mgridref_y")),
}


def seeds_for(n, seed=0):
    """seeds = default_rng(0).integers(1, 10**6, N, int32) (mirrors opt.py:94)."""
    return np.random.default_rng(seed).integers(1, 10**6, n).astype(np.int32)


def oracle_problem(name, dtype=torch.float32, N=None, K=None):
    c = dict(CONFIGS[name])
    if N:
        c["N"] = N
    if K:
        c["K"] = K
    log_prob, dim = OH.load_model(c["model"], OH.default_config(**c.get("model_cfg", {})), dtype=dtype)
    # parameters are always drawn in float32 (identical values for every dtype), then cast
    vdp = OM.vd_initialize(dim, c["sigma"])
    g = torch.Generator().manual_seed(7)
    vdp["mean"] = vdp["mean"] + 0.1 * torch.randn(dim, generator=g) + c.get("vd_mean", 0.0)  # non-trivial mean
    mgrid = 1.0 + 0.3 * torch.rand(min(32, c["K"]) + 1, generator=g)
    pf, unf, fixed = OM.initialize(dim, vdparams=vdp, nbridges=c["K"], eps=c["eps"], gamma=c.get("gamma", 10.0), eta=c.get("eta", 0.5), trainable=c["trainable"],
                                   emb_dim=c["emb_dim"], mode=c["mode"], nn_arch=c["nn_arch"], mgridref_y=mgrid,
                                   live=True)
    return c, log_prob, dim, pf.to(dtype), unf, fixed


def product_problem(name, pf_oracle, device="cuda", N=None, K=None):
    """Same pytree layout as the oracle (identical initialize arguments) -> the flat vectors are interchangeable."""
    from cmcd_b200 import mcdboundingmachine as PM
    from cmcd_b200 import model_handler as PH
    from cmcd_b200 import variationaldist as PV
    c = dict(CONFIGS[name])
    if N:
        c["N"] = N
    if K:
        c["K"] = K
    from types import SimpleNamespace
    out = PH.load_model(c["model"], config=SimpleNamespace(**c["model_cfg"]) if "model_cfg" in c else None, device=device)
    target, dim = out[0], out[1]
    mgrid = torch.ones(min(32, c["K"]) + 1)
    pf, unf, fixed = PM.initialize(dim, vdparams=PV.initialize(dim, device=device), nbridges=c["K"], eps=c["eps"], gamma=c.get("gamma", 10.0), eta=c.get("eta", 0.5),
                                   trainable=c["trainable"], emb_dim=c["emb_dim"], mode=c["mode"],
                                   nn_arch=c["nn_arch"], mgridref_y=mgrid, device=device)
    assert pf.numel() == pf_oracle.numel(), (pf.numel(), pf_oracle.numel())
    return c, target, dim, pf_oracle.to(torch.float32).to(device), unf, fixed


def rel_err(a, b, floor=1.0):
    """|a-b| / max(|b|, floor) elementwise (numpy)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)


# ---------------------------------------------------------------- UHA (boundingmachine.py + ais_utils.py, boundmode "UHA")
UHA_CONFIGS = {
    "UHA_gmm": dict(model="gmm", N=300, K=8, lfsteps=1, eps=0.1, eta=0.5, sigma=1.0),
    "UHA_funnel_lf3": dict(model="funnel", N=300, K=6, lfsteps=3, eps=0.05, eta=0.7, sigma=1.0),
    "UHA_manygmm_lf2": dict(model="many_gmm", N=300, K=16, lfsteps=2, eps=0.3, eta=0.3, sigma=15.0),
}
UHA_TRAINABLE = ("eps", "eta", "vd", "md", "mgridref_y")


def uha_oracle_problem(name, dtype=torch.float32, N=None, K=None):
    c = dict(UHA_CONFIGS[name])
    if N:
        c["N"] = N
    if K:
        c["K"] = K
    log_prob, dim = OH.load_model(c["model"], dtype=dtype)
    g = torch.Generator().manual_seed(11)
    vdp = OM.vd_initialize(dim, c["sigma"])
    vdp["mean"] = vdp["mean"] + 0.1 * torch.randn(dim, generator=g)
    md = 0.2 * torch.randn(dim, generator=g)                       # non-trivial momentum scales
    mgrid = 1.0 + 0.3 * torch.rand(min(32, c["K"]) + 1, generator=g)
    pf, unf, fixed = OM.uha_initialize(dim, vdparams=vdp, nbridges=c["K"], lfsteps=c["lfsteps"], eps=c["eps"], eta=c["eta"],
                                       mdparams=md, mgridref_y=mgrid, trainable=UHA_TRAINABLE)
    return c, log_prob, dim, pf.to(dtype), unf, fixed


def uha_product_problem(name, pf_oracle, device="cuda", N=None, K=None):
    from cmcd_b200 import boundingmachine as PB
    from cmcd_b200 import model_handler as PH
    c = dict(UHA_CONFIGS[name])
    if N:
        c["N"] = N
    if K:
        c["K"] = K
    out = PH.load_model(c["model"], device=device)
    target, dim = out[0], out[1]
    pf, unf, fixed = PB.initialize(dim, nbridges=c["K"], lfsteps=c["lfsteps"], eps=c["eps"], eta=c["eta"],
                                   mgridref_y=torch.ones(min(32, c["K"]) + 1), trainable=UHA_TRAINABLE, device=device)
    assert pf.numel() == pf_oracle.numel(), (pf.numel(), pf_oracle.numel())
    return c, target, dim, pf_oracle.to(torch.float32).to(device), unf, fixed
