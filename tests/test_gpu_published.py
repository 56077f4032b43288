"""The one external pin the reference tree offers: train the README commands through the CUDA path (opt.run / opt.sample /
log_final_losses) and compare the final ELBO / ln Z with the numbers printed in the reference's notebook
(/root/reference/src/notebooks/plotting_rebuttal.ipynb:413-418 funnel, :554-559 and :1007-1012 gmm, :3500-3508 / :6392 lgcp).

These are trained-model results from the authors' wandb sweeps (11 000 Adam iterations, 30 x n_samples evaluation), so they
pin the whole chain -- seeds, bridge, adjoint, optimizer, estimators -- statistically, not per call: network initial values
are drawn from the reference's distributions but not from its bit stream, and the ELBO of a trained model varies from run
to run.  Tolerance: |ELBO - published| <= max(0.1, 4 x published std) and |ln Z - published| <= max(0.15, 3 x published
std); additionally ln Z must sit within 0.3 of the true value 0 of the normalised targets (funnel, gmm).  The lgcp run
(README.md:63: MFVI pretraining 20 000 iterations + 37 500 bridge iterations at d = 1600, ~100 s on a B200) is the unnormalised Cox
process posterior: ELBO 469.48 +- 0.25 and ln Z 491.06 +- 3.5 in the notebook.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["funnel", "gmm_readme", "lgcp"])
def test_trained_elbo_lnz_match_reference_notebook(name):
    from train_published import RUNS
    from cmcd_b200 import experiment as E
    r = RUNS[name]
    res = E.main(E.get_config(**r["cfg"]), log=lambda s: None)
    assert res["diverged"] is None
    pub = r["published"]
    print(f"{name}: ELBO {res['elbo_final']:.4f} (published {pub['elbo']:.4f} +- {pub['elbo_std']:.3f}); "
          f"ln Z {res['final_ln_Z']:.4f} (published {pub['ln_Z']:.4f} +- {pub['ln_Z_std']:.3f})")
    assert abs(res["elbo_final"] - pub["elbo"]) <= max(0.1, 4 * pub["elbo_std"])
    assert abs(res["final_ln_Z"] - pub["ln_Z"]) <= max(0.15, 3 * pub["ln_Z_std"])
    if name != "lgcp":
        assert abs(res["final_ln_Z"]) < 0.3     # true ln Z = 0
    assert res["losses"][-1] < res["losses"][0]


def test_readme_40gmm_run_finds_the_analytic_ln_z():
    """README.md:26 -- the headline configuration of BASELINE.json at its own size (many_gmm, MCD_CAIS_sn, dds net, N = 2000,
    nbridges = 256, eps = 1 cos^2, sigma0 = 60, grad_clipping, lr 1e-3) -- has no trained-model numbers in the reference tree (wandb
    links only), but the 40-GMM is a normalised mixture: the true ln Z is 0.  5 % of the README's 150 000 iterations (26 s on a B200)
    already bring the estimate there: ln Z = -0.08 +- 0.33 over the 30 evaluation batches, ELBO -2.5; the full schedule
    (`tools/train_published.py many_gmm_dds`, 506 s, profiles/r2_published_many_gmm.json) ends at ln Z = +0.02 +- 0.12, ELBO -1.26."""
    from train_published import RUNS
    from cmcd_b200 import experiment as E
    cfg = E.get_config(**RUNS["many_gmm_dds"]["cfg"])
    cfg.iters = 7500
    res = E.main(cfg, log=lambda s: None)
    assert res["diverged"] is None
    print(f"40-GMM dds, 7500 iterations: ELBO {res['elbo_final']:.4f}, ln Z {res['final_ln_Z']:.4f} +- {res['final_ln_Z_std']:.3f} (true 0)")
    assert abs(res["final_ln_Z"]) < 0.3
    assert -3.5 < res["elbo_final"] < 0.0       # a lower bound of ln Z = 0, and far above the untrained bridge
    assert res["losses"][-1] < 3.5
