"""Forward bridge at hidden_pad 64, d = 2 with few particles: four threads per particle (csrc/bridge_fwd_tcw.cu, HT = 64) against one
thread per particle, three CTAs per SM (csrc/bridge_fwd_tc.cu): device time of one sampling pass and one train iteration with
CMCD_TC_QUAD unset / =0 (read at call time).  Dev tool.   python tools/tc_quad_crossover.py > gpurun_out/tc_quad_crossover.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from helpers import CONFIGS, seeds_for
from cmcd_b200 import mcdboundingmachine as PM
from cmcd_b200 import model_handler as PH
from cmcd_b200 import variationaldist as PV
from config_times import timed

RUNS = [("C_manygmm_dds", 300, 8), ("C_manygmm_dds", 2000, 256), ("C_manygmm_dds", 8192, 256), ("C_manygmm_dds", 16384, 256), ("C_manygmm_dds", 18944, 256),
        ("ULAsn_gmm_dds", 300, 8)]
for name, N, K in RUNS:
    c = dict(CONFIGS[name])
    target, dim = PH.load_model(c["model"], device="cuda")[:2]
    pf, unf, fixed = PM.initialize(dim, vdparams=PV.initialize(dim, c["sigma"], device="cuda"), nbridges=K, eps=c["eps"],
                                   trainable=c["trainable"], emb_dim=c["emb_dim"], mode=c["mode"], nn_arch=c["nn_arch"], device="cuda")
    kw = dict(eps_schedule=c["eps_schedule"], grad_clipping=c["clip"])
    gl = PM.grad_and_loss(lambda *a: PM.compute_bound(*a, **kw))
    seeds = torch.from_numpy(seeds_for(N)).cuda()
    row = dict(config=name, N=N, K=K)
    for tag, env in (("quad", None), ("one_thread", "0")):
        os.environ.pop("CMCD_TC_QUAD", None)
        if env is not None:
            os.environ["CMCD_TC_QUAD"] = env
        with torch.no_grad():
            row[tag + "_sampling_ms"] = round(timed(lambda: PM.compute_log_elbo(seeds, pf, unf, fixed, target, **kw), 7), 3)
        row[tag + "_train_ms"] = round(timed(lambda: gl(seeds, pf, unf, fixed, target), 7), 3)
    os.environ.pop("CMCD_TC_QUAD", None)
    print(json.dumps(row), flush=True)
