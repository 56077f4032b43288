"""Forward bridge of the README geffner net (hidden_pad 136) on 144-wide tcgen05 tiles (csrc/bridge_fwd_tc.cu, HT = 144) against
the block-cooperative FP32 mapping (csrc/bridge_blk.cu): device time of one sampling pass and of one train iteration with
CMCD_TC_WIDE unset / =0 (read at call time).  Dev tool.
    python tools/tc_wide_crossover.py > gpurun_out/tc_wide_crossover.json"""
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

RUNS = [("Ckl_manygmm_geffner", 300, 16), ("Ckl_manygmm_geffner", 2000, 256), ("Cvar_manygmm", 2000, 256), ("Ckl_manygmm_geffner", 8192, 64),
        ("Ckl_manygmm_geffner", 18944, 64), ("Ckl_manygmm_geffner", 32768, 64), ("Ckl_manygmm_geffner", 131072, 64)]

for name, N, K in RUNS:
    c = dict(CONFIGS[name])
    out = PH.load_model(c["model"], device="cuda")
    target, dim = out[0], out[1]
    pf, unf, fixed = PM.initialize(dim, vdparams=PV.initialize(dim, c["sigma"], device="cuda"), nbridges=K, eps=c["eps"],
                                   gamma=c.get("gamma", 10.0), trainable=c["trainable"], emb_dim=c["emb_dim"], mode=c["mode"],
                                   nn_arch=c["nn_arch"], device="cuda")
    kw = dict(eps_schedule=c["eps_schedule"], grad_clipping=c["clip"])
    bound = PM.compute_bound_var if c["mode"] == "MCD_CAIS_var_sn" else PM.compute_bound
    gl = PM.grad_and_loss(lambda *a: bound(*a, **kw))
    seeds = torch.from_numpy(seeds_for(N)).cuda()
    row = dict(config=name, N=N, K=K, hidden_pad=fixed[3].hidden_pad)
    for tag, env in (("tc144", None), ("fp32", "0")):
        os.environ.pop("CMCD_TC_WIDE", None)
        if env is not None:
            os.environ["CMCD_TC_WIDE"] = env
        with torch.no_grad():
            row[tag + "_sampling_ms"] = round(timed(lambda: PM.compute_log_elbo(seeds, pf, unf, fixed, target, **kw), 5), 3)
        row[tag + "_train_ms"] = round(timed(lambda: gl(seeds, pf, unf, fixed, target), 5), 3)
    os.environ.pop("CMCD_TC_WIDE", None)
    print(json.dumps(row), flush=True)
