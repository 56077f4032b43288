"""Where the block-cooperative mapping (csrc/bridge_blk.cu) beats one thread per particle: device time of one train iteration
for several (config, N) with CMCD_BLK_ALWAYS=1 and CMCD_DISABLE_BLK=1 (both read at call time).  Dev tool.
    python tools/blk_crossover.py > gpurun_out/blk_crossover.json"""
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

RUNS = [("Cvar_manygmm", 8192, 64), ("Cvar_manygmm", 32768, 64), ("A_gmm", 8192, 64), ("A_gmm", 131072, 64), ("B_funnel", 8192, 64),
        ("B_funnel", 65536, 64), ("lin_funnel", 8192, 64), ("lin_funnel", 65536, 64), ("LDVI_gmm", 65536, 64), ("LDVI_funnel_dds", 32768, 64)]

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
    for tag, env in (("block", "CMCD_BLK_ALWAYS"), ("one_thread", "CMCD_DISABLE_BLK")):
        os.environ.pop("CMCD_BLK_ALWAYS", None)
        os.environ.pop("CMCD_DISABLE_BLK", None)
        os.environ[env] = "1"
        row[tag + "_train_ms"] = round(timed(lambda: gl(seeds, pf, unf, fixed, target), 5), 3)
    print(json.dumps(row), flush=True)
