"""Sampling pass of README config C (many_gmm, dds) at a small particle count -- dev tool for ncu captures of the four-thread kernel.
    python tools/quick_fwd_small.py [N] [K] [reps]"""
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

N = int(sys.argv[1]) if len(sys.argv) > 1 else 18944
K = int(sys.argv[2]) if len(sys.argv) > 2 else 64
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
c = dict(CONFIGS["C_manygmm_dds"])
target, dim = PH.load_model(c["model"], device="cuda")[:2]
pf, unf, fixed = PM.initialize(dim, vdparams=PV.initialize(dim, c["sigma"], device="cuda"), nbridges=K, eps=c["eps"], trainable=c["trainable"],
                               emb_dim=c["emb_dim"], mode=c["mode"], nn_arch=c["nn_arch"], device="cuda")
kw = dict(eps_schedule=c["eps_schedule"], grad_clipping=c["clip"])
seeds = torch.from_numpy(seeds_for(N)).cuda()
with torch.no_grad():
    ms = timed(lambda: PM.compute_log_elbo(seeds, pf, unf, fixed, target, **kw), reps)
print(f"fwd N={N} K={K}: {ms:.3f} ms, {ms * 1e3 / (K + 1):.2f} us per node")
