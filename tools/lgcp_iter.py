"""A few train iterations of the README lgcp config (d = 1600, N = 20, K = 8) -- workload for launch lists / ncu captures (dev tool).
    python tools/lgcp_iter.py [reps]"""
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

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
c = dict(CONFIGS["D_lgcp"])
N, K = 20, 8
target, dim = PH.load_model(c["model"], device="cuda")[:2]
pf, unf, fixed = PM.initialize(dim, vdparams=PV.initialize(dim, c["sigma"], device="cuda"), nbridges=K, eps=c["eps"], trainable=c["trainable"],
                               emb_dim=c["emb_dim"], mode=c["mode"], nn_arch=c["nn_arch"], device="cuda")
kw = dict(eps_schedule=c["eps_schedule"], grad_clipping=c["clip"])
gl = PM.grad_and_loss(lambda *a: PM.compute_bound(*a, **kw))
seeds = torch.from_numpy(seeds_for(N)).cuda()
for _ in range(reps):
    gl(seeds, pf, unf, fixed, target)
torch.cuda.synchronize()
torch.cuda.profiler.start()
gl(seeds, pf, unf, fixed, target)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
