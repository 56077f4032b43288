"""Workload for ncu captures of the momentum-augmented kernels (dev tool): one train iteration each of LDVI (block path, N = 2000;
one-thread path, N = 2^16) and UHA on the 40-GMM, K = 64."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from helpers import CONFIGS, UHA_CONFIGS, seeds_for
from cmcd_b200 import boundingmachine as PB
from cmcd_b200 import mcdboundingmachine as PM
from cmcd_b200 import model_handler as PH
from cmcd_b200 import variationaldist as PV

K = 64
c = dict(CONFIGS["LDVI_manygmm_dds"])
target, dim = PH.load_model(c["model"], device="cuda")[:2]
pf, unf, fixed = PM.initialize(dim, vdparams=PV.initialize(dim, c["sigma"], device="cuda"), nbridges=K, eps=c["eps"], gamma=c["gamma"],
                               trainable=c["trainable"], emb_dim=c["emb_dim"], mode=c["mode"], nn_arch=c["nn_arch"], device="cuda")
gl = PM.grad_and_loss(PM.compute_bound)
for N in (2000, 1 << 16):
    seeds = torch.from_numpy(seeds_for(N)).cuda()
    for _ in range(2):
        gl(seeds, pf, unf, fixed, target)
u = dict(UHA_CONFIGS["UHA_manygmm_lf2"])
pf, unf, fixed = PB.initialize(dim, vdparams=PV.initialize(dim, u["sigma"], device="cuda"), nbridges=K, lfsteps=u["lfsteps"], eps=u["eps"],
                               eta=u["eta"], trainable=("eps", "eta", "vd", "md", "mgridref_y"), device="cuda")
gl = PM.grad_and_loss(PB.compute_bound)
seeds = torch.from_numpy(seeds_for(1 << 16)).cuda()
for _ in range(2):
    gl(seeds, pf, unf, fixed, target)
torch.cuda.synchronize()
print("done")
