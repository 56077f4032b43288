"""Wall-clock per training iteration of opt.run (seed draw + table chain + forward bridge + adjoint + fused Adam/project)
on the README configs, eager launches vs one captured CUDA graph per iteration (dev tool).

    python tools/train_iter_time.py [iters] [config ...]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from helpers import CONFIGS
from cmcd_b200 import mcdboundingmachine as M, model_handler as H, opt as O, variationaldist as V

SIZES = {"A_gmm": (300, 8), "B_funnel": (300, 8), "C_manygmm_dds": (2000, 256), "Cvar_manygmm": (2000, 256), "D_lgcp": (20, 8)}
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100
only = sys.argv[2:]
for name, (N, K) in SIZES.items():
    if only and name not in only:
        continue
    c = CONFIGS[name]
    out = H.load_model(c["model"])
    target, dim = out[0], out[1]
    pf, unf, fixed = M.initialize(dim, vdparams=V.initialize(dim, c["sigma"], device="cuda"), nbridges=K, eps=c["eps"],
                                  trainable=c["trainable"], emb_dim=c["emb_dim"], mode=c["mode"], nn_arch=c["nn_arch"])
    kw = dict(eps_schedule=c["eps_schedule"], grad_clipping=c["clip"])
    bound = M.compute_bound_var if c["mode"] == "MCD_CAIS_var_sn" else M.compute_bound
    gl = M.grad_and_loss(lambda *a: bound(*a, **kw))

    class Info:
        pass
    Info.N = N
    row = dict(config=name, N=N, K=K, iters=iters)
    for graph in (False, True):
        O.run(Info, 1e-3, 5, pf, unf, fixed, target, gl, c["trainable"], O.prng_key(1), sync_every=1000, graph=graph)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        O.run(Info, 1e-3, iters, pf, unf, fixed, target, gl, c["trainable"], O.prng_key(1), sync_every=1000, graph=graph)
        torch.cuda.synchronize()
        row["graph_ms_per_iter" if graph else "eager_ms_per_iter"] = round((time.perf_counter() - t0) / iters * 1e3, 3)
    print(json.dumps(row), flush=True)
