"""Device time of one train iteration (forward bridge + adjoint + table chain) and of one sampling pass for every
BASELINE.json / README config at its README size, CUDA events, median of `reps` (dev tool).

    python tools/config_times.py [reps] [config ...] > gpurun_out/config_times.json
"""
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

# README sizes (the parity tests run some of these at reduced N / K)
SIZES = {"A_gmm": (300, 8), "B_funnel": (300, 8), "C_manygmm_dds": (2000, 256), "C_manygmm_dds_16k": (16384, 256), "Cvar_manygmm": (2000, 256),
         "Ckl_manygmm_geffner": (2000, 256), "D_lgcp": (20, 8), "ULAsn_funnel": (300, 8)}


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 11
    only = sys.argv[2:]   # optional: config names to run
    rows = []
    for name, (N, K) in SIZES.items():
        if only and name not in only:
            continue
        c = dict(CONFIGS[name.replace("_16k", "")])
        out = PH.load_model(c["model"], device="cuda")
        target, dim = out[0], out[1]
        pf, unf, fixed = PM.initialize(dim, vdparams=PV.initialize(dim, c["sigma"], device="cuda"), nbridges=K, eps=c["eps"],
                                       trainable=c["trainable"], emb_dim=c["emb_dim"], mode=c["mode"], nn_arch=c["nn_arch"],
                                       device="cuda")
        kw = dict(eps_schedule=c["eps_schedule"], grad_clipping=c["clip"])
        bound = PM.compute_bound_var if c["mode"] == "MCD_CAIS_var_sn" else PM.compute_bound
        gl = PM.grad_and_loss(lambda *a: bound(*a, **kw))
        seeds = torch.from_numpy(seeds_for(N)).cuda()
        t_train = timed(lambda: gl(seeds, pf, unf, fixed, target), reps)
        with torch.no_grad():
            t_fwd = timed(lambda: bound(seeds, pf, unf, fixed, target, **kw), reps)
        rows.append(dict(config=name, model=c["model"], mode=c["mode"], nn_arch=c["nn_arch"], N=N, K=K, dim=dim,
                         train_iter_ms=round(t_train, 3), sampling_ms=round(t_fwd, 3),
                         train_particle_steps_per_s=round(N * K / t_train * 1e3, 1)))
        print(json.dumps(rows[-1]), flush=True)


if __name__ == "__main__":
    main()
