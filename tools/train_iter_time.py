"""Wall-clock per training iteration of opt.run on a README config (dev tool): python tools/train_iter_time.py [N] [iters]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cmcd_b200 import mcdboundingmachine as M, model_handler as H, opt as O, variationaldist as V

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
target, dim, _ = H.load_model("many_gmm")
trainable = ("eta", "gamma", "mgridref_y")
pf, unf, fixed = M.initialize(dim, vdparams=V.initialize(dim, 60.0), nbridges=256, eps=1.0, trainable=trainable, mode="MCD_CAIS_sn", nn_arch="dds")
kw = dict(eps_schedule="cos_sq", grad_clipping=True)
gl = M.grad_and_loss(lambda *a: M.compute_bound(*a, **kw))

class Info:
    pass
Info.N = N
O.run(Info, 1e-3, 5, pf, unf, fixed, target, gl, trainable, O.prng_key(1), sync_every=1000)
torch.cuda.synchronize()
for sync_every in (1, 1000):
    t0 = time.perf_counter()
    O.run(Info, 1e-3, iters, pf, unf, fixed, target, gl, trainable, O.prng_key(1), sync_every=sync_every)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / iters
    print(f"N={N} K=256 sync_every={sync_every}: {dt*1e3:.2f} ms / iteration  ({N*256/dt/1e6:.1f} M particle-steps/s)")
