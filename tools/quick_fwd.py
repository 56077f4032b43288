"""Quick forward-kernel timing on the scaling config (many_gmm, dds, K=256) -- dev tool, not the bench."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cmcd_b200 import mcdboundingmachine as M, model_handler as H

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
K = int(sys.argv[2]) if len(sys.argv) > 2 else 256
target, dim, _ = H.load_model("many_gmm")
from cmcd_b200 import variationaldist as V
pf, unf, fixed = M.initialize(dim, vdparams=V.initialize(dim, 60.0), nbridges=K, eps=1.0, trainable=("eta", "gamma", "mgridref_y"),
                              mode="MCD_CAIS_sn", nn_arch="dds")
# live head
pt, pn = unf(pf)
g = torch.Generator().manual_seed(0)
pt["sn"]["out"]["w"].copy_((torch.randn(64, 2, generator=g) * 0.01).cuda())
seeds = torch.from_numpy(np.random.default_rng(0).integers(1, 10**6, N).astype(np.int32)).cuda()
kw = dict(eps_schedule="cos_sq", grad_clipping=True)
with torch.no_grad():
    for _ in range(2):
        loss, (l, z) = M.compute_bound(seeds, pf, unf, fixed, target, **kw)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        loss, (l, z) = M.compute_bound(seeds, pf, unf, fixed, target, **kw)
    b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
print(f"fwd N={N} K={K}: {ms:.2f} ms  {N*K/ms*1e-6:.3f} G particle-steps/s  finite={torch.isfinite(l).float().mean().item():.4f} loss={l[torch.isfinite(l)].mean().item():.4f}")
