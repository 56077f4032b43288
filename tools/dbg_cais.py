"""Conditioning probe for the MCD_CAIS_UHA_sn parity config (dev tool): kernel and fp32-oracle gradient error against the fp64
oracle over K, N and step size."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import helpers
from helpers import oracle_problem, product_problem, seeds_for
from oracle import mcdboundingmachine as OM
from cmcd_b200 import mcdboundingmachine as PM
from cmcd_b200.pytree import tree_leaves

def leaf_errs(g, ref, unf):
    out = []
    for a, b in zip(tree_leaves(unf(g)), tree_leaves(unf(ref))):
        a, b = a.double().reshape(-1), b.double().reshape(-1)
        if b.numel() == 0: continue
        sc = b.abs().max().item()
        out.append((a - b).abs().max().item() / sc if sc > 0 else (a - b).abs().max().item())
    return np.array(out)

def run(name, K=None, **over):
    helpers.CONFIGS[name + "_x"] = dict(helpers.CONFIGS[name], **over)
    nm = name + "_x"
    c, lp, dim, pf, unf, fixed = oracle_problem(nm, torch.float32, K=K)
    _, lp64, _, pf64, unf64, fixed64 = oracle_problem(nm, torch.float64, K=K)
    seeds = seeds_for(c["N"])
    g32, (l32, _) = OM.grad_and_loss(OM.compute_bound, seeds, pf, unf, fixed, lp)
    g64, (l64, _) = OM.grad_and_loss(OM.compute_bound, seeds, pf64, unf64, fixed64, lp64)
    _, target, _, pf_p, unf_p, fixed_p = product_problem(nm, pf, K=K)
    gp, (lp_, zp_) = PM.grad_and_loss(PM.compute_bound)(torch.from_numpy(seeds), pf_p, unf_p, fixed_p, target)
    ek, eo = leaf_errs(gp.cpu(), g64, unf), leaf_errs(g32, g64, unf)
    el = ((lp_.cpu().double() - l64).abs() / l64.abs().clamp(min=1)).max().item()
    el32 = ((l32.double() - l64).abs() / l64.abs().clamp(min=1)).max().item()
    print("   leaf errs", np.array2string(ek, precision=1, max_line_width=250))
    print(f"{name} K={K} {over}: loss err {el:.2e} (fp32 oracle {el32:.2e}); grad kernel max {ek.max():.2e} oracle32 max {eo.max():.2e}", flush=True)

# eps = 0.3 (the first version of the parity config): the leapfrog is unstable between mixture modes and one particle's cotangent
# recursion amplifies fp32 rounding in BOTH fp32 implementations; eps = 0.1 (the committed config) is well conditioned
for K in (12, 16):
    run("CAISUHA_manygmm_dds", K=K, N=64, eps=0.3)
for N in (16, 32):
    run("CAISUHA_manygmm_dds", K=16, N=N, eps=0.3)
run("CAISUHA_manygmm_dds", K=16, N=64)
