"""Where does the full-size (N=2000, K=256) gradient difference of a config come from?  Splits the particles into chunks, differentiates
the same masked sum through kernel / fp32 oracle / fp64 oracle, and prints per-chunk errors (dev tool, GPU).
    python tools/diag_fullsize.py Ckl_manygmm_geffner [chunk]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from cmcd_b200 import mcdboundingmachine as PM
from oracle import mcdboundingmachine as OM
from helpers import oracle_problem, product_problem, seeds_for

name = sys.argv[1] if len(sys.argv) > 1 else "Ckl_manygmm_geffner"
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 100
N, K = 2000, 256
c, lp32, dim, pf, unf, fixed = oracle_problem(name, torch.float32, N=N, K=K)
_, lp64, _, pf64, unf64, fixed64 = oracle_problem(name, torch.float64, N=N, K=K)
_, target, _, pf_p, unf_p, fixed_p = product_problem(name, pf, N=N, K=K)
seeds = seeds_for(N)
kw = dict(eps_schedule=c["eps_schedule"], grad_clipping=c["clip"])
p32, p64, pp = pf.clone().requires_grad_(True), pf64.clone().requires_grad_(True), pf_p.clone().requires_grad_(True)
l32, z32 = OM.compute_log_elbo(seeds, p32, unf, fixed, lp32, c["eps_schedule"], c["clip"])
l64, z64 = OM.compute_log_elbo(seeds, p64, unf64, fixed64, lp64, c["eps_schedule"], c["clip"])
lP, (zP, _) = PM.compute_log_elbo(torch.from_numpy(seeds), pp, unf_p, fixed_p, target, **kw)
fin = torch.isfinite(l64.detach()) & torch.isfinite(l32.detach()) & torch.isfinite(lP.detach().cpu())
rel = lambda l: (l.detach().cpu().double() - l64.detach()).abs() / l64.detach().abs().clamp(min=1)
print("loss rel err: kernel max over finite", rel(lP)[fin].max().item(), "fp32", rel(l32)[fin].max().item())


def run(mask):
    g32 = torch.autograd.grad(l32[mask].sum() / N, p32, retain_graph=True)[0]
    g64 = torch.autograd.grad(l64[mask].sum() / N, p64, retain_graph=True)[0]
    gP = torch.autograd.grad(lP[mask.to(lP.device)].sum() / N, pp, retain_graph=True)[0].cpu()
    return g32.double(), g64, gP.double()


G32, G64, GP = run(fin)
scale = G64.abs().max().item()
print(f"all finite: |g64|max {scale:.3e}; kernel err {(GP - G64).abs().max().item() / scale:.3e}; fp32 err {(G32 - G64).abs().max().item() / scale:.3e}")
rows = []
for a in range(0, N, chunk):
    m = torch.zeros(N, dtype=torch.bool)
    m[a:a + chunk] = True
    m &= fin
    g32, g64, gP = run(m)
    ek, eo = (gP - g64).abs().max().item() / scale, (g32 - g64).abs().max().item() / scale
    rows.append((a, ek, eo, rel(lP)[m].max().item(), rel(l32)[m].max().item()))
    print(f"chunk {a:5d}: kernel {ek:.3e}  fp32 {eo:.3e}   loss err kernel {rows[-1][3]:.2e} fp32 {rows[-1][4]:.2e}", flush=True)
worst = max(rows, key=lambda r: r[1])
print("worst chunk", worst)
a = worst[0]
for n in range(a, a + chunk):
    if not fin[n]:
        continue
    m = torch.zeros(N, dtype=torch.bool)
    m[n] = True
    g32, g64, gP = run(m)
    ek, eo = (gP - g64).abs().max().item() / scale, (g32 - g64).abs().max().item() / scale
    if ek > 1e-4 or eo > 1e-4:
        print(f"  particle {n}: kernel {ek:.3e} fp32 {eo:.3e} |g64_n| {g64.abs().max().item() / scale:.3e} loss err kernel {rel(lP)[n].item():.2e} fp32 {rel(l32)[n].item():.2e} "
              f"z64 {z64[n].tolist()}")
