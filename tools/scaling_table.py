"""Assemble profiles/r2_scaling.md from the bench.py --sweep lines of the 1/2/4/8-GPU runs (gpurun_out/bench_r2_n{1,2,4,8}.json)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
runs = {}
for n in (1, 2, 4, 8):
    p = os.path.join(ROOT, "gpurun_out", f"bench_{tag}_n{n}.json")
    if os.path.exists(p):
        runs[n] = json.load(open(p))
        json.dump(runs[n], open(os.path.join(ROOT, "profiles", f"{tag}_bench_sweep_n{n}.json"), "w"))
out = [f"# Strong scaling of one CMCD train iteration ({tag}; B200, NVLink/NVSwitch, one process per GPU, NCCL)", "",
       "Workload: many_gmm (40-GMM, d = 2), MCD_CAIS_sn, nn_arch = dds, nbridges = 256, eps = 1 cos^2, sigma0 = 60, grad_clipping;",
       "`N_global` particles sharded over the ranks, per-rank step (fused O(K) chain + forward bridge + adjoint + loss statistics) captured",
       "in a CUDA graph, ONE fused all-reduce of [gradient | sum l | sum l^2 | n] per iteration, no host synchronisation.  Times are",
       "CUDA events, max over ranks, 5 timed iterations after 3 warm-ups (`bench.py --gpus N --sweep`).", "",
       "## Headline: N_global = 2^20 (BASELINE.json configs[4])", "",
       "| GPUs | ms / iteration | particle-steps/s | speed-up | efficiency | end to end (host seeds in, gradient out) | weak scaling (2^20 per GPU) |", "|---|---|---|---|---|---|---|"]
base = runs[1]["ms_per_step"] if 1 in runs else None
for n, d in sorted(runs.items()):
    sp = base / d["ms_per_step"] if base else float("nan")
    weak = d.get("weak_scaling")
    out.append(f"| {n} | {d['ms_per_step']:.2f} | {d['value']:.4g} | {sp:.2f}x | {sp / n:.3f} | {d['e2e']['value']:.4g} ({d['e2e']['ms_per_step']:.2f} ms) | "
               + (f"{weak['value']:.4g} ({weak['ms_per_step']:.1f} ms, {weak['value'] / (n * runs[1]['value']):.3f})" if weak and 1 in runs else "-") + " |")
out += ["", "## Sweep N_global = 2^16 .. 2^20: ms per iteration (efficiency vs 1 GPU at the same N_global)", "",
        "| N_global | " + " | ".join(f"{n} GPU" + ("s" if n > 1 else "") for n in sorted(runs)) + " |", "|---|" + "---|" * len(runs)]
sw = {n: {r["particles_global"]: r for r in d.get("sweep", [])} for n, d in runs.items()}
for e in range(16, 21):
    ng = 1 << e
    row = [f"2^{e}"]
    for n in sorted(runs):
        r = sw[n].get(ng)
        if not r:
            row.append("-")
            continue
        eff = sw[1][ng]["ms_per_step"] / (n * r["ms_per_step"]) if 1 in sw and ng in sw[1] else float("nan")
        row.append(f"{r['ms_per_step']:.2f}" + (f" ({eff:.2f})" if n > 1 else ""))
    out.append("| " + " | ".join(row) + " |")
out += ["", "## Where a rank's iteration goes (ms): forward kernel / adjoint kernel / all-reduce / rest (chain kernels, statistics, graph gaps)", "",
        "| N_global | " + " | ".join(f"{n} GPU" + ("s" if n > 1 else "") for n in sorted(runs)) + " |", "|---|" + "---|" * len(runs)]
for e in range(16, 21):
    ng = 1 << e
    row = [f"2^{e}"]
    for n in sorted(runs):
        r = sw[n].get(ng)
        row.append("-" if not r else f"{r['fwd_kernel_ms']:.2f} / {r['bwd_kernel_ms']:.2f} / {r['allreduce_ms']:.3f} / {r['rest_ms']:.2f}")
    out.append("| " + " | ".join(row) + " |")
out += ["", "Reading: the collective is 0.02-0.03 ms at every size (84 KB over NVLink) and the host chain 0.1-0.5 ms, so scaling is set by the two",
        "bridge kernels alone.  They process whole 128-particle tiles -- forward 3 CTAs per SM, adjoint 2 tiles per SM -- and one trajectory",
        "point of one tile is a ~5 us (forward) / ~10 us (adjoint) dependency chain, so a pass cannot take less than ~257 x that: ~1.3 ms +",
        "~2.5 ms.  At N_global = 2^20 on 8 GPUs a rank still has 1024 tiles (7 per SM) and the loss is tile quantisation (the adjoint's 3.46",
        "rounds cost 4); from 2^17 per 8 GPUs downwards (<= 128 tiles per rank, fewer tiles than SMs) the time is the latency floor and adding GPUs",
        "no longer helps -- the regime where more particles per GPU are free."]
open(os.path.join(ROOT, "profiles", f"{tag}_scaling.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
