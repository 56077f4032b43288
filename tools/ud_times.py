"""Device time of one train iteration / one sampling pass of the momentum-augmented operators (SURVEY.md section 8f row 3:
LDVI family through csrc/bridge_ud.cu, UHA through csrc/bridge_uha.cu) at README-like sizes and at a throughput size,
CUDA events, median of `reps`; next to it the CPU oracle restatement timed on this box's host cores on a bounded sample.

    python tools/ud_times.py [reps] > gpurun_out/ud_times.json
Algorithmic flops per particle-step (2 per MAC; activations / RNG excluded, SURVEY 8d convention): one network evaluation
2 (in H + H^2 + H d) plus two target scores per step forward; x3 for a train iteration (input-VJP, weight gradient, HVPs).
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from helpers import CONFIGS, UHA_CONFIGS, oracle_problem, seeds_for, uha_oracle_problem
from cmcd_b200 import boundingmachine as PB
from cmcd_b200 import mcdboundingmachine as PM
from cmcd_b200 import model_handler as PH
from cmcd_b200 import variationaldist as PV
from oracle import mcdboundingmachine as OM

SCORE_FLOPS = {"gmm": 150, "many_gmm": 600, "funnel": 40}
# (config, N, K): README-like small runs and a throughput size
RUNS = [("LDVI_gmm", 300, 8), ("LDVI_funnel_dds", 300, 8), ("LDVI_manygmm_dds", 2000, 256), ("LDVI_manygmm_dds", 1 << 17, 64),
        ("UDea_gmm", 300, 8), ("UDesna_funnel_dds", 300, 8), ("UD_gmm", 1 << 17, 64)]
UHA_RUNS = [("UHA_gmm", 300, 8), ("UHA_funnel_lf3", 300, 8), ("UHA_manygmm_lf2", 2000, 256), ("UHA_manygmm_lf2", 1 << 17, 64)]


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


def cpu_time(fn, budget_s=6.0):
    fn()
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        fn()
        n += 1
    return (time.perf_counter() - t0) / n


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 11
    torch.set_num_threads(os.cpu_count())
    for name, N, K in RUNS:
        c = dict(CONFIGS[name])
        out = PH.load_model(c["model"], device="cuda")
        target, dim = out[0], out[1]
        pf, unf, fixed = PM.initialize(dim, vdparams=PV.initialize(dim, c["sigma"], device="cuda"), nbridges=K, eps=c["eps"],
                                       gamma=c.get("gamma", 10.0), eta=c.get("eta", 0.5), trainable=c["trainable"],
                                       emb_dim=c["emb_dim"], mode=c["mode"], nn_arch=c["nn_arch"], device="cuda")
        gl = PM.grad_and_loss(PM.compute_bound)
        seeds = torch.from_numpy(seeds_for(N)).cuda()
        t_train = timed(lambda: gl(seeds, pf, unf, fixed, target), reps)
        with torch.no_grad():
            t_fwd = timed(lambda: PM.compute_bound(seeds, pf, unf, fixed, target), reps)
        af = fixed[3]
        net = 2 * (af.in_dim * af.hidden + af.hidden ** 2 + af.hidden * dim) if af is not None else 0
        flop_fwd = net + 2 * SCORE_FLOPS[c["model"]]
        row = dict(config=name, model=c["model"], mode=c["mode"], nn_arch=c["nn_arch"] if af is not None else None, N=N, K=K, dim=dim,
                   train_iter_ms=round(t_train, 3), sampling_ms=round(t_fwd, 3),
                   train_particle_steps_per_s=round(N * K / t_train * 1e3, 1), sampling_particle_steps_per_s=round(N * K / t_fwd * 1e3, 1),
                   algorithmic_flops_per_particle_step_train=3 * flop_fwd,
                   train_algorithmic_tflops=round(3 * flop_fwd * N * K / t_train * 1e-9, 3))
        if N <= 2000:   # CPU oracle restatement on the same config, bounded sample (fp32, all host cores)
            cN, cK = min(N, 300), min(K, 16)
            _, lp, _, pfo, unfo, fixedo = oracle_problem(name, torch.float32, N=cN, K=cK)
            so = seeds_for(cN)
            t_cpu = cpu_time(lambda: OM.grad_and_loss(OM.compute_bound, so, pfo, unfo, fixedo, lp))
            row.update(cpu_oracle_particle_steps_per_s=round(cN * cK / t_cpu, 1), cpu_cores=os.cpu_count(), cpu_sample=f"N={cN} K={cK} train iterations for 6 s")
        print(json.dumps(row), flush=True)
    for name, N, K in UHA_RUNS:
        c = dict(UHA_CONFIGS[name])
        out = PH.load_model(c["model"], device="cuda")
        target, dim = out[0], out[1]
        pf, unf, fixed = PB.initialize(dim, vdparams=PV.initialize(dim, c["sigma"], device="cuda"), nbridges=K, lfsteps=c["lfsteps"],
                                       eps=c["eps"], eta=c["eta"], trainable=("eps", "eta", "vd", "md", "mgridref_y"), device="cuda")
        gl = PM.grad_and_loss(PB.compute_bound)
        seeds = torch.from_numpy(seeds_for(N)).cuda()
        t_train = timed(lambda: gl(seeds, pf, unf, fixed, target), reps)
        with torch.no_grad():
            t_fwd = timed(lambda: PB.compute_bound(seeds, pf, unf, fixed, target), reps)
        flop_fwd = (c["lfsteps"] + 1) * SCORE_FLOPS[c["model"]]
        row = dict(config=name, model=c["model"], mode="UHA", lfsteps=c["lfsteps"], N=N, K=K, dim=dim,
                   train_iter_ms=round(t_train, 3), sampling_ms=round(t_fwd, 3),
                   train_particle_steps_per_s=round(N * K / t_train * 1e3, 1), sampling_particle_steps_per_s=round(N * K / t_fwd * 1e3, 1),
                   algorithmic_flops_per_particle_step_train=3 * flop_fwd,
                   train_algorithmic_tflops=round(3 * flop_fwd * N * K / t_train * 1e-9, 3))
        if N <= 2000:
            cN, cK = min(N, 300), min(K, 16)
            _, lp, _, pfo, unfo, fixedo = uha_oracle_problem(name, torch.float32, N=cN, K=cK)
            so = seeds_for(cN)
            t_cpu = cpu_time(lambda: OM.grad_and_loss(OM.uha_compute_bound, so, pfo, unfo, fixedo, lp))
            row.update(cpu_oracle_particle_steps_per_s=round(cN * cK / t_cpu, 1), cpu_cores=os.cpu_count(), cpu_sample=f"N={cN} K={cK} train iterations for 6 s")
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
