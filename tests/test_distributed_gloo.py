"""world_size-2 gloo test of the particle-sharded step: sharding + all-reduce reproduce the single-process result.

The local per-shard computation is injected (here: the CPU oracle) so the host-side combine logic is covered without
a GPU; the CUDA path plugs the kernel-backed compute_log_elbo into the same function (bench.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _problem():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import oracle_problem, seeds_for
    c, lp, dim, pf, unf, fixed = oracle_problem("A_gmm", torch.float64, N=48)
    return c, lp, pf, unf, fixed, seeds_for(48)


def _local_forward(c, lp, unf, fixed):
    from oracle import mcdboundingmachine as OM

    def f(seeds, p):
        if len(seeds) == 0:
            return p.new_zeros(0), p.new_zeros(0, fixed[0])
        l, z = OM.compute_log_elbo(seeds, p, unf, fixed, lp, c["eps_schedule"], c["clip"])
        return l, z
    return f


def _worker(rank, world, port, loss, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cmcd_b200.distributed import ShardedStep, global_ln_z, shard_bounds, sharded_grad_and_loss
    c, lp, pf, unf, fixed, seeds = _problem()
    lo, hi = shard_bounds(len(seeds), world, rank)
    if loss == "kl_step":     # the step object bench.py uses (eager form: CUDA graphs need a GPU), one fused all-reduce
        class _Fwd:           # seeds arrive as the step's static int32 tensor
            def __call__(self, s, p):
                return _local_forward(c, lp, unf, fixed)(s.numpy(), p)
        step = ShardedStep(_Fwd(), pf.float(), hi - lo, graph=False)
        g, value, (l, z) = step(torch.from_numpy(seeds[lo:hi]))
        assert step.n_global == len(seeds)
    else:
        g, value, (l, z) = sharded_grad_and_loss(_local_forward(c, lp, unf, fixed), seeds[lo:hi], pf, loss=loss)
    lnz = global_ln_z(l)
    if rank == 0:
        torch.save({"g": g, "value": value, "lnz": lnz}, out)
    dist.destroy_process_group()


@pytest.mark.parametrize("loss", ["kl", "var", "kl_step"])
def test_two_rank_step_matches_single_process(tmp_path, loss):
    from oracle import mcdboundingmachine as OM
    c, lp, pf, unf, fixed, seeds = _problem()
    fn = OM.compute_bound_var if loss == "var" else OM.compute_bound
    g_ref, (l_ref, _) = OM.grad_and_loss(fn, seeds, pf, unf, fixed, lp, eps_schedule=c["eps_schedule"], grad_clipping=c["clip"])
    v_ref = l_ref.var(unbiased=False).item() if loss == "var" else l_ref.mean().item()
    out = str(tmp_path / "r0.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, loss, out), nprocs=2, join=True)
    r = torch.load(out)
    if loss == "kl_step":    # the step works in float32 (the product's dtype)
        assert abs(float(r["value"]) - v_ref) < 1e-5 * max(1, abs(v_ref))
        torch.testing.assert_close(r["g"].double(), g_ref, rtol=1e-4, atol=1e-6)
        return
    assert abs(r["value"] - v_ref) < 1e-9 * max(1, abs(v_ref))
    torch.testing.assert_close(r["g"], g_ref, rtol=1e-9, atol=1e-12)
    lnz_ref = (torch.logsumexp(-l_ref, 0) - np.log(len(seeds))).item()
    assert abs(r["lnz"] - lnz_ref) < 1e-9


def test_shard_bounds_cover():
    from cmcd_b200.distributed import shard_bounds
    for n in (0, 1, 7, 48, 1 << 20):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
