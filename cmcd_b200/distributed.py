"""Particle-sharded train / eval step: one process per GPU, particles split across ranks.

The reference runs on one device only (``jax.vmap`` over particles, mcdboundingmachine.py:193); particles are
independent given the parameters and each particle's PRNG stream depends only on its own seed, so sharding
``seeds`` changes no per-particle result.  Collectives per train iteration:
  * one all-reduce(SUM) of [sum l, sum l^2, n] (needed *before* the backward pass only by the log-variance loss,
    whose cotangent is 2 (l_n - mean l) / N);
  * one all-reduce(SUM) of the flat parameter gradient (<= a few 100 KB except lgcp).
ln Z needs a max + sum-exp combine, done as all-reduce(MAX) then all-reduce(SUM).
Works with any torch.distributed backend (nccl on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def shard_bounds(n, world, rank):
    """Contiguous N/G slices (first ``n % world`` ranks get one extra particle)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def combine_stats(local_sum, local_sumsq, local_n, device, group=None):
    """-> (global sum l, sum l^2, N) as python floats/ints."""
    t = torch.tensor([float(local_sum), float(local_sumsq), float(local_n)], dtype=torch.float64, device=device)
    if _world(group)[0] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t[0].item(), t[1].item(), int(round(t[2].item()))


def sharded_grad_and_loss(local_forward, seeds_local, params_flat, loss="kl", group=None):
    """One train iteration on this rank's particle shard.

    ``local_forward(seeds_local, p)`` -> (l_local[N_r] differentiable w.r.t. ``p``, z_local).
    Returns (grad_flat [all-reduced], loss_value, (l_local, z_local)) where loss_value is the global
    mean (``kl``: compute_bound, mcdboundingmachine.py:205) or the clipped global variance (``var``:
    compute_bound_var, :231)."""
    p = params_flat.detach().requires_grad_(True)
    with torch.enable_grad():
        l, z = local_forward(seeds_local, p)
    ld = l.detach()
    s1, s2, n = combine_stats(ld.double().sum().item() if ld.numel() else 0.0,
                              (ld.double() ** 2).sum().item() if ld.numel() else 0.0, ld.numel(), p.device, group)
    mean = s1 / n
    if loss == "kl":
        value = mean
        cot = torch.full_like(ld, 1.0 / n)
    elif loss == "var":
        var = s2 / n - mean * mean
        value = min(max(var, -1e7), 1e7)
        inside = -1e7 <= var <= 1e7  # jnp.clip passes gradient only inside the interval
        cot = (2.0 / n) * (ld - mean) if inside else torch.zeros_like(ld)
    else:
        raise ValueError(loss)
    if l.requires_grad and l.numel():
        (g,) = torch.autograd.grad(l, p, grad_outputs=cot.to(l.dtype), allow_unused=True)
        g = torch.zeros_like(p) if g is None else g
    else:
        g = torch.zeros_like(p)
    if _world(group)[0] > 1:
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
    return g, value, (ld, z.detach())


def global_ln_z(neg_l_local, n_global=None, group=None):
    """ln Z estimate logsumexp(-l) - log N over all ranks (utils.py:231-237) from local losses l."""
    x = -neg_l_local.detach().double() if False else -neg_l_local.detach().double()
    m = x.max() if x.numel() else torch.tensor(-math.inf, dtype=torch.float64, device=x.device)
    world = _world(group)[0]
    if world > 1:
        dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    t = torch.stack([torch.exp(x - m).sum() if x.numel() else torch.zeros((), dtype=torch.float64, device=x.device),
                     torch.tensor(float(x.numel()), dtype=torch.float64, device=x.device)])
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    n = t[1].item() if n_global is None else n_global
    return (torch.log(t[0]) + m - math.log(n)).item()
