"""Particle-sharded train / eval step: one process per GPU, particles split across ranks.

The reference runs on one device only (``jax.vmap`` over particles, mcdboundingmachine.py:193); particles are
independent given the parameters and each particle's PRNG stream depends only on its own seed, so sharding
``seeds`` changes no per-particle result.  Collectives per train iteration:
  * KL loss (compute_bound, mcdboundingmachine.py:205): ONE all-reduce(SUM) of the fused buffer
    ``[flat gradient | sum l | sum l^2 | n]`` -- the cotangent 1/N_global does not depend on the losses, so nothing
    has to be exchanged between the forward and the backward pass;
  * log-variance loss (compute_bound_var, :231): one extra all-reduce(SUM) of ``[sum l, n]`` between forward and
    backward (the cotangent is 2 (l_n - mean l) / N), then the same fused buffer carrying ``sum (l - mean)^2``.
ln Z needs a max + sum-exp combine, done as all-reduce(MAX) then all-reduce(SUM) on the 4 statistics of
``cmcd_loss_stats`` (csrc/reduce.cu).
Nothing here reads a device value back on the host: the step enqueues and returns device tensors, so that the step
can be captured into a CUDA graph (``ShardedStep(graph=True)``) and successive iterations pipeline.
Works with any torch.distributed backend (nccl on the GPU box, gloo in the CPU tests -- there the per-shard
computation is injected and the loss statistics are plain torch reductions).
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


def _loss_stats(l):
    """[sum l, sum l^2, max(-l), sum exp(-l - max)] as a float64 device tensor, no host sync.
    CUDA: one launch of cmcd_loss_stats (double accumulation in the kernel); CPU tensors (gloo tests): torch reductions."""
    if l.is_cuda:
        from .utils import loss_stats
        return loss_stats(l).double() if l.numel() else torch.tensor([0.0, 0.0, -math.inf, 0.0], dtype=torch.float64, device=l.device)
    x = l.detach().double()
    if not x.numel():
        return torch.tensor([0.0, 0.0, -math.inf, 0.0], dtype=torch.float64)
    m = (-x).max()
    return torch.stack([x.sum(), (x * x).sum(), m, torch.exp(-x - m).sum()])


def global_particle_count(n_local, device, group=None):
    """Sum of the ranks' particle counts (host int; call once at set-up, not per step)."""
    world = _world(group)[0]
    if world == 1:
        return int(n_local)
    t = torch.tensor([float(n_local)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(round(t.item()))


def sharded_grad_and_loss(local_forward, seeds_local, params_flat, loss="kl", group=None, n_global=None):
    """One train iteration on this rank's particle shard, enqueue-only.

    ``local_forward(seeds_local, p)`` -> (l_local[N_r] differentiable w.r.t. ``p``, z_local).
    Returns (grad_flat [all-reduced], loss_value, (l_local, z_local)); ``loss_value`` is a 0-dim DEVICE tensor holding the
    global mean (``kl``) or the clipped global variance (``var``) -- read it with ``.item()`` only when the host needs it.
    ``n_global``: total particle count over the ranks if the caller knows it (saves the set-up all-reduce)."""
    world = _world(group)[0]
    p = params_flat.detach().requires_grad_(True)
    with torch.enable_grad():
        l, z = local_forward(seeds_local, p)
    ld = l.detach()
    dev = p.device
    n_local = ld.numel()
    if n_global is None:
        n_global = global_particle_count(n_local, dev, group)
    P = p.numel()
    buf = torch.zeros(P + 3, dtype=torch.float64 if p.dtype == torch.float64 else torch.float32, device=dev)
    if loss == "kl":
        cot = torch.full_like(ld, 1.0 / n_global)
        st = _loss_stats(ld)
        buf[P:P + 2] = st[:2].to(buf.dtype)
        mean = None
    elif loss == "var":
        s = torch.stack([ld.double().sum() if n_local else torch.zeros((), dtype=torch.float64, device=dev)])
        if world > 1:
            dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
        mean = s[0] / n_global
        dev_l = ld.double() - mean
        cot = ((2.0 / n_global) * dev_l).to(ld.dtype)
        buf[P] = s[0].to(buf.dtype)
        buf[P + 1] = (dev_l * dev_l).sum().to(buf.dtype)      # sum (l - mean)^2: the variance without cancellation
    else:
        raise ValueError(loss)
    buf[P + 2:].fill_(float(n_local))
    if l.requires_grad and n_local:
        (g,) = torch.autograd.grad(l, p, grad_outputs=cot.to(l.dtype), allow_unused=True)
        if g is not None:
            buf[:P] = g
    if world > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    g = buf[:P]
    if loss == "kl":
        value = buf[P] / n_global
    else:
        var = buf[P + 1] / n_global
        inside = (var >= -1e7) & (var <= 1e7)      # jnp.clip passes gradient only inside the interval
        value = torch.clamp(var, -1e7, 1e7)
        g = g * inside.to(g.dtype)
    return g, value, (ld, z.detach())


class ShardedStep:
    """The per-rank train step (table chain + forward bridge + adjoint + loss statistics + buffer packing) captured once into a
    CUDA graph and replayed, followed by the one fused all-reduce -- the same capture ``opt.run(graph=True)`` does for the
    single-process loop.  KL loss only (the log-variance loss has a collective between forward and backward and runs through
    ``sharded_grad_and_loss``).  ``__call__(seeds)`` copies the seeds into the static buffer, replays, all-reduces and returns
    device tensors (grad, loss_value, (l_local, z_local)); nothing synchronises."""

    def __init__(self, local_forward, params_flat, n_local, n_global=None, group=None, graph=True, warmup=2):
        self.local_forward, self.p, self.group = local_forward, params_flat, group
        self.n_local = int(n_local)
        dev = params_flat.device
        self.n_global = n_global if n_global is not None else global_particle_count(n_local, dev, group)
        self.world = _world(group)[0]
        self.P = params_flat.numel()
        self.seeds = torch.zeros(self.n_local, dtype=torch.int32, device=dev)
        self.graph = None
        if graph:
            self.seeds.fill_(1)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):          # lazily created workspaces / handles must exist before the capture
                for _ in range(warmup):
                    self._local()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.buf, self.l, self.z = self._local()

    def _local(self):
        p = self.p.detach().requires_grad_(True)
        with torch.enable_grad():
            l, z = self.local_forward(self.seeds, p)
        ld = l.detach()
        buf = torch.zeros(self.P + 3, dtype=torch.float32, device=p.device)
        buf[self.P:self.P + 2] = _loss_stats(ld)[:2].to(buf.dtype)
        buf[self.P + 2:].fill_(float(self.n_local))        # fill_, not an indexed host-scalar copy: legal inside a graph capture
        (g,) = torch.autograd.grad(l, p, grad_outputs=torch.full_like(ld, 1.0 / self.n_global), allow_unused=True)
        if g is not None:
            buf[:self.P] = g
        return buf, ld, z.detach()

    def __call__(self, seeds_local):
        self.seeds.copy_(seeds_local, non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
            buf, l, z = self.buf, self.l, self.z
        else:
            buf, l, z = self._local()
        if self.world > 1:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
        return buf[:self.P], buf[self.P] / self.n_global, (l, z)


def global_ln_z(neg_l_local, n_global=None, group=None):
    """ln Z estimate logsumexp(-l) - log N over all ranks (utils.py:231-237) from the local losses l.  Returns a python float
    (one host read at the end; the combine itself is all-reduce(MAX) + all-reduce(SUM) on device tensors)."""
    st = _loss_stats(neg_l_local.detach())
    m_local, s_local = st[2], st[3]
    world = _world(group)[0]
    m = m_local.clone()
    if world > 1:
        dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    scale = torch.where(torch.isfinite(m_local), torch.exp(m_local - m), torch.zeros_like(m))
    t = torch.stack([s_local * scale, torch.tensor(float(neg_l_local.numel()), dtype=torch.float64, device=st.device)])
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    n = t[1].item() if n_global is None else n_global
    return (torch.log(t[0]) + m - math.log(n)).item()
