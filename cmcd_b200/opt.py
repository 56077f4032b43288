"""Training and evaluation loops -- host side.  Mirrors /root/reference/src/opt.py:
``project`` (:14-24), ``create_optimizer`` (:26-35), ``run`` (:67-164), ``sample`` (:167-197).

What the reference does per iteration on the host or in separate XLA dispatches -- drawing the seeds
(``jax.random.randint``), ``optax.chain(clip(5.0), adam)`` + ``apply_updates`` + the un-jitted ``project``, the EMA
copy, and the NaN check with its device sync -- runs here as two launches (``cmcd_randint``,
``cmcd_adam_project_step``) with the divergence guard evaluated on the device, so iterations pipeline.  wandb
logging and plotting are the caller's business (``log_fn``)."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .model_handler import Target, _threefry2x32

_PROJECT_BOUNDS = {"eps": (0.0000001, 0.5), "eta": (0.0, 0.99), "gamma": (0.001, float("inf")),
                   "mgridref_y": (0.001, float("inf"))}   # relu(x - 0.001) + 0.001 == max(x, 0.001)  (opt.py:22-23)


# ------------------------------------------------------------------ PRNG keys (host side, O(1) per iteration)
def prng_key(seed):
    """jax.random.PRNGKey(seed) for a 32-bit seed."""
    return np.array([0, np.uint32(seed)], dtype=np.uint32)


def split_key(key):
    """jax.random.split(key) -> (first, second): threefry over counts iota(4) = blocks (0,2), (1,3)."""
    y0, y1 = _threefry2x32(np.uint32(key[0]), np.uint32(key[1]), np.array([0, 1], np.uint32), np.array([2, 3], np.uint32))
    return np.array([y0[0], y0[1]], np.uint32), np.array([y1[0], y1[1]], np.uint32)


def randint_seeds(key, n, minval=1, maxval=10**6, device="cuda"):
    """jax.random.randint(key, (n,), minval, maxval) on the device (opt.py:93-94)."""
    out = torch.empty(n, dtype=torch.int32, device=device)
    _lib.require_cuda(out)
    _lib.check(_lib.lib().cmcd_randint(_lib.current_stream(), int(key[0]), int(key[1]), n, int(minval), int(maxval), _lib.ptr(out)))
    return out


# ------------------------------------------------------------------ project / optimizer
def projection_bounds(params_flat, unflatten, trainable):
    """(lo, hi) vectors such that project(x) == clamp(x, lo, hi) (opt.py:14-24)."""
    n = params_flat.numel()
    idx_train, _ = unflatten(torch.arange(n, dtype=torch.float32, device=params_flat.device))
    lo = torch.full((n,), -float("inf"), device=params_flat.device)
    hi = torch.full((n,), float("inf"), device=params_flat.device)
    for name, (a, b) in _PROJECT_BOUNDS.items():
        if name in trainable and name in idx_train:
            ix = idx_train[name].reshape(-1).long()
            lo[ix], hi[ix] = a, b
    return lo, hi


def project(x, unflatten, trainable):
    """opt.py:14-24 (functional form)."""
    lo, hi = projection_bounds(x, unflatten, trainable)
    return torch.minimum(torch.maximum(x, lo), hi)


class Optimizer:
    """optax.chain(optax.clip(5.0), optax.adam(step_size, b1, b2, eps)) with the update, apply and project fused."""

    def __init__(self, step_size, b1=0.9, b2=0.999, eps=1e-8, clip=5.0):
        self.lr, self.b1, self.b2, self.eps, self.clip = float(step_size), b1, b2, eps, clip

    def init(self, params_flat):
        return {"m": torch.zeros_like(params_flat), "v": torch.zeros_like(params_flat), "count": 0}

    def step(self, params_flat, grad, state, lo=None, hi=None, ema=None, ema_step=0.001, skip_flag=None):
        """In place: params_flat, state (and ema).  Returns params_flat."""
        _lib.require_cuda(params_flat, grad)
        state["count"] += 1
        _lib.check(_lib.lib().cmcd_adam_project_step(
            _lib.current_stream(), _lib.ptr(params_flat), _lib.ptr(grad.contiguous()), _lib.ptr(state["m"]), _lib.ptr(state["v"]),
            _lib.ptr(lo), _lib.ptr(hi), params_flat.numel(), self.lr, self.b1, self.b2, self.eps, self.clip, state["count"],
            _lib.ptr(ema), float(ema_step), _lib.ptr(skip_flag)))
        return params_flat


def create_optimizer(step_size, b1=0.9, b2=0.999, eps=1e-8, trainable=None):
    """opt.py:26-35."""
    return Optimizer(step_size, b1, b2, eps)


# ------------------------------------------------------------------ loops
def run(info, lr, iters, params_flat, unflatten, params_fixed, log_prob_model, grad_and_loss, trainable, rng_key_gen,
        extra=True, log_prefix="", target_samples=None, use_ema=False, log_fn=None, sync_every=1, graph=False):
    """opt.py:67-164.  Returns (losses, params_flat, ema_params); on divergence prints "Diverged" and returns
    (params_flat, ema_params) exactly like the reference (:122-124).  ``sync_every``: how often the device-side
    divergence flag is read back (1 = the reference's per-iteration check; the update itself is always guarded on
    the device, so a larger value never applies a NaN gradient).  ``graph=True`` captures one iteration's device work
    (table chain, forward bridge, adjoint, loss mean, divergence flag) into a CUDA graph once and replays it: per
    iteration the host then issues three launches (seeds, graph, optimizer) instead of the ~100 small ones of the
    O(K) table chain -- the launch-bound README configs (nbridges = 8, N = 300) run several times faster.  Same
    arithmetic, same order; every C-ABI entry point is enqueue-only, which is what makes the capture legal."""
    optimizer = create_optimizer(lr, trainable=trainable)
    params_flat = params_flat.detach().clone()
    opt_state = optimizer.init(params_flat)
    ema_params = params_flat.clone() if use_ema else None
    lo, hi = projection_bounds(params_flat, unflatten, trainable)
    n_particles = int(getattr(info, "N"))
    losses = []
    flag = torch.zeros((), dtype=torch.int32, device=params_flat.device)

    def device_work(seeds):
        """One iteration up to (not including) the optimizer update; writes the sticky divergence flag in place."""
        grad, (loss, z) = grad_and_loss(seeds, params_flat, unflatten, params_fixed, log_prob_model)
        ema_loss = None
        if use_ema:
            _, (ema_loss, _z_ema) = grad_and_loss(seeds, ema_params, unflatten, params_fixed, log_prob_model)
            ema_loss = ema_loss.mean()
        mean_loss = loss.mean()
        flag.copy_(torch.logical_or(flag.bool(), torch.isnan(mean_loss)).to(torch.int32))   # sticky divergence flag (device)
        return grad, mean_loss, ema_loss

    cuda_graph = None
    if graph and iters > 0:
        static_seeds = torch.zeros(n_particles, dtype=torch.int32, device=params_flat.device)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):      # warm-up on a side stream (workspaces, lazily created handles), flag restored after
            static_seeds.fill_(1)
            for _ in range(2):
                device_work(static_seeds)
            flag.zero_()
        torch.cuda.current_stream().wait_stream(side)
        cuda_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cuda_graph):
            g_grad, g_mean, g_ema = device_work(static_seeds)
        flag.zero_()                       # the capture pass does not execute, but keep the invariant explicit

    for i in range(iters):
        rng_key, rng_key_gen = split_key(rng_key_gen)
        seeds = randint_seeds(rng_key, n_particles, 1, 10**6, device=params_flat.device)
        if cuda_graph is not None:
            static_seeds.copy_(seeds)
            cuda_graph.replay()
            grad, mean_loss, ema_loss = g_grad, g_mean, g_ema
        else:
            grad, mean_loss, ema_loss = device_work(seeds)
        if (i + 1) % max(sync_every, 1) == 0 and flag.item():
            print("Diverged")
            return params_flat, ema_params
        optimizer.step(params_flat, grad, opt_state, lo, hi, ema_params, 0.001, skip_flag=flag)
        if i % max(iters // 1000, 1) == 0:
            losses.append(mean_loss.clone() if cuda_graph is not None else mean_loss)   # stays on the device; converted once at the end
            if log_fn is not None:
                log_fn({f"{log_prefix}/loss": mean_loss, f"{log_prefix}/grad": grad.mean(), "train_step": i,
                        f"{log_prefix}/ema_loss": ema_loss})
    if flag.item():
        print("Diverged")
        return params_flat, ema_params
    return [float(v) for v in torch.stack(losses).cpu()] if losses else [], params_flat, ema_params


def sample(info, n_samples, n_input_dist_seeds, params_flat, unflatten, params_fixed, log_prob_model, loss_fn, rng_key_gen,
           log_prefix=""):
    """opt.py:167-197.  Returns (elbos [n_input_dist_seeds, n_samples] device tensor, zs [n_input_dist_seeds * n_samples, d]);
    the reference's per-element ``.item()`` loop (:193) is gone -- ``utils.log_final_losses`` reduces the tensor in one launch."""
    dev = params_flat.device
    eval_seeds = randint_seeds(rng_key_gen, n_samples * n_input_dist_seeds, 1, 10**6, device=dev)
    elbos, zs = [], []
    # Every particle's result depends on its own seed only, so the small-d targets take all seed batches in ONE pass (one table
    # chain, one bridge launch with n_input_dist_seeds x n_samples particles instead of 30 latency-bound launches); the wide path
    # (lgcp, callback targets: workspace proportional to the batch) keeps the reference's loop.
    if isinstance(log_prob_model, Target) and log_prob_model.kind != "lgcp":
        with torch.no_grad():
            _, (loss_list, z) = loss_fn(eval_seeds, params_flat, unflatten, params_fixed, log_prob_model)
        return loss_list.reshape(n_input_dist_seeds, n_samples), z
    with torch.no_grad():
        for i in range(n_input_dist_seeds):
            seeds = eval_seeds[i * n_samples:(i + 1) * n_samples]
            _, (loss_list, z) = loss_fn(seeds, params_flat, unflatten, params_fixed, log_prob_model)
            zs.append(z)
            elbos.append(loss_list)
    return torch.stack(elbos), torch.cat(zs, dim=0)
