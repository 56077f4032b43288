"""ELBO / ln Z estimators.  Mirrors /root/reference/src/utils.py:219-248 ``log_final_losses`` (the
wandb logging is the caller's business); the reductions run in one CUDA launch
(``cmcd_batched_elbo_lnz``) instead of the reference's per-element ``.item()`` loop (opt.py:193)."""
from __future__ import annotations

import torch

from . import _lib


def batched_elbo_lnz(eval_losses):
    """eval_losses [n_input_dist_seeds, n_samples] (CUDA) -> (elbo[b], lnz[b])."""
    e = torch.as_tensor(eval_losses, dtype=torch.float32)
    _lib.require_cuda(e)
    e = e.contiguous()
    b, n = e.shape
    elbo, lnz = torch.empty(b, device=e.device), torch.empty(b, device=e.device)
    _lib.check(_lib.lib().cmcd_batched_elbo_lnz(_lib.current_stream(), _lib.ptr(e), b, n, _lib.ptr(elbo), _lib.ptr(lnz)))
    return elbo, lnz


def log_final_losses(eval_losses, log_prefix=""):
    """utils.py:219-248 -> (final_elbo, final_ln_Z); also returns the stds as attributes of the result dict."""
    elbo, lnz = batched_elbo_lnz(eval_losses)
    out = {f"elbo_final{log_prefix}": elbo.mean().item(), f"final_ln_Z{log_prefix}": lnz.mean().item(),
           f"elbo_final_std{log_prefix}": elbo.std(unbiased=False).item(),
           f"final_ln_Z_std{log_prefix}": lnz.std(unbiased=False).item()}
    log_final_losses.last = out
    return out[f"elbo_final{log_prefix}"], out[f"final_ln_Z{log_prefix}"]


def loss_stats(negw):
    """[sum l, sum l^2, max(-l), sum exp(-l - max)] on the device (cmcd_loss_stats)."""
    _lib.require_cuda(negw)
    out = torch.empty(4, device=negw.device)
    _lib.check(_lib.lib().cmcd_loss_stats(_lib.current_stream(), _lib.ptr(negw.contiguous()), negw.numel(), _lib.ptr(out)))
    _lib.count_launches(1)
    return out
