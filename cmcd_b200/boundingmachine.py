"""MFVI-pretrain bound (nbridges = 0).  Mirrors /root/reference/src/boundingmachine.py:9-111
restricted to the path main.py:81-113 uses: ``initialize(dim, nbridges=0, trainable=("vd",), init_sigma)``
and ``compute_bound`` = mean over particles of log q(z) - log p(z), z ~ q.  (nbridges >= 1 is UHA,
outside the hot-path scope.)"""
from __future__ import annotations

import torch

from . import mcd_utils
from . import variationaldist as vd
from .pytree import ravel_pytree, tree_map


def initialize(dim, vdparams=None, nbridges=0, lfsteps=1, eps=0.0, eta=0.5, mdparams=None, ngridb=32,
               mgridref_y=None, trainable=("eps", "eta"), init_sigma=1.0, device="cuda"):
    if nbridges != 0:
        raise NotImplementedError("boundingmachine with nbridges >= 1 is UHA (underdamped), outside the hot-path scope")
    dev = torch.device(device)
    pt, pn = {}, {}
    vdp = vdparams if vdparams is not None else vd.initialize(dim, init_sigma=init_sigma, device=dev)
    (pt if "vd" in trainable else pn)["vd"] = tree_map(lambda t: t.to(dev), vdp)
    for name, val in (("eps", eps), ("eta", eta)):
        (pt if name in trainable else pn)[name] = torch.tensor(float(val), device=dev)
    params_flat, unflatten = ravel_pytree((pt, pn), device=dev)
    return params_flat, unflatten, (dim, 0, lfsteps)


def compute_log_elbo(seed, params_flat, unflatten, params_fixed, log_prob):
    pt, pn = unflatten(params_flat)
    params = {**pt, **tree_map(lambda t: t.detach(), pn)}
    dim = params_fixed[0]
    seeds = torch.as_tensor(seed, dtype=torch.int32).reshape(-1)
    negw, z = mcd_utils.bridge(seeds, params, None, (dim, 0, "MCD_ULA", None), log_prob)
    return negw, (z, torch.zeros((), device=z.device))


def compute_bound(seeds, params_flat, unflatten, params_fixed, log_prob):
    """boundingmachine.py:107-111."""
    ratios, (z, _) = compute_log_elbo(seeds, params_flat, unflatten, params_fixed, log_prob)
    return ratios.mean(), (ratios, z)
