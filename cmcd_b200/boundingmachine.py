"""The UHA / MFVI bound -- host side.  Mirrors /root/reference/src/boundingmachine.py:9-111.

``nbridges = 0``: the MFVI-pretrain bound of main.py:81-113 (mean over particles of log q(z) - log p(z), z ~ q).
``nbridges >= 1``: Uncorrected Hamiltonian Annealing (config.boundmode "UHA", main.py:115-133): ais_utils.evolve
(ais_utils.py:7-69) with the diagonal momentum distribution of momdist.py, fused into the CUDA kernels of
csrc/bridge_uha.cu behind ``mcd_utils.uha_bridge``.  Same argument order and return values as the reference."""
from __future__ import annotations

import torch

from . import mcd_utils
from . import variationaldist as vd
from .mcdboundingmachine import make_betas
from .pytree import ravel_pytree, tree_map


def initialize(dim, vdparams=None, nbridges=0, lfsteps=1, eps=0.0, eta=0.5, mdparams=None, ngridb=32,
               mgridref_y=None, trainable=("eps", "eta"), init_sigma=1.0, device="cuda"):
    """boundingmachine.py:9-70: the same pytree for every nbridges, including the MFVI machine (nbridges = 0: ngridb = 0, so
    mgridref_y = ones(1), gridref_x = linspace(0, 1, 2), target_x = empty, md = zeros(dim)) -- flat vectors and checkpoints are
    interchangeable with the reference's."""
    dev = torch.device(device)
    pt, pn = {}, {}
    vdp = vdparams if vdparams is not None else vd.initialize(dim, init_sigma=init_sigma, device=dev)
    (pt if "vd" in trainable else pn)["vd"] = tree_map(lambda t: t.to(dev), vdp)
    for name, val in (("eps", eps), ("eta", eta)):
        (pt if name in trainable else pn)[name] = torch.tensor(float(val), device=dev)
    md = mdparams.to(dev) if mdparams is not None else torch.zeros(dim, device=dev)   # momdist.py:9-11
    (pt if "md" in trainable else pn)["md"] = md
    if mgridref_y is not None:
        ngridb = mgridref_y.shape[0] - 1
        mgridref_y = mgridref_y.to(dev)
    else:
        ngridb = min(ngridb, nbridges)
        mgridref_y = torch.ones(ngridb + 1, device=dev)
    pn["gridref_x"] = torch.linspace(0, 1, ngridb + 2, device=dev)
    pn["target_x"] = torch.linspace(0, 1, nbridges + 2, device=dev)[1:-1]
    (pt if "mgridref_y" in trainable else pn)["mgridref_y"] = mgridref_y
    params_flat, unflatten = ravel_pytree((pt, pn), device=dev)
    return params_flat, unflatten, (dim, nbridges, lfsteps)


def compute_log_elbo(seed, params_flat, unflatten, params_fixed, log_prob):
    """boundingmachine.py:73-104, batched over seeds -> (-w[N], (z[N,d], delta_H)).  delta_H (the leapfrog energy error,
    a diagnostic compute_bound drops) is returned as 0."""
    pt, pn = unflatten(params_flat)
    params = {**pt, **tree_map(lambda t: t.detach(), pn)}
    dim, nbridges = params_fixed[0], params_fixed[1]
    seeds = torch.as_tensor(seed, dtype=torch.int32).reshape(-1)
    if nbridges >= 1:
        negw, z = mcd_utils.uha_bridge(seeds, params, make_betas(params), params_fixed, log_prob)
    else:
        negw, z = mcd_utils.bridge(seeds, params, None, (dim, 0, "MCD_ULA", None), log_prob)
    return negw, (z, torch.zeros((), device=z.device))


def compute_bound(seeds, params_flat, unflatten, params_fixed, log_prob):
    """boundingmachine.py:107-111."""
    ratios, (z, _) = compute_log_elbo(seeds, params_flat, unflatten, params_fixed, log_prob)
    return ratios.mean(), (ratios, z)
