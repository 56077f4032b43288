"""One experiment end to end -- host side.  Mirrors the body of /root/reference/src/main.py:60-262 (``main``) without wandb /
plotting / Sinkhorn: seeds (:77-79), optional MFVI pretrain (:81-113), ``mcdbm.initialize`` with the flags of
configs/base.py (:134-177), ``opt.run`` (:191-205), ``opt.sample`` + ``log_final_losses`` (:207-221), and the per-model
overrides of ``utils.setup_config`` (utils.py:181-204).  It exists so that the README commands can be replayed through the
CUDA path and compared with the numbers the reference's notebook holds (tests/test_gpu_published.py)."""
from __future__ import annotations

from types import SimpleNamespace

import torch

from . import boundingmachine as bm
from . import mcdboundingmachine as mcdbm
from . import opt
from .model_handler import load_model
from .utils import log_final_losses

# configs/base.py:55-72
LR_DICT = {"lgcp": {"MCD_CAIS_UHA_sn": 1e-3, "MCD_CAIS_sn": 1e-4, "MCD_U_a-lp-sn": 1e-3, "UHA": 1e-4, "MCD_ULA_sn": 1e-4, "MCD_ULA": 1e-4}}
FUNNEL_EPS_DICT = {8: (0.1, 0.01), 16: (0.1, 0.01), 32: (0.1, 0.005), 64: (0.1, 0.001), 128: (0.01, 0.01), 256: (0.01, 0.005)}


def get_config(**overrides):
    """configs/base.py:77-157 (hot-path flags only)."""
    c = dict(boundmode="UHA", model="lorenz", N=5, nbridges=8, lfsteps=1, emb_dim=20, nlayers=3, init_eta=0.0, init_eps=1e-5,
             init_sigma=1.0, pretrain_mfvi=True, train_vi=True, train_eps=True, train_betas=True, nn_arch="geffner",
             eps_schedule="", grad_clipping=False, mfvi_iters=150000, mfvi_lr=0.01, iters=150000, lr=0.0001, seed=1,
             n_samples=500, n_input_dist_seeds=30, use_ema=False, use_whitened=False)
    c.update(overrides)
    return SimpleNamespace(**c)


def setup_config(config):
    """utils.py:181-204: funnel takes (init_eps, lr) from FUNNEL_EPS_DICT[nbridges]; lgcp takes lr from LR_DICT."""
    if config.model == "funnel" and config.nbridges in FUNNEL_EPS_DICT:
        config.init_eps, config.lr = FUNNEL_EPS_DICT[config.nbridges]
    elif config.model in LR_DICT and config.boundmode in LR_DICT[config.model]:
        config.lr = LR_DICT[config.model][config.boundmode]
    return config


def main(config, device="cuda", graph=True, sync_every=500, log=print):
    """main.py:60-262 -> dict(elbo_final, final_ln_Z, elbo_final_std, final_ln_Z_std, losses, params_flat, diverged)."""
    config = setup_config(config)
    out = load_model(config.model, config, device=device)
    log_prob_model, dim = out[0], out[1]
    rng_key_gen = opt.prng_key(config.seed)
    train_rng_key_gen, eval_rng_key_gen = opt.split_key(rng_key_gen)

    trainable = ("vd",)
    params_flat, unflatten, params_fixed = bm.initialize(dim=dim, nbridges=0, trainable=trainable, init_sigma=config.init_sigma,
                                                         device=device)
    result = {}
    if config.pretrain_mfvi:
        gl0 = mcdbm.grad_and_loss(bm.compute_bound)
        r = opt.run(config, config.mfvi_lr, config.mfvi_iters, params_flat, unflatten, params_fixed, log_prob_model, gl0,
                    trainable, train_rng_key_gen, log_prefix="pretrain", use_ema=False, graph=graph, sync_every=sync_every)
        if len(r) == 2:
            return {"diverged": "pretrain"}
        losses, params_flat, _ = r
        result["elbo_init"] = -float(torch.tensor(losses[-500:]).mean())
        log("Done training initial parameters, got ELBO %.2f." % result["elbo_init"])
    vdparams_init = unflatten(params_flat)[0]["vd"]

    if config.boundmode == "UHA":
        trainable = ("eta",) + (("eps",) if config.train_eps else ()) + (("vd",) if config.train_vi else ()) + \
            (("mgridref_y",) if config.train_betas else ())
        params_flat, unflatten, params_fixed = bm.initialize(dim=dim, nbridges=config.nbridges, eta=config.init_eta,
                                                             eps=config.init_eps, lfsteps=config.lfsteps,
                                                             vdparams=vdparams_init, trainable=trainable, device=device)
        loss_fn = bm.compute_bound
    elif "MCD" in config.boundmode:
        trainable = ("eta", "gamma") + (("eps",) if config.train_eps else ()) + (("vd",) if config.train_vi else ()) + \
            (("mgridref_y",) if config.train_betas else ())
        params_flat, unflatten, params_fixed = mcdbm.initialize(dim=dim, nbridges=config.nbridges, vdparams=vdparams_init,
                                                                eta=config.init_eta, eps=config.init_eps, trainable=trainable,
                                                                mode=config.boundmode, emb_dim=config.emb_dim,
                                                                nlayers=config.nlayers, nn_arch=config.nn_arch, device=device)
        fn = mcdbm.compute_bound_var if "var" in config.boundmode else mcdbm.compute_bound
        kw = dict(eps_schedule=config.eps_schedule or None, grad_clipping=config.grad_clipping)
        loss_fn = lambda *a: fn(*a, **kw)
    else:
        raise NotImplementedError("Mode %s not implemented." % config.boundmode)
    grad_and_loss = mcdbm.grad_and_loss(loss_fn)

    r = opt.run(config, config.lr, config.iters, params_flat, unflatten, params_fixed, log_prob_model, grad_and_loss, trainable,
                train_rng_key_gen, log_prefix="train", use_ema=config.use_ema, graph=graph, sync_every=sync_every)
    if len(r) == 2:
        return {"diverged": "train", **result}
    losses, params_flat, ema_params = r
    eval_losses, samples = opt.sample(config, config.n_samples, config.n_input_dist_seeds, params_flat, unflatten, params_fixed,
                                      log_prob_model, loss_fn, eval_rng_key_gen, log_prefix="eval")
    final_elbo, final_ln_Z = log_final_losses(eval_losses)
    log("Done training, got ELBO %.2f." % final_elbo)
    log("Done training, got ln Z %.2f." % final_ln_Z)
    result.update(log_final_losses.last)
    result.update(losses=losses, params_flat=params_flat, samples=samples, diverged=None)
    return result
