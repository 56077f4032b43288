"""The CMCD bound ("machine") and bridge operators restated in torch-CPU.

ORACLE / TEST INFRASTRUCTURE ONLY.  Follows, under /root/reference/src:
  mcdboundingmachine.py:11-123 (initialize), :126-179 (compute_log_elbo), :183-231 (compute_bound[_var])
  mcd_utils.py:14-33 (sample_kernel / log_prob_kernel / evolve dispatch)
  mcd_cais.py:6-99, mcd_cais_var.py:7-112, mcd_over_orig.py:6-65 (the three step bodies)
  mcd_under_lp_a.py:6-87 (underdamped "LDVI" family: MCD_U_a-lp, MCD_U_a-lp-sna, MCD_U_a-lp-sn)
  mcd_under_lp_e.py:6-74 (MCD_U_e-lp, MCD_U_e-lp-sna), mcd_under_lp_ea.py:6-104 (MCD_U_ea-lp-sn)
  mcd_under_lp_a_cais.py:6-115 (MCD_CAIS_UHA_sn, "2nd order CMCD": the dispatcher mcd_utils.py:176-188 passes eps_schedule /
    grad_clipping keywords this function does not take -- a TypeError at HEAD; restated BY SPECIFICATION: the body as written,
    which hard-codes the cosine schedule (:33-40,50) and the 1e2 clip on the target score (:23-30,48, stable=True))
  vardist/diag_gauss.py:20-62, boundingmachine.py:73-111 (nbridges=0 MFVI bound)
  utils.py:219-248 (ELBO / ln Z estimators)
Particles are a leading batch axis (the reference vmaps a per-particle function).  Gaussians
come from the bit-exact threefry restatement (oracle/prng.py) and enter as constants, which is
what jax.grad sees too (noise does not depend on parameters).  Gradients: torch.autograd over
this graph, with target/q scores taken by autograd (create_graph) exactly like jax.grad inside
the step body.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import prng
from .nn import initialize_network
from .pytree import ravel_pytree

_LOG_SQRT_2PI_ARG = math.sqrt(2.0 * math.pi)


# ---------------------------------------------------------------- diag gaussian q (vardist/diag_gauss.py)
def vd_initialize(dim, init_sigma=1.0, dtype=torch.float32):
    return {"mean": torch.zeros(dim, dtype=dtype), "logdiag": torch.ones(dim, dtype=dtype) * math.log(init_sigma)}


def normal_log_prob(x, loc, scale):
    """numpyro 0.10.1 Normal.log_prob summed by Independent(.,1) (mcd_utils.py:19-21)."""
    v = (x - loc) / scale
    return (-0.5 * v * v - torch.log(_LOG_SQRT_2PI_ARG * scale)).sum(-1)


def vd_log_prob(vd, z):
    return normal_log_prob(z, vd["mean"], torch.exp(vd["logdiag"]))


def vd_sample_rep(vd, xi0):
    return torch.exp(vd["logdiag"]) * xi0 + vd["mean"]


# ---------------------------------------------------------------- initialize (mcdboundingmachine.py:11-123)
def initialize(dim, vdparams=None, nbridges=0, eps=0.01, gamma=10.0, eta=0.5, ngridb=32, mgridref_y=None,
               trainable=("eps",), emb_dim=48, seed=1, mode="MCD_CAIS_sn", nn_arch="geffner",
               live=False, dtype=torch.float32):
    pt, pn = {}, {}
    (pt if "vd" in trainable else pn)["vd"] = vdparams if vdparams is not None else vd_initialize(dim, dtype=dtype)
    for name, val in (("eps", eps), ("gamma", gamma), ("eta", eta)):
        (pt if name in trainable else pn)[name] = torch.tensor(float(val), dtype=dtype)
    if mode in ("MCD_ULA_sn", "MCD_CAIS_sn", "MCD_CAIS_var_sn"):
        sn, apply_fun_sn = initialize_network(dim, emb_dim, nbridges, nn_arch,
                                              torch.Generator().manual_seed(seed), live, dtype)
        pt["sn"] = sn
    elif mode in ("MCD_U_a-lp-sna", "MCD_U_e-lp-sna"):   # mcdboundingmachine.py:67-83: network on z only
        sn, apply_fun_sn = initialize_network(dim, emb_dim, nbridges, nn_arch,
                                              torch.Generator().manual_seed(seed), live, dtype)
        pt["sn"] = sn
    elif mode in ("MCD_U_a-lp-sn", "MCD_U_ea-lp-sn", "MCD_CAIS_UHA_sn"):    # :84-102: network on (z, rho), rho_dim = dim
        sn, apply_fun_sn = initialize_network(dim, emb_dim, nbridges, nn_arch,
                                              torch.Generator().manual_seed(seed), live, dtype, rho_dim=dim)
        pt["sn"] = sn
    elif mode in ("MCD_ULA", "MCD_U_a-lp", "MCD_U_e-lp"):
        apply_fun_sn = None
    else:
        raise NotImplementedError("Mode not implemented.")
    if mgridref_y is not None:
        ngridb = mgridref_y.shape[0] - 1
    else:
        ngridb = min(ngridb, nbridges)
        mgridref_y = torch.ones(ngridb + 1, dtype=dtype)
    pn["gridref_x"] = torch.linspace(0, 1, ngridb + 2, dtype=dtype)
    pn["target_x"] = torch.linspace(0, 1, nbridges + 2, dtype=dtype)[1:-1]
    (pt if "mgridref_y" in trainable else pn)["mgridref_y"] = mgridref_y
    params_fixed = (dim, nbridges, mode, apply_fun_sn)
    params_flat, unflatten = ravel_pytree((pt, pn), dtype)
    return params_flat, unflatten, params_fixed


def interp(x, xp, fp):
    """jnp.interp restated (differentiable in fp)."""
    i = torch.clamp(torch.searchsorted(xp, x, right=True), 1, xp.numel() - 1)
    df = fp[i] - fp[i - 1]
    dx = xp[i] - xp[i - 1]
    delta = x - xp[i - 1]
    return torch.where(dx == 0, fp[i], fp[i - 1] + (delta / dx) * df)


def make_betas(params):
    m = params["mgridref_y"]
    gridref_y = torch.cumsum(m, 0) / torch.sum(m)
    gridref_y = torch.cat([torch.zeros(1, dtype=m.dtype), gridref_y])
    return interp(params["target_x"], params["gridref_x"], gridref_y)


def eps_at(eps0, i, nbridges, eps_schedule):
    """mcd_cais.py:34-44,54-59."""
    if eps_schedule == "cos_sq":
        phase = torch.tensor(float(i), dtype=eps0.dtype) / nbridges
        decay = torch.cos((phase + 0.008) / 1.008 * 0.5 * math.pi) ** 2
        return eps0 * decay
    if eps_schedule == "linear":
        return (0.0001 - eps0) / (nbridges - 1) * i + eps0
    return eps0


_ANALYTIC = [False]


class analytic_scores:
    """Context manager for the CPU-BASELINE leg of bench.py only: densities that carry a closed-form ``.score`` (q, many_gmm) use
    it instead of the create_graph autograd call, so the timed baseline is not an eager double-backward strawman.  Same values
    up to rounding (tests/test_oracle_pins.py::test_analytic_scores_match_autograd); parity tests never enable it."""

    def __enter__(self):
        self.prev, _ANALYTIC[0] = _ANALYTIC[0], True

    def __exit__(self, *a):
        _ANALYTIC[0] = self.prev


class _QLogProb:
    """log q as a callable that also knows its closed-form score -(z - mean) / sigma^2 (vardist/diag_gauss.py:28-33)."""

    def __init__(self, vd):
        self.vd = vd

    def __call__(self, x):
        return vd_log_prob(self.vd, x)

    def score(self, x):
        return -(x - self.vd["mean"]) * torch.exp(-2.0 * self.vd["logdiag"])


def _score(fn, z):
    """jax.grad(fn)(z) per particle; differentiable again (second order) when z carries a graph."""
    if _ANALYTIC[0] and hasattr(fn, "score"):
        return fn.score(z)
    zz = z if z.requires_grad else z.detach().requires_grad_(True)
    with torch.enable_grad():
        (g,) = torch.autograd.grad(fn(zz).sum(), zz, create_graph=True)
    return g


# ---------------------------------------------------------------- the step bodies
UD_MODES = ("MCD_U_a-lp", "MCD_U_a-lp-sna", "MCD_U_a-lp-sn", "MCD_U_e-lp", "MCD_U_e-lp-sna", "MCD_U_ea-lp-sn", "MCD_CAIS_UHA_sn")


def evolve_underdamped_lp_a(z, betas, params, rho0, xi, params_fixed, log_prob_model, traj=None):
    """mcd_under_lp_a.py:6-87 (use_sn / full_sn from the mode, mcd_utils.py:83-118).  rho0 [N,d], xi [K,N,d]."""
    dim, nbridges, mode, apply_fun_sn = params_fixed
    if mode in ("MCD_U_e-lp", "MCD_U_e-lp-sna"):
        return evolve_underdamped_lp_e(z, betas, params, rho0, xi, params_fixed, log_prob_model, traj)
    if mode == "MCD_U_ea-lp-sn":
        return evolve_underdamped_lp_ea(z, betas, params, rho0, xi, params_fixed, log_prob_model, traj)
    if mode == "MCD_CAIS_UHA_sn":
        return evolve_underdamped_lp_a_cais(z, betas, params, rho0, xi, params_fixed, log_prob_model, traj)
    vd = params["vd"]
    use_sn, full_sn = mode != "MCD_U_a-lp", mode == "MCD_U_a-lp-sn"
    eps, gamma = params["eps"], params["gamma"]
    zero, one = torch.zeros((), dtype=z.dtype), torch.ones((), dtype=z.dtype)

    def gradU(x, beta):  # jax.grad(U)(z, beta), :18-21
        U = lambda y: -1.0 * (beta * log_prob_model(y) + (1.0 - beta) * vd_log_prob(vd, y))
        return _score(U, x)

    rho = rho0                                           # :62-63
    w = -normal_log_prob(rho, zero, one)                 # :66-67
    if traj is not None:
        traj[-1] = (traj[-1], rho.detach().clone(), None)
    for i in range(nbridges):                            # :23-60
        beta = betas[i]
        eta_aux = gamma * eps
        fk_rho_mean = rho * (1.0 - eta_aux)
        scale = torch.sqrt(2.0 * eta_aux)
        rho_prime = fk_rho_mean + scale * xi[i]
        rho_pp = rho_prime - eps * gradU(z, beta) / 2.0
        z_new = z + eps * rho_pp
        rho_new = rho_pp - eps * gradU(z_new, beta) / 2.0
        bk_rho_mean = rho_prime * (1.0 - eta_aux)
        if use_sn:
            inp = torch.cat([z, rho_prime], dim=-1) if full_sn else z
            bk_rho_mean = bk_rho_mean + 2 * eta_aux * apply_fun_sn(params["sn"], inp, i)
        w = w + normal_log_prob(rho, bk_rho_mean, scale) - normal_log_prob(rho_prime, fk_rho_mean, scale)
        if traj is not None:
            traj[-1] = (traj[-1][0], traj[-1][1], rho_prime.detach().clone())
            traj.append((z_new.detach().clone(), rho_new.detach().clone(), None))
        z, rho = z_new, rho_new
    w = w + normal_log_prob(rho, zero, one)              # :83-84
    return z, w


def _ud_scan(z, betas, params, rho0, xi, nbridges, log_prob_model, kernels, traj):
    """Shared scaffolding of the three underdamped scans (initial / final momentum terms, leapfrog); ``kernels(i, z, rho)``
    returns (fk_mean, fk_scale, bk(rho_prime) -> (bk_mean, bk_scale)) exactly as the respective reference body forms them."""
    vd = params["vd"]
    eps = params["eps"]
    zero, one = torch.zeros((), dtype=z.dtype), torch.ones((), dtype=z.dtype)

    def gradU(x, beta):
        U = lambda y: -1.0 * (beta * log_prob_model(y) + (1.0 - beta) * vd_log_prob(vd, y))
        return _score(U, x)

    rho = rho0
    w = -normal_log_prob(rho, zero, one)
    if traj is not None:
        traj[-1] = (traj[-1], rho.detach().clone(), None)
    for i in range(nbridges):
        beta = betas[i]
        fk_rho_mean, scale_f, bk = kernels(i, z, rho)
        rho_prime = fk_rho_mean + scale_f * xi[i]
        rho_pp = rho_prime - eps * gradU(z, beta) / 2.0
        z_new = z + eps * rho_pp
        rho_new = rho_pp - eps * gradU(z_new, beta) / 2.0
        bk_rho_mean, scale_b = bk(rho_prime)
        w = w + normal_log_prob(rho, bk_rho_mean, scale_b) - normal_log_prob(rho_prime, fk_rho_mean, scale_f)
        if traj is not None:
            traj[-1] = (traj[-1][0], traj[-1][1], rho_prime.detach().clone())
            traj.append((z_new.detach().clone(), rho_new.detach().clone(), None))
        z, rho = z_new, rho_new
    w = w + normal_log_prob(rho, zero, one)
    return z, w


def evolve_underdamped_lp_a_cais(z, betas, params, rho0, xi, params_fixed, log_prob_model, traj=None):
    """mcd_under_lp_a_cais.py:6-115 as written (see the module header): controlled underdamped step -- the network enters the
    forward-kernel mean at (z, rho) and the backward-kernel mean at (z, rho'), eps follows the cosine schedule, the target
    score is clipped at 1e2."""
    dim, nbridges, mode, apply_fun_sn = params_fixed
    vd = params["vd"]
    gamma = params["gamma"]
    zero, one = torch.zeros((), dtype=z.dtype), torch.ones((), dtype=z.dtype)

    def gradU(x, beta, clip=1e2):                                         # :23-30
        gp = _score(lambda y: vd_log_prob(vd, y), x)
        gu = _score(log_prob_model, x)
        guc = torch.clamp(gu, -clip, clip)
        return -1.0 * (beta * guc + (1.0 - beta) * gp)

    rho = rho0                                                            # :94-95
    w = -normal_log_prob(rho, zero, one)                                  # :98-99
    for i in range(nbridges):                                             # :42-90
        beta = betas[i]
        uf = gradU(z, beta)                                               # :47
        eps = eps_at(params["eps"], i, nbridges, "cos_sq")                # :50 -> :33-40
        eta_aux = gamma * eps
        input_sn_old = torch.cat([z, rho], dim=-1)
        fk_rho_mean = rho * (1.0 - eta_aux) - 2.0 * eta_aux * apply_fun_sn(params["sn"], input_sn_old, i)   # :53-56
        scale = torch.sqrt(2.0 * eta_aux)
        rho_prime = fk_rho_mean + scale * xi[i]
        rho_pp = rho_prime - eps * uf / 2.0                               # :64
        z_new = z + eps * rho_pp
        ub = gradU(z_new, beta)
        rho_new = rho_pp - eps * ub / 2.0
        input_sn = torch.cat([z, rho_prime], dim=-1)
        bk_rho_mean = rho_prime * (1.0 - eta_aux) + 2.0 * eta_aux * apply_fun_sn(params["sn"], input_sn, i)  # :79-82
        w = w + normal_log_prob(rho, bk_rho_mean, scale) - normal_log_prob(rho_prime, fk_rho_mean, scale)
        z, rho = z_new, rho_new
    w = w + normal_log_prob(rho, zero, one)                               # :113
    return z, w


def evolve_underdamped_lp_e(z, betas, params, rho0, xi, params_fixed, log_prob_model, traj=None):
    """mcd_under_lp_e.py:6-74: exact OU momentum refresh rho' ~ N(eta rho, 1 - eta^2); MCD_U_e-lp / MCD_U_e-lp-sna."""
    dim, nbridges, mode, apply_fun_sn = params_fixed
    use_sn = mode == "MCD_U_e-lp-sna"
    eta = params["eta"]

    def kernels(i, z_, rho):
        fk_rho_mean = eta * rho                              # :27
        scale = torch.sqrt(1.0 - eta ** 2)                   # :28

        def bk(rho_prime):
            m = eta * rho_prime                              # :39
            if use_sn:
                m = m + 2 * apply_fun_sn(params["sn"], z_, i) * (1.0 - eta)   # :41-43
            return m, scale
        return fk_rho_mean, scale, bk

    return _ud_scan(z, betas, params, rho0, xi, nbridges, log_prob_model, kernels, traj)


def evolve_underdamped_lp_ea(z, betas, params, rho0, xi, params_fixed, log_prob_model, traj=None):
    """mcd_under_lp_ea.py:6-104 (MCD_U_ea-lp-sn: use_sn, full_sn): exact refresh forward, Euler-type backward kernel."""
    dim, nbridges, mode, apply_fun_sn = params_fixed
    eps, gamma = params["eps"], params["gamma"]

    def kernels(i, z_, rho):
        eta = torch.exp(-gamma * eps)                        # :28
        eta_aux = gamma * eps                                # :29
        scale = torch.sqrt(2.0 * eta_aux)                    # :31
        fk_rho_mean = rho * eta                              # :32
        scale_f = torch.sqrt(1.0 - eta ** 2)                 # :33

        def bk(rho_prime):
            inp = torch.cat([z_, rho_prime], dim=-1)         # :54
            m = rho_prime * (1.0 - eta_aux) + 2 * eta_aux * apply_fun_sn(params["sn"], inp, i)   # :55-57
            return m, scale
        return fk_rho_mean, scale_f, bk

    return _ud_scan(z, betas, params, rho0, xi, nbridges, log_prob_model, kernels, traj)


def evolve(z, betas, params, xi, params_fixed, log_prob_model, eps_schedule=None, grad_clipping=False, traj=None):
    """mcd_utils.py:24-190 dispatch + the three scan bodies.  xi [K,N,d] are the per-step Gaussians."""
    dim, nbridges, mode, apply_fun_sn = params_fixed
    vd = params["vd"]
    q_lp = _QLogProb(vd)
    w = torch.zeros(z.shape[0], dtype=z.dtype)
    if mode in ("MCD_ULA", "MCD_ULA_sn"):
        use_sn = mode == "MCD_ULA_sn"
        for i in range(nbridges):  # mcd_over_orig.py:18-55
            beta, eps = betas[i], params["eps"]
            gU = lambda x: -(beta * _score(log_prob_model, x) + (1.0 - beta) * _score(q_lp, x))
            fk_mean = z - eps * gU(z)
            scale = torch.sqrt(2 * eps)
            z_new = fk_mean + scale * xi[i]
            bk_mean = z_new - eps * gU(z_new)
            if use_sn:
                bk_mean = bk_mean + eps * apply_fun_sn(params["sn"], z_new, i)
            w = w + normal_log_prob(z, bk_mean, scale) - normal_log_prob(z_new, fk_mean, scale)
            z = z_new
            if traj is not None:
                traj.append(z.detach().clone())
        return z, w
    if mode not in ("MCD_CAIS_sn", "MCD_CAIS_var_sn"):
        raise NotImplementedError("Mode not implemented.")
    var = mode == "MCD_CAIS_var_sn"
    clip = 1e2 if var else 1e3

    def gradU(x, beta):
        gp = _score(q_lp, x)
        gu = _score(log_prob_model, x)
        if grad_clipping:
            gu = torch.clamp(gu, -clip, clip)
            if var:  # mcd_cais_var.py:33-40 clips both
                gp = torch.clamp(gp, -clip, clip)
        return -(beta * gu + (1.0 - beta) * gp)

    for i in range(nbridges):  # mcd_cais.py:46-89 / mcd_cais_var.py:56-101
        beta = betas[i]
        if var:
            z = z.detach()
        uf = gradU(z, beta)
        eps = eps_at(params["eps"], i, nbridges, eps_schedule)
        fk_mean = z - eps * uf - eps * apply_fun_sn(params["sn"], z, i)
        scale = torch.sqrt(2 * eps)
        z_new = fk_mean + scale * xi[i]
        if var:
            z_new = z_new.detach()
        ub = gradU(z_new, beta)
        bk_mean = z_new - eps * ub + eps * apply_fun_sn(params["sn"], z_new, i + 1)
        w = w + normal_log_prob(z, bk_mean, scale) - normal_log_prob(z_new, fk_mean, scale)
        z = z_new
        if traj is not None:
            traj.append(z.detach().clone())
    return z, w


def compute_log_elbo(seeds, params_flat, unflatten, params_fixed, log_prob, eps_schedule=None,
                     grad_clipping=False, traj=None):
    """mcdboundingmachine.py:126-179, batched over ``seeds``.  Returns (-w [N], z_K [N,d])."""
    pt, pn = unflatten(params_flat)
    pn = _detach_tree(pn)
    params = {**pt, **pn}
    dim, nbridges = params_fixed[0], params_fixed[1]
    dtype = params_flat.dtype
    ud = params_fixed[2] in UD_MODES
    if ud:
        xi0_np, rho0_np, xi_np = prng.particle_noise_ud(np.asarray(seeds), dim, nbridges)
        rho0 = torch.tensor(rho0_np, dtype=dtype)
    else:
        xi0_np, xi_np = prng.particle_noise(np.asarray(seeds), dim, nbridges)
    xi0, xi = torch.tensor(xi0_np, dtype=dtype), torch.tensor(xi_np, dtype=dtype)
    z = vd_sample_rep(params["vd"], xi0)
    w = -vd_log_prob(params["vd"], z)
    if traj is not None:
        traj.append(z.detach().clone())
    if nbridges >= 1:
        betas = make_betas(params)
        if ud:
            z, w_mom = evolve_underdamped_lp_a(z, betas, params, rho0, xi, params_fixed, log_prob, traj)
        else:
            z, w_mom = evolve(z, betas, params, xi, params_fixed, log_prob, eps_schedule, grad_clipping, traj)
        w = w + w_mom
    w = w + log_prob(z)
    return -1.0 * w, z


def _detach_tree(t):
    if isinstance(t, dict):
        return {k: _detach_tree(v) for k, v in t.items()}
    if isinstance(t, (list, tuple)):
        return type(t)(_detach_tree(v) for v in t)
    return t.detach() if isinstance(t, torch.Tensor) else t


def compute_bound(seeds, params_flat, unflatten, params_fixed, log_prob, eps_schedule=None, grad_clipping=False):
    """mcdboundingmachine.py:183-205."""
    l, z = compute_log_elbo(seeds, params_flat, unflatten, params_fixed, log_prob, eps_schedule, grad_clipping)
    return l.mean(), (l, z)


def compute_bound_var(seeds, params_flat, unflatten, params_fixed, log_prob, eps_schedule=None, grad_clipping=False):
    """mcdboundingmachine.py:208-231."""
    l, z = compute_log_elbo(seeds, params_flat, unflatten, params_fixed, log_prob, eps_schedule, grad_clipping)
    return torch.clamp(l.var(unbiased=False), -1e7, 1e7), (l, z)


def grad_and_loss(bound_fn, seeds, params_flat, unflatten, params_fixed, log_prob, **kw):
    """jax.jit(jax.grad(bound_fn, 1, has_aux=True)) of main.py:174-176 -> (grad_flat, (loss[N], z[N,d]))."""
    p = params_flat.detach().clone().requires_grad_(True)
    loss, (l, z) = bound_fn(seeds, p, unflatten, params_fixed, log_prob, **kw)
    (g,) = torch.autograd.grad(loss, p, allow_unused=True)
    if g is None:
        g = torch.zeros_like(p)
    return g, (l.detach(), z.detach())


# ---------------------------------------------------------------- boundingmachine.py (nbridges = 0)
def bm_initialize(dim, vdparams=None, trainable=("vd",), init_sigma=1.0, dtype=torch.float32):
    """boundingmachine.py:9-70 at nbridges=0 (main.py:83-85): the SAME pytree as for nbridges >= 1 -- vd, eps = 0.0, eta = 0.5,
    md = zeros(dim), and the beta-grid leaves of an empty bridge (ngridb = 0: mgridref_y = ones(1), gridref_x = linspace(0, 1, 2),
    target_x = empty) -- so flat vectors are interchangeable with the reference's."""
    return uha_initialize(dim, vdparams=vdparams, nbridges=0, lfsteps=1, eps=0.0, eta=0.5, trainable=trainable, init_sigma=init_sigma,
                          dtype=dtype)


def bm_compute_bound(seeds, params_flat, unflatten, params_fixed, log_prob):
    """boundingmachine.py:73-111 with nbridges=0: loss_n = log q(z) - log p(z), z ~ q."""
    pt, pn = unflatten(params_flat)
    params = {**pt, **_detach_tree(pn)}
    dim = params_fixed[0]
    xi0_np, _ = prng.particle_noise(np.asarray(seeds), dim, 0)
    xi0 = torch.tensor(xi0_np, dtype=params_flat.dtype)
    z = vd_sample_rep(params["vd"], xi0)
    w = -vd_log_prob(params["vd"], z) + log_prob(z)
    l = -1.0 * w
    return l.mean(), (l, z)


# ---------------------------------------------------------------- boundingmachine.py + ais_utils.py (UHA, nbridges >= 1)
def uha_initialize(dim, vdparams=None, nbridges=1, lfsteps=1, eps=0.05, eta=0.5, mdparams=None, ngridb=32, mgridref_y=None,
                   trainable=("eps", "eta"), init_sigma=1.0, dtype=torch.float32):
    """boundingmachine.py:9-70 (boundmode "UHA", main.py:115-131): vd, eps, eta, md (momentum log-scales, momdist.py:9-11),
    mgridref_y; params_fixed = (dim, nbridges, lfsteps)."""
    pt, pn = {}, {}
    (pt if "vd" in trainable else pn)["vd"] = vdparams if vdparams is not None else vd_initialize(dim, init_sigma, dtype)
    for name, val in (("eps", eps), ("eta", eta)):
        (pt if name in trainable else pn)[name] = torch.tensor(float(val), dtype=dtype)
    (pt if "md" in trainable else pn)["md"] = mdparams if mdparams is not None else torch.zeros(dim, dtype=dtype)
    if mgridref_y is not None:
        ngridb = mgridref_y.shape[0] - 1
    else:
        ngridb = min(ngridb, nbridges)
        mgridref_y = torch.ones(ngridb + 1, dtype=dtype)
    pn["gridref_x"] = torch.linspace(0, 1, ngridb + 2, dtype=dtype)
    pn["target_x"] = torch.linspace(0, 1, nbridges + 2, dtype=dtype)[1:-1]
    (pt if "mgridref_y" in trainable else pn)["mgridref_y"] = mgridref_y
    flat, unflatten = ravel_pytree((pt, pn), dtype)
    return flat, unflatten, (dim, nbridges, lfsteps)


def uha_evolve(z, betas, params, rho_noise0, xi, params_fixed, log_prob):
    """ais_utils.py:7-69.  rho_noise0 [N,d], xi [K,N,d]: the N(0,1) draws of md.sample (momdist.py:14-22)."""
    dim, nbridges, lfsteps = params_fixed
    vd, eps, eta, mdp = params["vd"], params["eps"], params["eta"], params["md"]
    zero = torch.zeros((), dtype=z.dtype)
    md_scale = torch.exp(mdp)
    md_log_prob = lambda r: normal_log_prob(r, zero, md_scale)            # momdist.py:25-29

    def gradU(x, beta):                                                   # ais_utils.py:8-9
        U = lambda y: -1.0 * (beta * log_prob(y) + (1.0 - beta) * vd_log_prob(vd, y))
        return _score(U, x)

    gradK = lambda r: _score(lambda y: -1.0 * md_log_prob(y), r)          # :28-29

    rho = md_scale * rho_noise0                                           # :62 -> momdist.py:17-19
    w = torch.zeros(z.shape[0], dtype=z.dtype)
    for i in range(nbridges):                                             # :11-26
        beta = betas[i]
        rho_r = eta * rho + torch.sqrt(1.0 - eta ** 2) * (md_scale * xi[i])   # momdist.py:21
        # leapfrog (:27-56)
        r = rho_r - eps * gradU(z, beta) / 2.0
        zz = z + eps * gradK(r)
        for _ in range(lfsteps - 1):
            r = r - eps * gradU(zz, beta)
            zz = zz + eps * gradK(r)
        r = r - eps * gradU(zz, beta) / 2.0
        w = w + md_log_prob(r) - md_log_prob(rho_r)                       # :21
        z, rho = zz, r
    return z, w


def uha_compute_bound(seeds, params_flat, unflatten, params_fixed, log_prob):
    """boundingmachine.py:73-111 with nbridges >= 1 (the key chain is the one of the underdamped MCD operators:
    one split for the initial momentum, one discarded, then two per bridge; ais_utils.py:61-66,15,22)."""
    pt, pn = unflatten(params_flat)
    params = {**pt, **_detach_tree(pn)}
    dim, nbridges = params_fixed[0], params_fixed[1]
    dtype = params_flat.dtype
    xi0_np, rho0_np, xi_np = prng.particle_noise_ud(np.asarray(seeds), dim, nbridges)
    xi0, rho0, xi = (torch.tensor(a, dtype=dtype) for a in (xi0_np, rho0_np, xi_np))
    z = vd_sample_rep(params["vd"], xi0)
    w = -vd_log_prob(params["vd"], z)
    if nbridges >= 1:
        betas = make_betas(params)
        z, w_mom = uha_evolve(z, betas, params, rho0, xi, params_fixed, log_prob)
        w = w + w_mom
    w = w + log_prob(z)
    l = -1.0 * w
    return l.mean(), (l, z)


# ---------------------------------------------------------------- estimators (utils.py:219-248)
def log_final_losses(eval_losses):
    """eval_losses [n_input_dist_seeds, n_samples] -> dict(elbo, elbo_std, ln_Z, ln_Z_std)."""
    e = torch.as_tensor(eval_losses)
    n = e.shape[1]
    elbos = -e.mean(1)
    lnz = torch.logsumexp(-e, dim=1) - math.log(n)
    return {"elbo": elbos.mean().item(), "elbo_std": elbos.std(unbiased=False).item(),
            "ln_Z": lnz.mean().item(), "ln_Z_std": lnz.std(unbiased=False).item()}
