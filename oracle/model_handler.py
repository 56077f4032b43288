"""Target log-densities of the CMCD hot path, restated in torch-CPU (batched over particles).

ORACLE / TEST INFRASTRUCTURE ONLY.  Follows /root/reference/src/model_handler.py
(funnel :124-154, gmm :157-242, many_gmm :245-284, lgcp :287-409) and
/root/reference/src/cp_utils.py.  Each ``log_prob`` maps [N,d] -> [N]; the reference's
functions are per-particle and vmapped (mcdboundingmachine.py:193), which is the same thing.
Operation order follows the reference's libraries (numpyro 0.10.1 / distrax 0.1.2 /
jax.scipy) where it matters for fp32 rounding.
"""
from __future__ import annotations

import math
import os
from types import SimpleNamespace

import numpy as np
import torch

from . import prng

_HALF_LOG2PI = 0.5 * math.log(2.0 * math.pi)


def default_config(**kw):
    """The subset of configs/base.py:77-157 that load_model reads."""
    cfg = dict(funnel_d=10, funnel_sig=3, funnel_clipy=11, use_whitened=False, n_mixes=40,
               loc_scaling=40, file_path=os.path.join(os.path.dirname(__file__), "..", "cmcd_b200", "data", "pines.csv"))
    cfg.update(kw)
    return SimpleNamespace(**cfg)


# ------------------------------------------------------------------ funnel (model_handler.py:124-154)
def load_model_funnel(config, dtype=torch.float32):
    d = config.funnel_d

    def neg_energy(x):
        v = x[:, 0]
        # norm.logpdf(v, 0, 3): -(v-0)^2/(2*9) - log(3) - 0.5 log(2 pi)
        log_density_v = -0.5 * (v / 3.0) ** 2 - math.log(3.0) - _HALF_LOG2PI
        # multivariate_normal.logpdf(x[1:], 0, eye*exp(v)): L = sqrt(exp(v)) I, y = x/L
        var = torch.exp(v)
        ldiag = torch.sqrt(var)
        y = x[:, 1:] / ldiag[:, None]
        n = d - 1
        log_density_other = -0.5 * (y * y).sum(-1) - 0.5 * n * math.log(2 * math.pi) - n * torch.log(ldiag)
        return log_density_v + log_density_other

    return neg_energy, d


# ------------------------------------------------------------------ gmm (model_handler.py:157-242)
GMM2_MEANS = [[3.0, 0.0], [-2.5, 0.0], [2.0, 3.0]]
GMM2_COVS = [[[0.7, 0.0], [0.0, 0.05]], [[0.7, 0.0], [0.0, 0.05]], [[1.0, 0.95], [0.95, 1.0]]]


def load_model_gmm(config=None, dtype=torch.float32):
    means = torch.tensor(GMM2_MEANS, dtype=dtype)
    covs = torch.tensor(GMM2_COVS, dtype=dtype)
    chol = torch.linalg.cholesky(covs)
    log_w = torch.log(torch.tensor([1.0 / 3] * 3, dtype=dtype))
    norm_term = -1.0 * math.log(2 * math.pi) - torch.log(torch.diagonal(chol, dim1=-2, dim2=-1)).sum(-1)

    def raw(x):  # x [N,2]
        diff = x[:, None, :] - means[None]  # [N,3,2]
        y = torch.linalg.solve_triangular(chol[None].expand(x.shape[0], -1, -1, -1), diff[..., None], upper=False)[..., 0]
        maha = -0.5 * (y * y).sum(-1)
        return torch.logsumexp(maha + norm_term + log_w, dim=-1)

    def log_prob(x):
        a = raw(x)
        b = raw(torch.flip(x, dims=(-1,)))
        # np.logaddexp(a, b) (:194); torch.logaddexp's *second* derivative is NaN when one branch underflows
        # (jnp.logaddexp has a custom JVP that is not), so the mathematically identical 2-way logsumexp is used
        return torch.logsumexp(torch.stack([a, b], -1), -1) - math.log(2.0)

    return log_prob, 2


# ------------------------------------------------------------------ many_gmm (model_handler.py:245-284)
def many_gmm_params(n_mixes=40, loc_scaling=40, dim=2, seed=0):
    """means = uniform(PRNGKey(seed), (n_mixes, dim), -1, 1) * loc_scaling (:256-261); scale = softplus(0.1) (:262-263)."""
    u = prng.uniform(prng.prng_key(seed), n_mixes * dim, -1.0, 1.0).reshape(n_mixes, dim)
    means = (u * np.float32(loc_scaling)).astype(np.float32)
    scale = np.float32(np.log1p(np.exp(np.float32(0.1))))  # jax.nn.softplus(0.1) = logaddexp(0.1, 0)
    return means, scale


def load_model_manygmm(config, dtype=torch.float32):
    means_np, scale_np = many_gmm_params(config.n_mixes, config.loc_scaling)
    means = torch.tensor(means_np, dtype=dtype)
    scale = float(scale_np)
    log_mix = -math.log(config.n_mixes)  # log_softmax(ones)

    def log_prob(x):
        st = (x[:, None, :] - means[None]) / scale
        lp_comp = (-0.5 * st * st - (_HALF_LOG2PI + math.log(scale))).sum(-1)  # distrax Normal.log_prob, Independent
        lp = torch.logsumexp(lp_comp + log_mix, dim=-1)
        valid = lp > -1e4  # :279-280
        return torch.where(valid, lp, torch.full_like(lp, -float("inf")))

    def score(x):
        """Closed form of grad log p (responsibility-weighted pull to the means; zero where the -inf override fires) -- used only
        by the CPU-baseline timing leg (oracle.mcdboundingmachine.analytic_scores)."""
        diff = means[None] - x[:, None, :]
        lp_comp = (-0.5 * (diff / scale) ** 2).sum(-1)
        r = torch.softmax(lp_comp, dim=-1)
        sc = (r[..., None] * diff).sum(1) / (scale * scale)
        lp = torch.logsumexp(lp_comp, dim=-1) - 2.0 * (_HALF_LOG2PI + math.log(scale)) + log_mix
        return torch.where((lp > -1e4)[:, None], sc, torch.zeros_like(sc))

    log_prob.score = score
    return log_prob, 2


# ------------------------------------------------------------------ lgcp (model_handler.py:287-409, cp_utils.py)
def lgcp_constants(file_path, num_dim=1600):
    """float64 numpy construction of the LGCP constants; consumers cast to their dtype."""
    m = int(round(math.sqrt(num_dim)))
    pts = np.genfromtxt(file_path, delimiter=",")
    counts = np.zeros((m, m))
    for elem in pts * m:  # cp_utils.py:16-42
        r, c = int(np.floor(elem[0])), int(np.floor(elem[1]))
        r -= r == m
        c -= c == m
        counts[r, c] += 1
    gi = np.arange(m)
    bin_vals = np.array([[a, b] for a in gi for b in gi], dtype=np.float64)  # itertools.product order
    dist = np.linalg.norm(bin_vals[:, None, :] - bin_vals[None], axis=-1)
    beta = 1.0 / 33
    gram = 1.91 * np.exp(-dist / (m * beta))  # cp_utils.py:81-84
    chol = np.linalg.cholesky(gram)
    mu_zero = math.log(126.0) - 0.5 * 1.91
    return dict(counts=counts.reshape(-1), gram=gram, chol=chol, mu_zero=mu_zero,
                half_log_det=float(np.sum(np.log(np.abs(np.diag(chol))))), poisson_a=1.0 / num_dim)


def load_model_lgcp(config, dtype=torch.float32):
    c = lgcp_constants(config.file_path)
    d = 1600
    chol = torch.tensor(c["chol"], dtype=dtype)
    counts = torch.tensor(c["counts"], dtype=dtype)
    mu0 = c["mu_zero"]
    log_norm = -0.5 * d * math.log(2.0 * math.pi) - c["half_log_det"]
    a = c["poisson_a"]
    if config.use_whitened:
        # model_handler.py:373-384: density of the WHITENED variable e, latent = L e + mu0 (cp_utils.py:107-128)
        white_norm = -0.5 * d * math.log(2.0 * math.pi)

        def log_prob_white(white):
            latent = white @ chol.T + mu0
            return white_norm - 0.5 * (white * white).sum(-1) + (latent * counts - a * torch.exp(latent)).sum(-1)

        return log_prob_white, d

    def log_prob(x):
        white = torch.linalg.solve_triangular(chol, (x - mu0).T, upper=False).T  # cp_utils.py:153
        prior = -0.5 * (white * white).sum(-1) + log_norm
        lik = (x * counts - a * torch.exp(x)).sum(-1)  # cp_utils.py:102-104
        return prior + lik

    return log_prob, d


def load_model(model, config=None, dtype=torch.float32):
    """model_handler.py:30-43 dispatch (hot-path targets only)."""
    config = config or default_config()
    if "funnel" in model:
        return load_model_funnel(config, dtype)
    if "lgcp" in model:
        return load_model_lgcp(config, dtype)
    if "many_gmm" in model:
        return load_model_manygmm(config, dtype)
    if "gmm" in model:
        return load_model_gmm(config, dtype)
    raise NotImplementedError(f"target {model!r} is outside the hot-path scope")
