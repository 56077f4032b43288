"""Target registry -- host side.  Mirrors /root/reference/src/model_handler.py:30-43 ``load_model``.

The reference returns an opaque JAX callable; a fused kernel needs a closed registry, so
``load_model`` returns a ``Target``: still callable on a batch ``x[N,d]`` (log density, via the
CUDA ``cmcd_target_eval`` entry -- used for plotting, utils.py:54) and tagged with the device
buffers the bridge kernels read.  Targets outside the registry raise (no generic fallback).
Constants follow model_handler.py: funnel :124-154, gmm :157-242, many_gmm :245-284,
lgcp :287-409 + cp_utils.py.
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch

from . import _lib
from ._lib import CmcdTarget, MIX_STRIDE, TARGET, TARGET_FN

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


# ---- host-side threefry for the many_gmm means (jax.random.uniform(PRNGKey(0), (40,2), -1, 1)) ----
def _threefry2x32(k0, k1, x0, x1):
    u = np.uint32
    rot = ((13, 15, 26, 6), (17, 29, 16, 24))
    with np.errstate(over="ignore"):
        ks = (u(k0), u(k1), u(k0) ^ u(k1) ^ u(0x1BD11BDA))
        x0 = x0.astype(u) + ks[0]
        x1 = x1.astype(u) + ks[1]
        for g in range(5):
            for r in rot[g % 2]:
                x0 = x0 + x1
                x1 = (x1 << u(r)) | (x1 >> u(32 - r))
                x1 = x1 ^ x0
            x0 = x0 + ks[(g + 1) % 3]
            x1 = x1 + ks[(g + 2) % 3] + u(g + 1)
    return x0, x1


def many_gmm_means(n_mixes=40, dim=2, loc_scaling=40.0, seed=0):
    n = n_mixes * dim
    m = (n + 1) // 2
    c = np.arange(2 * m, dtype=np.uint32)
    c[n:] = 0
    y0, y1 = _threefry2x32(0, seed, c[:m], c[m:])
    bits = np.concatenate([y0, y1])[:n]
    f = ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0)
    u = np.maximum(np.float32(-1.0), f * np.float32(2.0) + np.float32(-1.0))
    return (u.reshape(n_mixes, dim) * np.float32(loc_scaling)).astype(np.float32)


class Target:
    """A registry target: callable log density + kernel descriptor."""

    def __init__(self, kind, dim, device, ncomp=0, scale=1.0, invalid_below=-math.inf, mix=None, lgcp=None):
        self.kind, self.dim, self.device = kind, dim, torch.device(device)
        self.ncomp, self.scale, self.invalid_below = ncomp, scale, invalid_below
        self.mix = None if mix is None else torch.as_tensor(mix, dtype=torch.float32).contiguous().to(self.device)
        self.lgcp = lgcp  # dict of device tensors + scalars

    def desc(self):
        t = CmcdTarget()
        t.kind, t.ncomp, t.scale, t.invalid_below = TARGET[self.kind], self.ncomp, self.scale, self.invalid_below
        t.mix = _lib.ptr(self.mix)
        if self.lgcp is not None:
            t.lgcp_kinv, t.lgcp_linv = _lib.ptr(self.lgcp["kinv"]), _lib.ptr(self.lgcp["linv"])
            t.lgcp_counts = _lib.ptr(self.lgcp["counts"])
            t.lgcp_mu0, t.lgcp_log_norm = self.lgcp["mu0"], self.lgcp["log_norm"]
            t.lgcp_bin_area = self.lgcp["bin_area"]
        return t

    def evaluate(self, x, v=None):
        """(log p, score[, hvp]) at x[N,d] on the GPU."""
        _lib.require_cuda(x)
        x = x.to(torch.float32).contiguous()
        n = x.shape[0]
        lp = torch.empty(n, device=x.device)
        sc = torch.empty_like(x)
        hv = torch.empty_like(x) if v is not None else None
        vv = None if v is None else v.to(torch.float32).contiguous()
        d = self.desc()
        _lib.check(_lib.lib().cmcd_target_eval(d, self.dim, _lib.current_stream(), _lib.ptr(x), n, _lib.ptr(vv),
                                               _lib.ptr(lp), _lib.ptr(sc), _lib.ptr(hv)))
        return (lp, sc) if v is None else (lp, sc, hv)

    def __call__(self, x):
        single = x.dim() == 1
        lp = self.evaluate(x[None] if single else x)[0]
        return lp[0] if single else lp


class _DevArray:
    """Raw device pointer -> torch tensor view (through __cuda_array_interface__, no copy)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


def _view(ptr, shape, device):
    return torch.as_tensor(_DevArray(ptr, shape), device=device)


class CallbackTarget:
    """A target outside the registry, given as a batched torch log density ``log_prob(x[N,d]) -> [N]`` on the device --
    what the reference gets from an arbitrary ``log_prob_model`` (inference-gym / numpyro models, model_handler.py:46-86) and
    differentiates with jax.grad inside the step body.  The bridge's step-wise CUDA path calls back between half-steps
    (``cmcd_target_fn``, include/cmcd_b200.h): score by torch autograd, Hessian-vector product by double backward, both
    enqueued on the stream the bridge runs on.  The O(N K) bridge arithmetic stays in the CUDA kernels."""

    kind = "callback"

    def __init__(self, log_prob, dim, device="cuda", closed_form=None):
        """``closed_form(x, v) -> (log p [N], score [N,d], H v [N,d] or None)``: optional analytic evaluation used instead of
        autograd / double backward (same contract, fewer launches)."""
        self.log_prob, self.dim, self.device = log_prob, int(dim), torch.device(device)
        self.closed_form = closed_form
        self.calls = 0

        def _eval(user, stream, x, n, dim_, v, out_logp, out_score, out_hvp):
            try:
                with torch.cuda.stream(torch.cuda.ExternalStream(int(stream or 0), device=self.device)), torch.enable_grad():
                    self.calls += 1
                    if self.closed_form is not None:
                        with torch.no_grad():
                            vt = _view(v, (n, dim_), self.device) if (v and out_hvp) else None
                            lp, sc, hv = self.closed_form(_view(x, (n, dim_), self.device), vt)
                            if out_logp:
                                _view(out_logp, (n,), self.device).copy_(lp)
                            if out_score:
                                _view(out_score, (n, dim_), self.device).copy_(sc)
                            if vt is not None:
                                _view(out_hvp, (n, dim_), self.device).copy_(hv)
                        return 0
                    xt = _view(x, (n, dim_), self.device).detach().clone().requires_grad_(True)
                    lp = self.log_prob(xt)
                    want_hvp = bool(v) and bool(out_hvp)
                    (sc,) = torch.autograd.grad(lp.sum(), xt, create_graph=want_hvp)
                    if out_logp:
                        _view(out_logp, (n,), self.device).copy_(lp.detach())
                    if out_score:
                        _view(out_score, (n, dim_), self.device).copy_(sc.detach())
                    if want_hvp:
                        (hv,) = torch.autograd.grad((sc * _view(v, (n, dim_), self.device)).sum(), xt)
                        _view(out_hvp, (n, dim_), self.device).copy_(hv)
                return 0
            except Exception as e:  # a Python exception must not unwind through the C caller
                self.error = e
                return 1

        self._cb = TARGET_FN(_eval)   # keep the trampoline alive as long as the target

    def desc(self):
        import ctypes
        t = CmcdTarget()
        t.kind = TARGET["callback"]
        t.eval = ctypes.cast(self._cb, ctypes.c_void_p).value
        t.user = None
        return t

    def evaluate(self, x, v=None):
        if self.closed_form is not None:
            with torch.no_grad():
                lp, sc, hv = self.closed_form(x.detach().to(torch.float32), v)
            return (lp, sc) if v is None else (lp, sc, hv)
        x = x.detach().to(torch.float32).clone().requires_grad_(True)
        with torch.enable_grad():
            lp = self.log_prob(x)
            (sc,) = torch.autograd.grad(lp.sum(), x, create_graph=v is not None)
            if v is None:
                return lp.detach(), sc.detach()
            (hv,) = torch.autograd.grad((sc * v).sum(), x)
        return lp.detach(), sc.detach(), hv

    def __call__(self, x):
        single = x.dim() == 1
        lp = self.log_prob(x[None] if single else x)
        return lp[0] if single else lp


def callback_target(log_prob, dim, device="cuda"):
    """(log_prob_model, dim) for a target that is not in the fused registry (model_handler.py:66-86 returns the same pair)."""
    return CallbackTarget(log_prob, dim, device), int(dim)


def _gmm2_components():
    """model_handler.py:164-195: 3 full-covariance components, symmetrised by the coordinate flip."""
    means = np.array([[3.0, 0.0], [-2.5, 0.0], [2.0, 3.0]])
    covs = np.array([[[0.7, 0.0], [0.0, 0.05]], [[0.7, 0.0], [0.0, 0.05]], [[1.0, 0.95], [0.95, 1.0]]])
    rows = []
    for flip in (False, True):
        for m, c in zip(means, covs):
            if flip:  # raw(flip x): mean and covariance with swapped coordinates
                m = m[::-1]
                c = c[::-1, ::-1]
            p = np.linalg.inv(c)
            logc = -math.log(2 * math.pi) - 0.5 * math.log(np.linalg.det(c)) + math.log(1.0 / 3) - math.log(2.0)
            rows.append([m[0], m[1], p[0, 0], p[0, 1], p[1, 1], logc])
    return np.array(rows, dtype=np.float32)


def lgcp_constants(file_path, num_dim=1600):
    """cp_utils.py:16-84 + model_handler.py:305-346 in float64 numpy (one-off setup, not the hot path)."""
    m = int(round(math.sqrt(num_dim)))
    pts = np.genfromtxt(file_path, delimiter=",")
    counts = np.zeros((m, m))
    for elem in pts * m:
        r, c = int(np.floor(elem[0])), int(np.floor(elem[1]))
        r -= r == m
        c -= c == m
        counts[r, c] += 1
    gi = np.arange(m)
    bv = np.array([[a, b] for a in gi for b in gi], dtype=np.float64)
    gram = 1.91 * np.exp(-np.linalg.norm(bv[:, None] - bv[None], axis=-1) / (m * (1.0 / 33)))
    chol = np.linalg.cholesky(gram)
    linv = np.linalg.inv(chol)
    return dict(counts=counts.reshape(-1), kinv=linv.T @ linv, linv=np.tril(linv), chol=chol,
                mu0=math.log(126.0) - 0.5 * 1.91,
                log_norm=-0.5 * num_dim * math.log(2 * math.pi) - float(np.sum(np.log(np.abs(np.diag(chol))))),
                bin_area=1.0 / num_dim)


def load_model(model="many_gmm", config=None, device="cuda"):
    """model_handler.py:30-43.  Returns (log_prob, dim) (+ None sample_fn for the tractable targets)."""
    g = lambda k, dflt: getattr(config, k, dflt) if config is not None else dflt
    if "funnel" in model:
        return Target("funnel", int(g("funnel_d", 10)), device), int(g("funnel_d", 10)), None
    if "lgcp" in model:
        c = lgcp_constants(g("file_path", os.path.join(_DATA, "pines.csv")))
        dev = torch.device(device)
        if g("use_whitened", False):
            # config.use_whitened (model_handler.py:348-351,373-384; configs/base.py:134 default False): the density of the whitened
            # variable e with latent = L e + mu0 (cp_utils.py:107-128).  Not in the fused registry: served through the batched
            # score callback of the step-wise path with CLOSED-FORM score and Hessian-vector product (two dense products each),
            #   score = -e + L^T (counts - a exp(latent)),   H v = -v - L^T (a exp(latent) * (L v)).
            chol = torch.tensor(c["chol"], dtype=torch.float32, device=dev)
            counts = torch.tensor(c["counts"], dtype=torch.float32, device=dev)
            mu0, area, wn = float(c["mu0"]), float(c["bin_area"]), -0.5 * 1600 * math.log(2.0 * math.pi)

            def log_prob_white(white):
                latent = white @ chol.T + mu0
                return wn - 0.5 * (white * white).sum(-1) + (latent * counts - area * torch.exp(latent)).sum(-1)

            def closed_form(white, v):
                latent = white @ chol.T + mu0
                e = area * torch.exp(latent)
                lp = wn - 0.5 * (white * white).sum(-1) + (latent * counts).sum(-1) - e.sum(-1)
                sc = -white + (counts - e) @ chol
                hv = None if v is None else -v - (e * (v @ chol.T)) @ chol
                return lp, sc, hv

            return CallbackTarget(log_prob_white, 1600, device, closed_form=closed_form), 1600
        lg = {k: torch.tensor(c[k], dtype=torch.float32).contiguous().to(dev) for k in ("kinv", "linv", "counts")}
        lg.update(mu0=c["mu0"], log_norm=c["log_norm"], bin_area=c["bin_area"])
        return Target("lgcp", 1600, device, lgcp=lg), 1600
    if "many_gmm" in model:
        n_mixes, loc = int(g("n_mixes", 40)), float(g("loc_scaling", 40))
        mix = np.zeros((n_mixes, MIX_STRIDE), np.float32)
        mix[:, :2] = many_gmm_means(n_mixes, 2, loc)
        scale = float(np.float32(np.log1p(np.exp(np.float32(0.1)))))  # softplus(0.1) passed as *scale* (:262-267)
        return Target("many_gmm", 2, device, ncomp=n_mixes, scale=scale, invalid_below=-1e4, mix=mix), 2, None
    if "gmm" in model:
        return Target("gmm", 2, device, ncomp=6, mix=_gmm2_components()), 2, None
    raise NotImplementedError(f"target {model!r} is not in the fused-kernel registry (gmm, many_gmm, funnel, lgcp)")
