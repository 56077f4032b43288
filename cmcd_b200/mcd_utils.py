"""Bridge operator dispatch -- host side of the fused CUDA bridge kernels.

Mirrors /root/reference/src/mcd_utils.py:24-190 (``evolve``: dispatch on ``params_fixed[2]``)
for the overdamped modes on the hot path: MCD_ULA / MCD_ULA_sn (src/mcd_over_orig.py),
MCD_CAIS_sn (src/mcd_cais.py), MCD_CAIS_var_sn (src/mcd_cais_var.py), and the underdamped "LDVI" family
MCD_U_a-lp / -sna / -sn (src/mcd_under_lp_a.py), MCD_U_e-lp / -sna (src/mcd_under_lp_e.py), MCD_U_ea-lp-sn
(src/mcd_under_lp_ea.py) -- SURVEY section 8f row 3.  Unknown modes raise
``NotImplementedError("Mode not implemented.")`` like mcd_utils.py:190.

The reference's per-particle ``evolve(z, betas, params, rng_key_gen, ...)`` runs under
``jax.vmap``; the CUDA kernel fuses the whole per-particle program (key chain from the integer
seed, z0 ~ q, K bridge steps, log p(z_K)) -- ``bridge`` below is that fused call as a
``torch.autograd.Function`` (forward = ``cmcd_bridge_fwd``, backward = ``cmcd_bridge_bwd``).
"""
from __future__ import annotations

import math

import torch

from . import _lib
from ._lib import ARCH, MODE, UD_MODES, CmcdBridgeDesc, CmcdNet, CmcdNetGrad
from .nn import build_tables

SUPPORTED_MODES = tuple(MODE)
_NET_KEYS = ("U1", "U2", "U3", "W2", "W3", "c1", "c2", "c3", "out_scale")


def eps_table(eps0, nbridges, eps_schedule=None):
    """Per-step step sizes (mcd_cais.py:34-44,54-59), differentiable in eps0.  Returns [K]."""
    i = torch.arange(nbridges, device=eps0.device, dtype=torch.float32)
    if eps_schedule == "cos_sq":
        phase = i / nbridges
        decay = torch.cos((phase + 0.008) / 1.008 * 0.5 * math.pi) ** 2
        return eps0 * decay
    if eps_schedule == "linear":
        return (0.0001 - eps0) / (nbridges - 1) * i + eps0
    return eps0 * torch.ones_like(i)


def ud_coeff_table(mode, params, nbridges):
    """[7, K] rows (eps, a_f, s_f, a_b, c_n, s_b, c_f) of the underdamped step (csrc/bridge_ud.cu), formed from the reference's
    scalars with differentiable ops so the kernel's row cotangents chain into eps / gamma / eta:
      forward kernel  rho' ~ N(a_f rho + c_f NN, s_f);  backward kernel  rho ~ N(a_b rho' + c_n NN, s_b)."""
    eps, gamma, eta = params["eps"], params["gamma"], params["eta"]
    c_f = torch.zeros((), device=eps.device)
    if mode == "MCD_CAIS_UHA_sn":                                          # mcd_under_lp_a_cais.py:33-40,50-58,79-82
        eps = eps_table(eps, nbridges, "cos_sq")                           # the body hard-codes the cosine schedule
        eta_aux = gamma * eps
        a_f = a_b = 1.0 - eta_aux
        s_f = s_b = torch.sqrt(2.0 * eta_aux)
        c_n, c_f = 2.0 * eta_aux, -2.0 * eta_aux
        return torch.stack([eps, a_f, s_f, a_b, c_n, s_b, c_f])
    if mode in ("MCD_U_a-lp", "MCD_U_a-lp-sna", "MCD_U_a-lp-sn"):        # mcd_under_lp_a.py:28-51
        eta_aux = gamma * eps
        a_f = a_b = 1.0 - eta_aux
        s_f = s_b = torch.sqrt(2.0 * eta_aux)
        c_n = 2 * eta_aux
    elif mode in ("MCD_U_e-lp", "MCD_U_e-lp-sna"):                       # mcd_under_lp_e.py:27-43
        a_f = a_b = eta * torch.ones_like(eps)
        s_f = s_b = torch.sqrt(1.0 - eta ** 2)
        c_n = 2 * (1.0 - eta)
    elif mode == "MCD_U_ea-lp-sn":                                       # mcd_under_lp_ea.py:28-57
        eta_aux = gamma * eps
        a_f = torch.exp(-gamma * eps)
        s_f = torch.sqrt(1.0 - a_f ** 2)
        a_b = 1.0 - eta_aux
        c_n = 2 * eta_aux
        s_b = torch.sqrt(2.0 * eta_aux)
    else:
        raise NotImplementedError("Mode not implemented.")
    one = torch.ones(nbridges, device=eps.device, dtype=torch.float32)
    return torch.stack([v * one for v in (eps, a_f, s_f, a_b, c_n, s_b, c_f)])


def _clips(mode, grad_clipping):
    """grad_clipping -> (clip_target, clip_q).  mcd_cais.py:24-30 (1e3, target only);
    mcd_cais_var.py:33-40 (1e2, both); mcd_over_orig.py never clips."""
    inf = float("inf")
    if mode == "MCD_CAIS_UHA_sn":   # mcd_under_lp_a_cais.py:23-30,48: stable=True -> the target score is always clipped at 1e2
        return 1e2, inf
    if not grad_clipping or mode in ("MCD_ULA", "MCD_ULA_sn") or mode in UD_MODES:   # mcd_under_lp_a.py takes no grad_clipping
        return inf, inf
    return (1e2, 1e2) if mode == "MCD_CAIS_var_sn" else (1e3, inf)


def _make_desc(mode, dim, nbridges, n, clip_t, clip_q, lfsteps=0):
    d = CmcdBridgeDesc()
    d.lfsteps = lfsteps
    d.mode, d.dim, d.nbridges, d.n_particles = MODE[mode], dim, nbridges, n
    d.clip_target, d.clip_q = clip_t, clip_q
    return d


def _make_net(apply_fun, tabs, nbridges):
    net = CmcdNet()
    if apply_fun is None or tabs is None:
        net.arch = 0
        return net
    net.arch, net.hidden, net.hidden_pad, net.n_rows = ARCH[apply_fun.arch], apply_fun.hidden, apply_fun.hidden_pad, nbridges + 1
    for k in _NET_KEYS[:-1]:
        setattr(net, k, _lib.ptr(tabs[k]))
    net.out_scale = 1.0
    net.out_scale_dev = _lib.ptr(tabs["out_scale"]) if apply_fun.arch == "geffner" else None   # factor_sn stays on the device
    net.out_clip = 1.0e4 if apply_fun.arch == "dds" else float("inf")
    return net


def _check_cb(rc, target):
    """_lib.check, but a Python exception raised inside a callback target's log density is re-raised as itself."""
    err = getattr(target, "error", None)
    if rc != 0 and err is not None:
        target.error = None
        raise err
    _lib.check(rc)


def _launch_fwd(cfg, seeds, vd_mean, vd_logdiag, betas, eps, tabs):
    """cmcd_bridge_fwd on detached, contiguous float32 device buffers -> (negw, z, traj)."""
    mode, dim, K, apply_fun, target, clip_t, clip_q, need_grad = cfg[:8]
    lfsteps = cfg[8] if len(cfg) > 8 else 0
    dev = vd_mean.device
    n = seeds.numel()
    negw = torch.empty(n, device=dev, dtype=torch.float32)
    z = torch.empty(n, dim, device=dev, dtype=torch.float32)
    rows = 3 * dim if (mode in UD_MODES or mode == "UHA") else dim   # underdamped: (z_j, rho_j, rho'_j) per node
    traj = torch.empty((K + 1, rows, n), device=dev, dtype=torch.float32) if need_grad else None
    desc, net, tg = _make_desc(mode, dim, K, n, clip_t, clip_q, lfsteps), _make_net(apply_fun, tabs, K), target.desc()
    L = _lib.lib()
    ws_bytes = L.cmcd_bridge_fwd_workspace_bytes(desc, net, tg)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if ws_bytes else None
    with _lib.timed("fwd"):
        rc = L.cmcd_bridge_fwd(desc, _lib.current_stream(), _lib.ptr(seeds), _lib.ptr(vd_mean),
                               _lib.ptr(vd_logdiag), _lib.ptr(betas), _lib.ptr(eps), net, tg,
                               _lib.ptr(negw), _lib.ptr(z), _lib.ptr(traj), _lib.ptr(ws), ws_bytes)
        _check_cb(rc, target)
    _lib.count_launches(1 if ws is None else 3 + 16 * K)
    return negw, z, traj


def _launch_bwd(cfg, seeds, vd_mean, vd_logdiag, betas, eps, tabs, traj, cot_negw):
    """cmcd_bridge_bwd -> (g_mean, g_logdiag, g_betas, g_eps, {table name: cotangent})."""
    mode, dim, K, apply_fun, target, clip_t, clip_q, need_grad = cfg[:8]
    lfsteps = cfg[8] if len(cfg) > 8 else 0
    if traj is None:
        raise RuntimeError("bridge was run without a trajectory; cannot differentiate")
    dev = vd_mean.device
    n = seeds.numel()
    cot = cot_negw.detach().to(torch.float32).contiguous()
    wide_rows = mode in UD_MODES or mode == "UHA"   # eps carries several coefficient rows
    desc, net, tg = _make_desc(mode, dim, K, n, clip_t, clip_q, lfsteps), _make_net(apply_fun, tabs, K), target.desc()
    # every cotangent buffer is a view of ONE zero-filled allocation (one fill launch instead of a dozen)
    shapes = [("mean", vd_mean.shape), ("logdiag", vd_logdiag.shape), ("betas", (max(K, 1),)),
              ("eps", tuple(eps.shape) if wide_rows else (max(K, 1),))]
    if apply_fun is not None:
        shapes += [(k, tuple(tabs[k].shape)) for k in _NET_KEYS if tabs[k] is not None]
    sizes = [(math.prod(sh) + 3) // 4 * 4 for _, sh in shapes]          # 16-byte aligned views
    pool = torch.zeros(sum(sizes), device=dev, dtype=torch.float32)
    views, o = {}, 0
    for (name, sh), sz in zip(shapes, sizes):
        views[name] = pool[o:o + math.prod(sh)].view(sh)
        o += sz
    g_mean, g_logdiag, g_betas, g_eps = views["mean"], views["logdiag"], views["betas"], views["eps"]
    gnet, gt = CmcdNetGrad(), {}
    if apply_fun is not None:
        for k in _NET_KEYS:
            if tabs[k] is not None:
                gt[k] = views[k]
                setattr(gnet, k, _lib.ptr(gt[k]))
    L = _lib.lib()
    ws_bytes = L.cmcd_bridge_bwd_workspace_bytes_for_target(desc, net, tg)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
    with _lib.timed("bwd"):
        rc = L.cmcd_bridge_bwd(desc, _lib.current_stream(), _lib.ptr(seeds), _lib.ptr(vd_mean),
                               _lib.ptr(vd_logdiag), _lib.ptr(betas), _lib.ptr(eps), net, tg, _lib.ptr(traj),
                               _lib.ptr(cot), _lib.ptr(g_mean), _lib.ptr(g_logdiag), _lib.ptr(g_betas),
                               _lib.ptr(g_eps), gnet, _lib.ptr(ws), ws_bytes)
        _check_cb(rc, target)
    _lib.count_launches(2)  # adjoint kernel + partial-gradient reduce kernel
    return g_mean, g_logdiag, g_betas, g_eps, gt, gnet


def _save(ctx, seeds, vd_mean, vd_logdiag, betas, eps, tabs, traj, *extra):
    """Forward-time tensors go through ctx.save_for_backward (freed with the graph, visible to saved-tensor hooks); only the
    table names stay on ctx."""
    ctx.tab_keys = list(tabs) if tabs is not None else None
    ctx.n_extra = len(extra)
    ctx.save_for_backward(seeds, vd_mean, vd_logdiag, betas, eps, traj, *extra, *((tabs[k] for k in ctx.tab_keys) if tabs is not None else ()))


def _saved(ctx):
    """-> (seeds, vd_mean, vd_logdiag, betas, eps, tabs, traj) [+ extras via ctx.saved_tensors[6:6 + n_extra]]."""
    t = ctx.saved_tensors
    seeds, vd_mean, vd_logdiag, betas, eps, traj = t[:6]
    rest = t[6 + ctx.n_extra:]
    tabs = dict(zip(ctx.tab_keys, rest)) if ctx.tab_keys is not None else None
    return seeds, vd_mean, vd_logdiag, betas, eps, tabs, traj


class _Bridge(torch.autograd.Function):
    """(-w[N], z_K[N,d]) = bridge(seeds; vd, betas, eps, net tables)."""

    @staticmethod
    def forward(ctx, cfg, seeds, vd_mean, vd_logdiag, betas, eps, *net_t):
        apply_fun = cfg[3]
        _lib.require_cuda(seeds, vd_mean)
        f32 = lambda t: None if t is None else t.detach().to(torch.float32).contiguous()
        vd_mean, vd_logdiag, betas, eps = f32(vd_mean), f32(vd_logdiag), f32(betas), f32(eps)
        tabs = {k: f32(t) for k, t in zip(_NET_KEYS, net_t)} if apply_fun is not None else None
        negw, z, traj = _launch_fwd(cfg, seeds, vd_mean, vd_logdiag, betas, eps, tabs)
        ctx.cfg = cfg
        _save(ctx, seeds, vd_mean, vd_logdiag, betas, eps, tabs, traj)
        ctx.mark_non_differentiable(z)
        return negw, z

    @staticmethod
    def backward(ctx, cot_negw, _cot_z):
        mode, dim, K, apply_fun = ctx.cfg[:4]
        g_mean, g_logdiag, g_betas, g_eps, gt, _ = _launch_bwd(ctx.cfg, *_saved(ctx), cot_negw)
        net_grads = tuple(gt.get(k) for k in _NET_KEYS) if apply_fun is not None else ()
        if mode in UD_MODES or mode == "UHA":
            return (None, None, g_mean, g_logdiag, g_betas[:K] if K else None, g_eps if K else None, *net_grads)
        return (None, None, g_mean, g_logdiag, g_betas[:K] if K else None, g_eps[:K] if K else None, *net_grads)


# ------------------------------------------------------------------ fused O(K) chain (csrc/chain.cu)
_CHAIN_MODES = ("MCD_ULA", "MCD_ULA_sn", "MCD_CAIS_sn", "MCD_CAIS_var_sn")
_CHAIN_CACHE = {}


def _chain_desc(unflatten, params_fixed, eps_schedule, device):
    """cmcd_chain for this (pytree layout, problem) pair, or None when the fused chain does not serve it.  Cached: the
    descriptor depends only on static data (the reference passes the same objects as static jit arguments)."""
    import os
    dim, nbridges, mode, apply_fun = params_fixed
    if os.environ.get("CMCD_DISABLE_CHAIN") or mode not in _CHAIN_MODES or nbridges < 1 or not hasattr(unflatten, "leaf_offsets"):
        return None
    key = (id(unflatten), params_fixed, eps_schedule, str(device))
    if key in _CHAIN_CACHE:
        return _CHAIN_CACHE[key]
    from ._lib import EPS_SCHEDULE, LEAF, CmcdChain
    from .nn import dds_timestep_coeff
    c = None
    try:
        pt, pn = unflatten.leaf_offsets()
        uses_net = mode != "MCD_ULA"
        arch = apply_fun.arch if (uses_net and apply_fun is not None) else None
        if uses_net and (apply_fun is None or apply_fun.rho_dim != 0 or "sn" not in pt):
            raise KeyError("network")
        c = CmcdChain()
        c.arch, c.dim, c.in_dim, c.nbridges = ARCH[arch], dim, dim, nbridges
        c.emb_dim = apply_fun.emb_dim if arch == "geffner" else 0
        c.hidden = apply_fun.hidden if arch else 0
        c.hidden_pad = apply_fun.hidden_pad if arch else 0
        c.eps_schedule = EPS_SCHEDULE[eps_schedule if mode in ("MCD_CAIS_sn", "MCD_CAIS_var_sn") else None]   # orig ignores the schedule
        c.n_params = unflatten.size
        for i in range(len(c.off)):
            c.off[i] = -1
        mask = 0

        def put(name, tree_train, tree_fixed, *path):
            nonlocal mask
            for tree, trained in ((tree_train, True), (tree_fixed, False)):
                node = tree
                try:
                    for k in path:
                        node = node[k]
                except (KeyError, IndexError, TypeError):
                    continue
                c.off[LEAF[name]] = node[0]
                if trained:
                    mask |= 1 << LEAF[name]
                return node[1]
            raise KeyError(path)

        put("VD_MEAN", pt, pn, "vd", "mean"); put("VD_LOGDIAG", pt, pn, "vd", "logdiag"); put("EPS", pt, pn, "eps")
        c.ngrid = put("MGRID_Y", pt, pn, "mgridref_y")[0]
        put("GRID_X", pt, pn, "gridref_x"); put("TARGET_X", pt, pn, "target_x")
        if c.ngrid > 39:
            raise KeyError("ngrid")
        if arch == "dds":
            put("DDS_PHASE", pt, pn, "sn", "timestep_phase")
            for nm, k in (("TC1", "tc1"), ("TC2", "tc2"), ("ST1", "st1"), ("ST2", "st2"), ("OUT", "out")):
                put(f"DDS_{nm}_W", pt, pn, "sn", k, "w"); put(f"DDS_{nm}_B", pt, pn, "sn", k, "b")
            c.dds_coeff = _lib.ptr(dds_timestep_coeff(device))
        elif arch == "geffner":
            put("GEF_EMB", pt, pn, "sn", "emb"); put("GEF_FACTOR", pt, pn, "sn", "factor_sn")
            for l in range(3):
                put(f"GEF_W{l + 1}", pt, pn, "sn", "nn", l, "w"); put(f"GEF_B{l + 1}", pt, pn, "sn", "nn", l, "b")
        c.train_mask = mask
    except KeyError:
        c = None
    _CHAIN_CACHE[key] = c
    return c


def chain_supported(params_flat, unflatten, params_fixed, eps_schedule=None):
    return (isinstance(params_flat, torch.Tensor) and params_flat.is_cuda and params_flat.dtype == torch.float32
            and _chain_desc(unflatten, params_fixed, eps_schedule, params_flat.device) is not None)


def _chain_forward(chain, p, apply_fun, dim, K):
    """cmcd_chain_fwd on the (detached, contiguous) flat vector -> (vd_mean, vd_logdiag, betas, eps, tables): betas / eps and the
    per-step tables are fresh buffers, everything else is a view of ``p``."""
    from ._lib import LEAF
    dev = p.device
    T, hp = K + 1, chain.hidden_pad
    betas, eps = torch.empty(K, device=dev), torch.empty(K, device=dev)
    tabs, pads = None, [None] * 4
    view = lambda leaf, *shape: p[chain.off[leaf]:chain.off[leaf] + math.prod(shape)].view(*shape)
    c1 = c2 = c3 = None
    if apply_fun is not None:
        c1, c2, c3 = torch.empty(T, hp, device=dev), torch.empty(T, hp, device=dev), torch.empty(T, dim, device=dev)
        if apply_fun.arch == "dds":
            tabs = {"U1": view(LEAF["DDS_ST1_W"], dim, hp), "U2": None, "U3": None, "W2": view(LEAF["DDS_ST2_W"], hp, hp),
                    "W3": view(LEAF["DDS_OUT_W"], hp, dim), "out_scale": None}
        else:
            pads = [torch.empty(dim, hp, device=dev), torch.empty(dim, hp, device=dev), torch.empty(hp, hp, device=dev),
                    torch.empty(hp, dim, device=dev)]
            tabs = {"U1": pads[0], "U2": pads[1], "U3": view(LEAF["GEF_W3"], dim, dim), "W2": pads[2], "W3": pads[3],
                    "out_scale": view(LEAF["GEF_FACTOR"], 1)}
        tabs.update(c1=c1, c2=c2, c3=c3)
    _lib.check(_lib.lib().cmcd_chain_fwd(chain, _lib.current_stream(), _lib.ptr(p), _lib.ptr(betas), _lib.ptr(eps), _lib.ptr(c1),
                                         _lib.ptr(c2), _lib.ptr(c3), *[_lib.ptr(t) for t in pads]))
    _lib.count_launches(1)
    return view(LEAF["VD_MEAN"], dim), view(LEAF["VD_LOGDIAG"], dim), betas, eps, tabs


class _FusedBridge(torch.autograd.Function):
    """(-w[N], z_K[N,d]) = bridge(seeds; params_flat) with the O(K) chain fused: forward = cmcd_chain_fwd + cmcd_bridge_fwd,
    backward = cmcd_bridge_bwd + cmcd_chain_bwd (flat gradient written directly)."""

    @staticmethod
    def forward(ctx, cfg, chain, seeds, params_flat):
        mode, dim, K, apply_fun = cfg[:4]
        _lib.require_cuda(seeds, params_flat)
        p = params_flat.detach().contiguous()
        vd_mean, vd_logdiag, betas, eps, tabs = _chain_forward(chain, p, apply_fun, dim, K)
        negw, z, traj = _launch_fwd(cfg, seeds, vd_mean, vd_logdiag, betas, eps, tabs)
        ctx.cfg, ctx.chain = cfg, chain
        _save(ctx, seeds, vd_mean, vd_logdiag, betas, eps, tabs, traj, p)
        ctx.mark_non_differentiable(z)
        return negw, z

    @staticmethod
    def backward(ctx, cot_negw, _cot_z):
        chain, p = ctx.chain, ctx.saved_tensors[6]
        g_mean, g_logdiag, g_betas, g_eps, gt, gnet = _launch_bwd(ctx.cfg, *_saved(ctx), cot_negw)
        L = _lib.lib()
        n_scratch = L.cmcd_chain_bwd_scratch_floats(chain)
        scratch = torch.empty(n_scratch, device=p.device)
        grad = torch.empty_like(p)
        _lib.check(L.cmcd_chain_bwd(chain, _lib.current_stream(), _lib.ptr(p), _lib.ptr(g_betas), _lib.ptr(g_eps), _lib.ptr(g_mean),
                                    _lib.ptr(g_logdiag), gnet, _lib.ptr(scratch), n_scratch, _lib.ptr(grad)))
        _lib.count_launches(2)
        return None, None, None, grad


def fused_bridge(seeds, params_flat, unflatten, params_fixed, log_prob_model, eps_schedule=None, grad_clipping=False):
    """``bridge`` for the overdamped modes straight from the flat parameter vector: betas, the step-size schedule and the
    per-step network tables are formed by ONE prologue kernel and their transposes by TWO epilogue kernels (csrc/chain.cu)
    instead of the ~100 framework launches of ``make_betas`` / ``eps_table`` / ``nn.build_tables`` and their autograd mirror.
    Same values up to the summation order of the 64-wide time-coder products.  ``CMCD_DISABLE_CHAIN=1`` keeps the framework
    chain (A/B runs, tests)."""
    dim, nbridges, mode, apply_fun = params_fixed
    chain = _chain_desc(unflatten, params_fixed, eps_schedule, params_flat.device)
    uses_net = mode != "MCD_ULA"
    clip_t, clip_q = _clips(mode, grad_clipping)
    need_grad = torch.is_grad_enabled() and params_flat.requires_grad
    cfg = (mode, dim, nbridges, apply_fun if uses_net else None, log_prob_model, clip_t, clip_q, need_grad)
    seeds = torch.as_tensor(seeds, dtype=torch.int32, device=params_flat.device).contiguous()
    return _FusedBridge.apply(cfg, chain, seeds, params_flat)


def bridge(seeds, params, betas, params_fixed, log_prob_model, eps_schedule=None, grad_clipping=False):
    """Fused per-particle program of compute_log_elbo (mcdboundingmachine.py:126-179).

    Returns (-w[N], z_K[N,d]); differentiable w.r.t. everything in ``params`` / ``betas``."""
    dim, nbridges, mode, apply_fun = params_fixed
    if mode not in MODE:
        raise NotImplementedError("Mode not implemented.")
    vd = params["vd"]
    dev = vd["mean"].device
    uses_net = mode not in ("MCD_ULA", "MCD_U_a-lp", "MCD_U_e-lp") and nbridges >= 1
    if uses_net and apply_fun is None:
        raise RuntimeError(f"mode {mode} needs a score network")
    clip_t, clip_q = _clips(mode, grad_clipping)
    if nbridges >= 1 and mode in UD_MODES:
        eps = ud_coeff_table(mode, params, nbridges)   # [7, K]: eps and the kernel coefficients, constant over the steps
    elif nbridges >= 1:
        sched = eps_schedule if mode in ("MCD_CAIS_sn", "MCD_CAIS_var_sn") else None  # orig ignores the schedule
        eps = eps_table(params["eps"], nbridges, sched)
    else:
        betas = torch.zeros(1, device=dev)
        eps = torch.zeros(1, device=dev)
    net_t = ()
    if uses_net:
        tabs = build_tables(apply_fun, params["sn"])
        net_t = tuple(tabs[k] for k in _NET_KEYS)
    need_grad = torch.is_grad_enabled() and any(
        t is not None and t.requires_grad for t in (vd["mean"], vd["logdiag"], betas, eps, *net_t))
    cfg = (mode, dim, nbridges, apply_fun if uses_net else None, log_prob_model, clip_t, clip_q, need_grad)
    seeds = torch.as_tensor(seeds, dtype=torch.int32, device=dev).contiguous()
    return _Bridge.apply(cfg, seeds, vd["mean"], vd["logdiag"], betas, eps, *net_t)


def uha_bridge(seeds, params, betas, params_fixed, log_prob_model):
    """Fused per-particle program of boundingmachine.compute_log_elbo (boundingmachine.py:73-104) for nbridges >= 1 =
    ais_utils.evolve (ais_utils.py:7-69).  Returns (-w[N], z_K[N,d]); differentiable w.r.t. vd, eps, eta, md, betas."""
    dim, nbridges, lfsteps = params_fixed
    vd = params["vd"]
    dev = vd["mean"].device
    eps, eta = params["eps"], params["eta"]
    one = torch.ones(nbridges, device=dev, dtype=torch.float32)
    rows = torch.stack([eps * one, eta * one, torch.sqrt(1.0 - eta ** 2) * one])   # momdist.py:21
    scales = torch.stack([vd["logdiag"], params["md"]])                            # [2, d]: q log-scales, momentum log-scales
    need_grad = torch.is_grad_enabled() and any(t.requires_grad for t in (vd["mean"], scales, betas, rows))
    cfg = ("UHA", dim, nbridges, None, log_prob_model, float("inf"), float("inf"), need_grad, int(lfsteps))
    seeds = torch.as_tensor(seeds, dtype=torch.int32, device=dev).contiguous()
    return _Bridge.apply(cfg, seeds, vd["mean"], scales, betas, rows)


def _forward_inputs(params, betas, params_fixed, eps_schedule, grad_clipping):
    """(vd_mean, vd_logdiag, betas, eps rows, tables, clip levels) of a forward-only call, detached and contiguous."""
    dim, nbridges, mode, apply_fun = params_fixed
    if mode not in MODE:
        raise NotImplementedError("Mode not implemented.")
    uses_net = mode not in ("MCD_ULA", "MCD_U_a-lp", "MCD_U_e-lp") and nbridges >= 1
    if uses_net and apply_fun is None:
        raise RuntimeError(f"mode {mode} needs a score network")
    f32 = lambda t: t.detach().to(torch.float32).contiguous()
    sched = eps_schedule if mode in ("MCD_CAIS_sn", "MCD_CAIS_var_sn") else None
    with torch.no_grad():
        eps = f32(eps_table(params["eps"], nbridges, sched))
        tabs = {k: (None if t is None else f32(t)) for k, t in build_tables(apply_fun, params["sn"]).items()} if uses_net else None
    return f32(params["vd"]["mean"]), f32(params["vd"]["logdiag"]), f32(betas), eps, tabs, _clips(mode, grad_clipping), uses_net


def evolve(z, betas, params, rng_key_gen, params_fixed, log_prob_model, eps_schedule=None, grad_clipping=False):
    """mcd_utils.py:24-33: ``(z, w, None) = evolve(z, betas, params, rng_key_gen, ...)`` -- the K bridge steps from a given state and
    PRNG key, dispatching on ``params_fixed[2]`` (mcd_cais.py:6-99, mcd_cais_var.py:7-112, mcd_over_orig.py:6-65).  The reference
    calls it per particle under ``jax.vmap``; here ``z`` is ``[N, d]`` (or ``[d]``) and ``rng_key_gen`` the matching ``[N, 2]``
    (or ``[2]``) uint32 keys, and the whole batch runs in one launch of ``cmcd_bridge_evolve``.  Forward only: the differentiable
    entry is ``bridge`` / ``mcdboundingmachine.compute_log_elbo`` (whose kernels fuse the key chain, z0 ~ q and log p(z_K) around
    these same steps).  Overdamped modes; unknown modes raise ``NotImplementedError("Mode not implemented.")`` (mcd_utils.py:190)."""
    dim, nbridges, mode, apply_fun = params_fixed
    if mode not in MODE:
        raise NotImplementedError("Mode not implemented.")
    if mode in UD_MODES or mode == "UHA":
        raise NotImplementedError(f"Mode not implemented. (evolve() from a given (z, key) serves the overdamped modes; {mode} runs "
                                  "through mcd_utils.bridge / boundingmachine.compute_log_elbo)")
    single = z.dim() == 1
    zz = (z[None] if single else z).detach().to(torch.float32).contiguous()
    _lib.require_cuda(zz)
    dev = zz.device
    import numpy as np
    keys = rng_key_gen.cpu().numpy() if isinstance(rng_key_gen, torch.Tensor) else np.asarray(rng_key_gen)
    keys = np.ascontiguousarray(keys.astype(np.uint32).reshape(-1, 2))
    n = zz.shape[0]
    if keys.shape[0] != n:
        raise ValueError(f"evolve: {n} states but {keys.shape[0]} keys")
    keys_dev = torch.from_numpy(keys.view(np.int32)).to(dev)        # bit pattern travels as int32 (torch has no uint32 arithmetic)
    vd_mean, vd_logdiag, betas_c, eps, tabs, (clip_t, clip_q), uses_net = _forward_inputs(params, betas, params_fixed, eps_schedule,
                                                                                         grad_clipping)
    out_z, out_w = torch.empty(n, dim, device=dev), torch.empty(n, device=dev)
    desc, net = _make_desc(mode, dim, nbridges, n, clip_t, clip_q), _make_net(apply_fun if uses_net else None, tabs, nbridges)
    _lib.check(_lib.lib().cmcd_bridge_evolve(desc, _lib.current_stream(), _lib.ptr(zz), _lib.ptr(keys_dev), _lib.ptr(vd_mean),
                                             _lib.ptr(vd_logdiag), _lib.ptr(betas_c), _lib.ptr(eps), net, log_prob_model.desc(),
                                             _lib.ptr(out_z), _lib.ptr(out_w)))
    _lib.count_launches(1)
    return (out_z[0], out_w[0], None) if single else (out_z, out_w, None)


_HOST_SCRATCH = {}


def sample_host(seeds_host, params_flat, unflatten, params_fixed, log_prob_model, out_negw_host, out_z_host=None,
                eps_schedule=None, grad_clipping=False):
    """Sampling pass with HOST buffers through ``cmcd_bridge_fwd_host``: ``seeds_host`` int32[N] (pinned for an asynchronous
    copy) in, per-particle losses ``out_negw_host`` f32[N] (and optionally ``out_z_host`` f32[N, d]) back on the host; the call
    returns after the device-to-host copy completed.  Parameters stay resident on the device.  Small-d registry targets."""
    from .mcdboundingmachine import make_betas
    pt, pn = unflatten(params_flat)
    params = {**pt, **pn}
    dim, nbridges, mode, apply_fun = params_fixed
    dev = params_flat.device
    _lib.require_cuda(params_flat)
    if seeds_host.is_cuda or out_negw_host.is_cuda or seeds_host.dtype != torch.int32 or out_negw_host.dtype != torch.float32:
        raise ValueError("sample_host takes host int32 seeds and a host float32 output buffer")
    n = seeds_host.numel()
    chain = _chain_desc(unflatten, params_fixed, eps_schedule, dev) if params_flat.dtype == torch.float32 else None
    if chain is not None:       # same prologue as compute_log_elbo: bit-identical results
        uses_net = mode != "MCD_ULA"
        clip_t, clip_q = _clips(mode, grad_clipping)
        vd_mean, vd_logdiag, betas_c, eps, tabs = _chain_forward(chain, params_flat.detach().contiguous(), apply_fun if uses_net else None,
                                                                 dim, nbridges)
    else:
        with torch.no_grad():
            betas = make_betas(params)
        vd_mean, vd_logdiag, betas_c, eps, tabs, (clip_t, clip_q), uses_net = _forward_inputs(params, betas, params_fixed, eps_schedule,
                                                                                             grad_clipping)
    key = (str(dev), n, dim)
    if key not in _HOST_SCRATCH:
        _HOST_SCRATCH.clear()
        _HOST_SCRATCH[key] = (torch.empty(n, dtype=torch.int32, device=dev), torch.empty(n, device=dev), torch.empty(n, dim, device=dev))
    seeds_dev, negw_dev, z_dev = _HOST_SCRATCH[key]
    desc, net = _make_desc(mode, dim, nbridges, n, clip_t, clip_q), _make_net(apply_fun if uses_net else None, tabs, nbridges)
    _lib.check(_lib.lib().cmcd_bridge_fwd_host(desc, _lib.current_stream(), seeds_host.data_ptr(), _lib.ptr(vd_mean), _lib.ptr(vd_logdiag),
                                               _lib.ptr(betas_c), _lib.ptr(eps), net, log_prob_model.desc(), _lib.ptr(seeds_dev),
                                               _lib.ptr(negw_dev), _lib.ptr(z_dev), out_negw_host.data_ptr(),
                                               None if out_z_host is None else out_z_host.data_ptr()))
    _lib.count_launches(1)
    return out_negw_host
