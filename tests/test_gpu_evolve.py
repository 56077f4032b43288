"""The reference's per-particle operator entry ``mcd_utils.evolve(z, betas, params, rng_key_gen, ...)`` (src/mcd_utils.py:24-33)
started from caller-supplied states and PRNG keys (C ABI: cmcd_bridge_evolve), and the host-buffer sampling entry
(cmcd_bridge_fwd_host) -- both against the CPU oracle."""
import numpy as np
import pytest
import torch

from cmcd_b200 import mcd_utils as PU
from cmcd_b200 import mcdboundingmachine as PM
from oracle import mcdboundingmachine as OM
from oracle import prng as P
from helpers import oracle_problem, product_problem, rel_err, seeds_for

pytestmark = pytest.mark.gpu


def _both(name, N, single=False, dtype=torch.float32):
    c, lp, dim, pf, unf, fixed = oracle_problem(name, dtype)
    K = c["K"]
    g = np.random.default_rng(5)
    z0 = (g.normal(size=(N, dim)) * c["sigma"]).astype(np.float32) + np.float32(c.get("vd_mean", 0.0))
    keys = g.integers(0, 2**32, size=(N, 2), dtype=np.uint32)
    pt, pn = unf(pf)
    params = {**pt, **pn}
    xi = P.evolve_noise(keys, dim, K)
    with torch.no_grad():
        z_o, w_o = OM.evolve(torch.from_numpy(z0).to(dtype), OM.make_betas(params), params, torch.from_numpy(xi).to(dtype), fixed, lp,
                             c["eps_schedule"], c["clip"])
    if dtype == torch.float64:
        return z_o, w_o
    _, target, _, pf_p, unf_p, fixed_p = product_problem(name, pf)
    ptp, pnp = unf_p(pf_p)
    pp = {**ptp, **pnp}
    zin = torch.from_numpy(z0).cuda()
    if single:
        z_p, w_p, aux = PU.evolve(zin[0], PM.make_betas(pp), pp, keys[0], fixed_p, target, c["eps_schedule"], c["clip"])
        return (z_o[0], w_o[0]), (z_p.cpu(), w_p.cpu()), aux
    z_p, w_p, aux = PU.evolve(zin, PM.make_betas(pp), pp, keys, fixed_p, target, c["eps_schedule"], c["clip"])
    return (z_o, w_o), (z_p.cpu(), w_p.cpu()), aux


@pytest.mark.parametrize("name", ["A_gmm", "B_funnel", "C_manygmm_dds_small", "Cvar_manygmm", "ULA_gmm", "ULAsn_funnel",
                                  "ULAsn_gmm_dds", "lin_funnel"])
def test_evolve_from_state_and_key(name):
    (z_o, w_o), (z_p, w_p), aux = _both(name, 257)
    z_64, w_64 = _both(name, 257, dtype=torch.float64)      # ground truth; the fp32 oracle's distance to it is the floor
    assert aux is None and z_p.shape == z_o.shape and w_p.shape == w_o.shape
    fin = torch.isfinite(w_64)
    assert (torch.isfinite(w_p) == fin).all()
    assert rel_err(w_p[fin], w_64[fin]).max() < max(1e-4, 2 * rel_err(w_o[fin], w_64[fin]).max())
    assert rel_err(z_p[fin], z_64[fin]).max() < max(1e-4, 2 * rel_err(z_o[fin], z_64[fin]).max())


def test_evolve_single_particle_signature():
    """The reference signature is per particle: z [d], rng_key_gen [2] -> (z [d], w scalar, None)."""
    (z_o, w_o), (z_p, w_p), aux = _both("A_gmm", 3, single=True)
    assert z_p.shape == (2,) and w_p.dim() == 0 and aux is None
    assert rel_err(w_p, w_o).max() < 1e-4 and rel_err(z_p, z_o).max() < 1e-4


def test_evolve_unknown_and_unserved_modes_raise():
    c, lp, dim, pf, unf, fixed = oracle_problem("A_gmm", torch.float32)
    _, target, _, pf_p, unf_p, fixed_p = product_problem("A_gmm", pf)
    pt, pn = unf_p(pf_p)
    pp = {**pt, **pn}
    z, k = torch.zeros(4, 2, device="cuda"), np.zeros((4, 2), np.uint32)
    with pytest.raises(NotImplementedError, match="Mode not implemented"):
        PU.evolve(z, PM.make_betas(pp), pp, k, (fixed_p[0], fixed_p[1], "MCD_DNF", fixed_p[3]), target)
    with pytest.raises(NotImplementedError):
        PU.evolve(z, PM.make_betas(pp), pp, k, (fixed_p[0], fixed_p[1], "MCD_U_a-lp-sn", fixed_p[3]), target)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        PU.evolve(z.cpu(), PM.make_betas(pp), pp, k, fixed_p, target)


@pytest.mark.parametrize("name", ["A_gmm", "C_manygmm_dds_small", "lin_funnel"])
def test_fwd_host_entry_matches_device_entry(name):
    """cmcd_bridge_fwd_host (host seeds in, losses / z back on the host, synchronous) == cmcd_bridge_fwd on device buffers,
    bit for bit, and both match the oracle."""
    c, lp, dim, pf, unf, fixed = oracle_problem(name, torch.float32)
    _, target, _, pf_p, unf_p, fixed_p = product_problem(name, pf)
    seeds = torch.from_numpy(seeds_for(c["N"]))
    kw = dict(eps_schedule=c["eps_schedule"], grad_clipping=c["clip"])
    out_l = torch.empty(c["N"]).pin_memory()
    out_z = torch.empty(c["N"], dim).pin_memory()
    PU.sample_host(seeds.pin_memory(), pf_p, unf_p, fixed_p, target, out_l, out_z, **kw)
    with torch.no_grad():
        l_d, (z_d, _) = PM.compute_log_elbo(seeds, pf_p, unf_p, fixed_p, target, **kw)
        l_o = OM.compute_bound(seeds.numpy(), pf, unf, fixed, lp, **kw)[1][0]
    assert torch.equal(out_l, l_d.cpu()) and torch.equal(out_z, z_d.cpu())
    fin = torch.isfinite(l_o)
    assert rel_err(out_l[fin], l_o[fin]).max() < 1e-4
