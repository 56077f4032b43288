"""GPU PRNG parity: the in-kernel threefry / normal chain is bit-exact against the oracle."""
import numpy as np
import pytest
import torch

from cmcd_b200 import _lib
from oracle import prng as P
from helpers import seeds_for

pytestmark = pytest.mark.gpu


def test_threefry_block_bit_exact():
    rng = np.random.default_rng(1)
    n = 100_000
    key = rng.integers(0, 2**32, 2, dtype=np.uint64).astype(np.uint32)
    x0 = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    x1 = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    # KAT rows first
    x0[:2], x1[:2] = [0x243F6A88, 0], [0x85A308D3, 0]
    t = lambda a: torch.from_numpy(a.view(np.int32)).cuda()
    y0, y1 = torch.empty(n, dtype=torch.int32, device="cuda"), torch.empty(n, dtype=torch.int32, device="cuda")
    for k in (key, np.array([0x13198A2E, 0x03707344], np.uint32)):
        kd, x0d, x1d = t(k), t(x0), t(x1)  # keep the device buffers alive across the launch
        _lib.check(_lib.lib().cmcd_threefry2x32(_lib.current_stream(), _lib.ptr(kd), _lib.ptr(x0d), _lib.ptr(x1d),
                                                n, _lib.ptr(y0), _lib.ptr(y1)))
        r0, r1 = P.threefry2x32(k[0], k[1], x0, x1)
        np.testing.assert_array_equal(y0.cpu().numpy().view(np.uint32), r0)
        np.testing.assert_array_equal(y1.cpu().numpy().view(np.uint32), r1)
    assert (int(r0[0]), int(r1[0])) == (0xC4923A9C, 0x483DF7A0)  # Random123 KAT through the GPU path


@pytest.mark.parametrize("dim,K,n", [(2, 16, 4096), (10, 8, 1024), (5, 3, 257), (1600, 2, 8)])
def test_particle_gaussians_bit_exact(dim, K, n):
    seeds = seeds_for(n, seed=3)
    xi0 = torch.empty(n, dim, device="cuda")
    xi = torch.empty(K, n, dim, device="cuda")
    s = torch.from_numpy(seeds).cuda()
    _lib.check(_lib.lib().cmcd_particle_noise(_lib.current_stream(), _lib.ptr(s), n, dim, K, _lib.ptr(xi0), _lib.ptr(xi)))
    r0, r = P.particle_noise(seeds, dim, K)
    np.testing.assert_array_equal(xi0.cpu().numpy().view(np.uint32), r0.view(np.uint32))
    np.testing.assert_array_equal(xi.cpu().numpy().view(np.uint32), r.view(np.uint32))


def test_jax_documented_normals_on_gpu():
    # normal(PRNGKey(0),(3,)) = [1.8160863, -0.48262316, 0.33988908]: reproduce via the particle chain helper
    # (xi0 of a particle is normal(split(PRNGKey(seed))[0], (d,)); check against the oracle's same chain)
    seeds = np.array([0, 42], np.int32)
    xi0 = torch.empty(2, 3, device="cuda")
    xi = torch.empty(1, 2, 3, device="cuda")
    sd = torch.from_numpy(seeds).cuda()
    _lib.check(_lib.lib().cmcd_particle_noise(_lib.current_stream(), _lib.ptr(sd), 2, 3, 1, _lib.ptr(xi0), _lib.ptr(xi)))
    a, _ = P.split(P.prng_key(seeds))
    np.testing.assert_array_equal(xi0.cpu().numpy(), P.normal(a, 3))
