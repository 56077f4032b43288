"""Generate the committed golden fixtures under tests/golden/.

Two kinds of fixture:

* ``prng_kat.json`` -- PUBLISHED known-answer vectors that pin the PRNG restatement to the outside world: the
  Random123 Threefry-2x32-20 KATs (Salmon et al.) and the values printed in JAX's public documentation for
  ``split(PRNGKey(0))`` / ``normal(PRNGKey(k), shape)``.  These are constants, not generated.
* ``bridge_<config>.npz`` -- per-config input/output vectors of the bridge path (seeds, flat parameter vector ->
  per-particle loss, final state, flat gradient) computed with the float64 ORACLE (oracle/).  The reference itself
  (pure JAX) cannot run in this image (no jax/jaxlib), so these are restatement-derived
  regression pins ("parity unpinned" against a running reference): they freeze the oracle across rounds and let
  the GPU tests check the CUDA path without executing the oracle.

Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import UHA_CONFIGS, oracle_problem, seeds_for, uha_oracle_problem  # noqa: E402
from oracle import mcdboundingmachine as OM  # noqa: E402

# name -> particles in the fixture (kept small: the whole directory stays well under 1 MB)
GOLDEN_CONFIGS = {"A_gmm": 64, "B_funnel": 64, "C_manygmm_dds_small": 96, "Cvar_manygmm": 48, "ULA_gmm": 64,
                  "ULAsn_gmm_dds": 64, "lin_funnel": 48, "LDVI_gmm": 64, "LDVI_funnel_dds": 48, "UDsna_funnel": 48,
                  "UDe_gmm": 64, "UDesna_funnel_dds": 48, "UDea_gmm": 64, "CAISUHA_gmm": 64, "CAISUHA_manygmm_dds": 48}

PRNG_KAT = {
    "source": "Random123 Threefry-2x32-20 known-answer tests; JAX documentation (jax.random.split / normal examples)",
    "threefry2x32": [
        {"key": [0, 0], "ctr": [0, 0], "out": [0x6B200159, 0x99BA4EFE]},
        {"key": [0xFFFFFFFF, 0xFFFFFFFF], "ctr": [0xFFFFFFFF, 0xFFFFFFFF], "out": [0x1CB996FC, 0xBB002BE7]},
        {"key": [0x13198A2E, 0x03707344], "ctr": [0x243F6A88, 0x85A308D3], "out": [0xC4923A9C, 0x483DF7A0]},
    ],
    "split": [{"seed": 0, "out": [[4146024105, 967050713], [2718843009, 1272950319]]}],
    "normal": [
        {"seed": 0, "n": 1, "out": [-0.20584226]},
        {"seed": 42, "n": 1, "out": [-0.18471177]},
        {"seed": 0, "n": 3, "out": [1.8160863, -0.48262316, 0.33988908]},
    ],
}


def main():
    with open(os.path.join(HERE, "prng_kat.json"), "w") as f:
        json.dump(PRNG_KAT, f, indent=1)
    for name, n in GOLDEN_CONFIGS.items():
        c, lp, dim, pf, unf, fixed = oracle_problem(name, torch.float64, N=n)
        seeds = seeds_for(n)
        kw = dict(eps_schedule=c["eps_schedule"], grad_clipping=c["clip"])
        fn = OM.compute_bound_var if "var" in c["mode"] else OM.compute_bound
        g, (l, z) = OM.grad_and_loss(fn, seeds, pf, unf, fixed, lp, **kw)
        with torch.no_grad():
            loss = fn(seeds, pf, unf, fixed, lp, **kw)[0]
        # how far the float32 oracle is from the float64 one, per pytree leaf (max-norm relative): the floor any fp32
        # implementation is judged against (same criterion as tests/test_gpu_parity_bwd.py)
        c32, lp32, _, pf32, unf32, fixed32 = oracle_problem(name, torch.float32, N=n)
        g32, (l32, z32) = OM.grad_and_loss(OM.compute_bound_var if "var" in c["mode"] else OM.compute_bound, seeds, pf32, unf32,
                                           fixed32, lp32, **kw)
        from cmcd_b200.pytree import tree_leaves
        floor = []
        for a_, b_ in zip(tree_leaves(unf32(g32)), tree_leaves(unf(g))):
            a_, b_ = a_.double().reshape(-1), b_.double().reshape(-1)
            if b_.numel() == 0:
                continue
            sc = b_.abs().max().item()
            floor.append((a_ - b_).abs().max().item() / sc if sc > 0 else (a_ - b_).abs().max().item())
        np.savez_compressed(os.path.join(HERE, f"bridge_{name}.npz"), seeds=seeds, params_flat=pf.numpy().astype(np.float32),
                            loss=np.float64(loss.item()), l=l.numpy(), z=z.numpy(), grad=g.numpy().astype(np.float32), grad_fp32_floor=np.array(floor))
        print(f"{name}: N={n} P={pf.numel()} loss={loss.item():.6f}")
    # UHA (boundingmachine.py + ais_utils.py): uha_<config>.npz, same fields
    from cmcd_b200.pytree import tree_leaves
    for name in UHA_CONFIGS:
        n = 64
        c, lp, dim, pf, unf, fixed = uha_oracle_problem(name, torch.float64, N=n)
        seeds = seeds_for(n)
        g, (l, z) = OM.grad_and_loss(OM.uha_compute_bound, seeds, pf, unf, fixed, lp)
        _, lp32, _, pf32, unf32, fixed32 = uha_oracle_problem(name, torch.float32, N=n)
        g32, _ = OM.grad_and_loss(OM.uha_compute_bound, seeds, pf32, unf32, fixed32, lp32)
        floor = []
        for a_, b_ in zip(tree_leaves(unf32(g32)), tree_leaves(unf(g))):
            a_, b_ = a_.double().reshape(-1), b_.double().reshape(-1)
            if b_.numel() == 0:
                continue
            sc = b_.abs().max().item()
            floor.append((a_ - b_).abs().max().item() / sc if sc > 0 else (a_ - b_).abs().max().item())
        np.savez_compressed(os.path.join(HERE, f"uha_{name}.npz"), seeds=seeds, params_flat=pf.numpy().astype(np.float32),
                            loss=np.float64(l.mean().item()), l=l.numpy(), z=z.numpy(), grad=g.numpy().astype(np.float32),
                            grad_fp32_floor=np.array(floor))
        print(f"{name}: N={n} P={pf.numel()} loss={l.mean().item():.6f}")


if __name__ == "__main__":
    main()
