"""The C-ABI library loads and exports every symbol include/cmcd_b200.h declares (no compute without a GPU)."""
import os
import re

import pytest
import torch

from cmcd_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "cmcd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cmcd_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    names = _declared()
    assert len(names) >= 12
    l = _lib.lib()
    for n in names:
        assert hasattr(l, n), f"{n} declared in cmcd_b200.h but not exported"
        assert n in _lib.EXPORTS, f"{n} has no ctypes binding"
    assert set(_lib.EXPORTS) == set(names)
    assert l.cmcd_version() >= 100


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    from cmcd_b200 import mcdboundingmachine as M, model_handler as H
    t, dim, _ = H.load_model("gmm", device="cpu")
    pf, unf, fixed = M.initialize(dim, nbridges=4, trainable=("vd",), mode="MCD_ULA", device="cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        M.compute_bound(torch.arange(1, 5), pf, unf, fixed, t)


def test_unknown_mode_and_target_raise():
    from cmcd_b200 import mcdboundingmachine as M, model_handler as H, mcd_utils
    with pytest.raises(NotImplementedError):
        M.initialize(2, nbridges=4, mode="MCD_DNF", device="cpu")   # not built (broken at the reference HEAD)
    for mode in ("MCD_U_a-lp", "MCD_U_a-lp-sna", "MCD_U_a-lp-sn"):          # built: the evolve_underdamped_lp_a family
        pf, unf, fixed = M.initialize(2, nbridges=4, mode=mode, emb_dim=6, nn_arch="geffner", device="cpu")
        assert fixed[2] == mode and (fixed[3] is None) == (mode == "MCD_U_a-lp")
        if fixed[3] is not None:
            assert fixed[3].in_dim == (4 if mode == "MCD_U_a-lp-sn" else 2)
    with pytest.raises(NotImplementedError):
        H.load_model("lorenz", device="cpu")
    with pytest.raises(NotImplementedError, match="Mode not implemented"):
        mcd_utils.evolve(None, None, None, None, (2, 4, "bogus", None), None)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "cmcd_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh")):
                s = open(os.path.join(dp, f)).read()
                assert "import oracle" not in s and "from oracle" not in s, f


def test_every_kernel_source_is_built():
    """Every .cu / .cc under csrc/ is in the build list (a source left out would silently drop its entry points' kernels)."""
    from cmcd_b200 import build as B
    on_disk = {f for f in os.listdir(B.CSRC) if f.endswith((".cu", ".cc"))}
    assert on_disk == set(B.SOURCES), (sorted(on_disk - set(B.SOURCES)), sorted(set(B.SOURCES) - on_disk))


def test_xla_ffi_shim_compiles_against_stub():
    """cmcd_b200/csrc/xla_ffi.cc (the jax.ffi custom-call handlers) cannot be built against jaxlib here; a header-only stub of
    the public xla/ffi/api/ffi.h surface it uses (tests/xla_ffi_stub) makes the compiler check that every handler's Ctx / Arg /
    Ret / Attr list matches its implementation's parameter list and that the C ABI is called with the right types -- and the
    check really bites: dropping one .Arg<>() from a binding fails the build."""
    import subprocess
    import tempfile
    src = os.path.join(ROOT, "cmcd_b200", "csrc", "xla_ffi.cc")
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "tests", "xla_ffi_stub"), "-I", "/usr/local/cuda/include"]
    ok = subprocess.run(cmd + [src], capture_output=True, text=True)
    assert ok.returncode == 0, ok.stderr[-2000:]
    text = open(src).read()
    marker = ".Arg<F32>()                                                                 // mix"
    assert marker in text
    bad = text.replace(marker, "// mix", 1).replace('"../../include/cmcd_b200.h"', f'"{os.path.join(ROOT, "include", "cmcd_b200.h")}"')
    with tempfile.NamedTemporaryFile("w", suffix=".cc", delete=False) as f:
        f.write(bad)
    try:
        r = subprocess.run(cmd + [f.name], capture_output=True, text=True)
    finally:
        os.unlink(f.name)
    assert r.returncode != 0 and "does not match" in r.stderr
