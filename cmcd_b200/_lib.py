"""ctypes binding of libcmcd_b200.so (include/cmcd_b200.h).  Fails loudly when the library is absent."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CMCD_B200_LIB") or os.path.join(_HERE, "libcmcd_b200.so")  # env override: A/B builds

# 4-6: the underdamped operators (mcd_utils.py:59-133); the kernel mode only says what the network sees (none / z / (z, rho'))
MODE = {"MCD_ULA": 0, "MCD_ULA_sn": 1, "MCD_CAIS_sn": 2, "MCD_CAIS_var_sn": 3,
        "MCD_U_a-lp": 4, "MCD_U_a-lp-sna": 5, "MCD_U_a-lp-sn": 6,       # evolve_underdamped_lp_a ("LDVI")
        "MCD_U_e-lp": 4, "MCD_U_e-lp-sna": 5,                           # evolve_underdamped_lp_e
        "MCD_U_ea-lp-sn": 6,                                            # evolve_underdamped_lp_ea
        "UHA": 7,                                                       # boundingmachine + ais_utils (config.boundmode "UHA")
        "MCD_CAIS_UHA_sn": 8}                                           # evolve_underdamped_lp_a_cais ("2nd order CMCD")
UD_MODES = ("MCD_U_a-lp", "MCD_U_a-lp-sna", "MCD_U_a-lp-sn", "MCD_U_e-lp", "MCD_U_e-lp-sna", "MCD_U_ea-lp-sn",
            "MCD_CAIS_UHA_sn")
TARGET = {"gmm": 0, "many_gmm": 1, "funnel": 2, "lgcp": 3, "callback": 4}
ARCH = {None: 0, "none": 0, "geffner": 1, "dds": 2}
MIX_STRIDE = 6

_fp = C.c_void_p  # device pointers travel as integers


class CmcdNet(C.Structure):
    _fields_ = [("arch", C.c_int32), ("hidden", C.c_int32), ("hidden_pad", C.c_int32), ("n_rows", C.c_int32),
                ("U1", _fp), ("U2", _fp), ("U3", _fp), ("W2", _fp), ("W3", _fp), ("c1", _fp), ("c2", _fp), ("c3", _fp),
                ("out_scale", C.c_float), ("out_clip", C.c_float), ("out_scale_dev", _fp)]


class CmcdNetGrad(C.Structure):
    _fields_ = [(n, _fp) for n in ("U1", "U2", "U3", "W2", "W3", "c1", "c2", "c3", "out_scale")]


class CmcdTarget(C.Structure):
    _fields_ = [("kind", C.c_int32), ("ncomp", C.c_int32), ("scale", C.c_float), ("invalid_below", C.c_float),
                ("mix", _fp), ("lgcp_kinv", _fp), ("lgcp_linv", _fp), ("lgcp_counts", _fp),
                ("lgcp_mu0", C.c_float), ("lgcp_log_norm", C.c_float), ("lgcp_bin_area", C.c_float),
                ("eval", C.c_void_p), ("user", C.c_void_p)]


# cmcd_target_fn (include/cmcd_b200.h): int f(user, stream, x, n, dim, v, out_logp, out_score, out_hvp)
TARGET_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                        C.c_void_p, C.c_void_p)


class CmcdBridgeDesc(C.Structure):
    _fields_ = [("mode", C.c_int32), ("dim", C.c_int32), ("nbridges", C.c_int32), ("n_particles", C.c_int32),
                ("clip_target", C.c_float), ("clip_q", C.c_float), ("lfsteps", C.c_int32)]


LEAVES = ("VD_MEAN", "VD_LOGDIAG", "EPS", "MGRID_Y", "GRID_X", "TARGET_X",
          "DDS_PHASE", "DDS_TC1_W", "DDS_TC1_B", "DDS_TC2_W", "DDS_TC2_B", "DDS_ST1_W", "DDS_ST1_B", "DDS_ST2_W", "DDS_ST2_B",
          "DDS_OUT_W", "DDS_OUT_B", "GEF_EMB", "GEF_FACTOR", "GEF_W1", "GEF_B1", "GEF_W2", "GEF_B2", "GEF_W3", "GEF_B3")
LEAF = {n: i for i, n in enumerate(LEAVES)}       # CMCD_LEAF_* (include/cmcd_b200.h)
EPS_SCHEDULE = {None: 0, "": 0, "linear": 1, "cos_sq": 2}


class CmcdChain(C.Structure):
    _fields_ = [("arch", C.c_int32), ("dim", C.c_int32), ("in_dim", C.c_int32), ("nbridges", C.c_int32), ("emb_dim", C.c_int32),
                ("hidden", C.c_int32), ("hidden_pad", C.c_int32), ("eps_schedule", C.c_int32), ("ngrid", C.c_int32),
                ("train_mask", C.c_uint32), ("n_params", C.c_int64), ("off", C.c_int64 * len(LEAVES)), ("dds_coeff", _fp)]


EXPORTS = {
    "cmcd_last_error": (C.c_char_p, []),
    "cmcd_version": (C.c_int, []),
    "cmcd_num_sms": (C.c_int, []),
    "cmcd_xla_ffi_available": (C.c_int, []),
    "cmcd_bridge_fwd_workspace_bytes": (C.c_size_t, [C.POINTER(CmcdBridgeDesc), C.POINTER(CmcdNet), C.POINTER(CmcdTarget)]),
    "cmcd_bridge_fwd": (C.c_int, [C.POINTER(CmcdBridgeDesc), _fp, _fp, _fp, _fp, _fp, _fp, C.POINTER(CmcdNet),
                                  C.POINTER(CmcdTarget), _fp, _fp, _fp, _fp, C.c_size_t]),
    "cmcd_bridge_bwd_workspace_bytes": (C.c_size_t, [C.POINTER(CmcdBridgeDesc), C.POINTER(CmcdNet)]),
    "cmcd_bridge_bwd_workspace_bytes_for_target": (C.c_size_t, [C.POINTER(CmcdBridgeDesc), C.POINTER(CmcdNet), C.POINTER(CmcdTarget)]),
    "cmcd_bridge_bwd": (C.c_int, [C.POINTER(CmcdBridgeDesc), _fp, _fp, _fp, _fp, _fp, _fp, C.POINTER(CmcdNet),
                                  C.POINTER(CmcdTarget), _fp, _fp, _fp, _fp, _fp, _fp, C.POINTER(CmcdNetGrad),
                                  _fp, C.c_size_t]),
    "cmcd_loss_stats": (C.c_int, [_fp, _fp, C.c_int64, _fp]),
    "cmcd_batched_elbo_lnz": (C.c_int, [_fp, _fp, C.c_int32, C.c_int32, _fp, _fp]),
    "cmcd_bridge_fwd_host": (C.c_int, [C.POINTER(CmcdBridgeDesc), _fp, _fp, _fp, _fp, _fp, _fp, C.POINTER(CmcdNet),
                                       C.POINTER(CmcdTarget), _fp, _fp, _fp, _fp, _fp]),
    "cmcd_bridge_evolve": (C.c_int, [C.POINTER(CmcdBridgeDesc), _fp, _fp, _fp, _fp, _fp, _fp, _fp, C.POINTER(CmcdNet),
                                     C.POINTER(CmcdTarget), _fp, _fp]),
    "cmcd_chain_fwd": (C.c_int, [C.POINTER(CmcdChain), _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp]),
    "cmcd_chain_bwd_scratch_floats": (C.c_size_t, [C.POINTER(CmcdChain)]),
    "cmcd_chain_bwd": (C.c_int, [C.POINTER(CmcdChain), _fp, _fp, _fp, _fp, _fp, _fp, C.POINTER(CmcdNetGrad), _fp, C.c_size_t, _fp]),
    "cmcd_target_eval": (C.c_int, [C.POINTER(CmcdTarget), C.c_int32, _fp, _fp, C.c_int64, _fp, _fp, _fp, _fp]),
    "cmcd_adam_project_step": (C.c_int, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float,
                                         C.c_float, C.c_int32, _fp, C.c_float, _fp]),
    "cmcd_randint": (C.c_int, [_fp, C.c_uint32, C.c_uint32, C.c_int64, C.c_int32, C.c_int32, _fp]),
    "cmcd_ffma_probe": (C.c_int, [_fp, _fp, C.c_int32, C.c_int32]),
    "cmcd_threefry2x32": (C.c_int, [_fp, _fp, _fp, _fp, C.c_int64, _fp, _fp]),
    "cmcd_particle_noise": (C.c_int, [_fp, _fp, C.c_int64, C.c_int32, C.c_int32, _fp, _fp]),
}

_lib = None

# bookkeeping for bench.py: kernels launched by this library and optional CUDA-event timing of the two bridge launches
LAUNCHES = {"count": 0}
TIMING = {"enabled": False, "fwd": [], "bwd": []}


def count_launches(n):
    LAUNCHES["count"] += n


def timed(kind):
    """Context manager recording CUDA events around one bridge launch on the current stream (no sync)."""
    import contextlib

    import torch

    @contextlib.contextmanager
    def cm():
        if not TIMING["enabled"]:
            yield
            return
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        yield
        b.record()
        TIMING[kind].append((a, b))
    return cm()


def lib():
    """Load (once) and return the CDLL.  No fallback: a missing library is an error."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing -- build it with `python -m cmcd_b200.build` (nvcc, sm_100a). "
                "cmcd_b200 has no CPU fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(l, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().cmcd_last_error().decode()
        if "not implemented" in msg.lower() or "not in the registry" in msg or "no small-d" in msg:
            raise NotImplementedError(msg)
        raise RuntimeError(f"cmcd_b200 error {rc}: {msg}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_contiguous(), "cmcd_b200 needs contiguous tensors"
    return t.data_ptr()


def require_cuda(*tensors):
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("cmcd_b200: no CUDA device -- the bridge hot path has no CPU fallback")
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("cmcd_b200: tensors must live on a CUDA device (no CPU fallback)")


def current_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream
