"""ctypes binding of libffno_b200.so (the C ABI declared in include/ffno_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, a RuntimeError is
raised.  The library is built in-tree by ``python -m fourierflow_b200.build`` (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FFNO_B200_LIB: load another build of the same library (the FFNO_TIMELINE=1 diagnostics build of tools/*_timeline.py)
LIB_PATH = os.environ.get("FFNO_B200_LIB") or os.path.join(_HERE, "lib", "libffno_b200.so")

ABI_VERSION = 2
MAX_DIMS = 3
MAX_FF_LAYERS = 4

OK = 0
MODE = {"full": 0, "low-pass": 1, "no-fourier": 2}
PATH = {"auto": 0, "generic": 1, "umma": 2}

c_float_p = C.POINTER(C.c_float)


class Desc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("ndim", C.c_int32),
        ("size", C.c_int32 * MAX_DIMS), ("pad", C.c_int32 * MAX_DIMS), ("modes", C.c_int32 * MAX_DIMS),
        ("width", C.c_int32), ("in_features", C.c_int32), ("append_grid", C.c_int32),
        ("out_features", C.c_int32), ("head_hidden", C.c_int32), ("n_layers", C.c_int32),
        ("ff_factor", C.c_int32), ("n_ff_layers", C.c_int32), ("layer_norm", C.c_int32),
        ("use_fork", C.c_int32), ("spectral_mode", C.c_int32), ("path", C.c_int32), ("transform", C.c_int32),
    ]


class LinearParams(C.Structure):
    _fields_ = [("weight", C.c_void_p), ("weight_g", C.c_void_p), ("weight_v", C.c_void_p),
                ("bias", C.c_void_p), ("in_features", C.c_int32), ("out_features", C.c_int32)]


class FFParams(C.Structure):
    _fields_ = [("linear", LinearParams * MAX_FF_LAYERS), ("ln_weight", C.c_void_p), ("ln_bias", C.c_void_p)]


class LayerParams(C.Structure):
    _fields_ = [("fourier_weight", C.c_void_p * MAX_DIMS), ("backcast_ff", FFParams), ("forecast_ff", FFParams)]


class BlockParams(C.Structure):
    _fields_ = [("in_proj", LinearParams), ("out0", LinearParams), ("out1", LinearParams),
                ("layers", C.POINTER(LayerParams)), ("n_layers", C.c_int32)]


class LinearGrads(C.Structure):
    _fields_ = [("weight", C.c_void_p), ("weight_g", C.c_void_p), ("weight_v", C.c_void_p), ("bias", C.c_void_p)]


class LayerGrads(C.Structure):
    _fields_ = [("fourier_weight", C.c_void_p * MAX_DIMS), ("backcast_ff", LinearGrads * MAX_FF_LAYERS)]


class BlockGrads(C.Structure):
    _fields_ = [("in_proj", LinearGrads), ("out0", LinearGrads), ("out1", LinearGrads),
                ("layers", C.POINTER(LayerGrads)), ("n_layers", C.c_int32)]


class RolloutExtras(C.Structure):
    _fields_ = [("use_velocity", C.c_int32), ("force_steps", C.c_int32), ("force", C.c_void_p), ("mu", C.c_void_p)]


class Taps(C.Structure):
    _fields_ = [("lift", C.c_void_p), ("x_after", C.POINTER(C.c_void_p)), ("spectral", C.POINTER(C.c_void_p)),
                ("b_last", C.c_void_p), ("forecast_list", C.POINTER(C.c_void_p))]


EXPORTS = {
    # name: (restype, argtypes)
    "ffno_last_error": (C.c_char_p, []),
    "ffno_abi_version": (C.c_int, []),
    "ffno_device_ok": (C.c_int, []),
    "ffno_plan_create": (C.c_int, [C.POINTER(Desc), C.POINTER(C.c_void_p)]),
    "ffno_plan_destroy": (C.c_int, [C.c_void_p]),
    "ffno_plan_uses_umma": (C.c_int, [C.c_void_p]),
    "ffno_plan_load_params": (C.c_int, [C.c_void_p, C.POINTER(BlockParams), C.c_void_p]),
    "ffno_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int32]),
    "ffno_block_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(Taps), C.c_void_p,
                                 C.c_size_t, C.c_void_p]),
    "ffno_workspace_bytes_host": (C.c_size_t, [C.c_void_p, C.c_int32]),
    "ffno_block_fwd_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t,
                                      C.c_void_p]),
    "ffno_spectral_fwd": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                    C.c_size_t, C.c_void_p]),
    "ffno_spectral_split_fwd": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.c_void_p,
                                          C.c_size_t, C.c_void_p]),
    "ffno_ff_fwd": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                              C.c_void_p, C.c_size_t, C.c_void_p]),
    "ffno_linear_fwd": (C.c_int, [C.POINTER(LinearParams), C.c_void_p, C.c_int64, C.c_void_p, C.c_int32,
                                  C.c_void_p, C.c_size_t, C.c_void_p]),
    "ffno_layernorm_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                     C.c_void_p]),
    "ffno_rel_l2": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int32,
                              C.c_int64, C.c_void_p, C.c_void_p]),
    "ffno_rollout_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int32, C.c_int32]),
    "ffno_rollout_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, c_float_p, c_float_p,
                                   C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ffno_rollout_workspace_bytes_ex": (C.c_size_t, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "ffno_rollout_fwd_ex": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, c_float_p, c_float_p,
                                      C.c_float, C.c_float, C.POINTER(RolloutExtras), C.c_void_p, C.c_void_p, C.c_size_t,
                                      C.c_void_p]),
    "ffno_block_bwd_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int32]),
    "ffno_block_bwd": (C.c_int, [C.c_void_p, C.POINTER(BlockParams), C.c_void_p, C.c_void_p, C.c_int32,
                                 C.POINTER(BlockGrads), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ffno_layers_bwd": (C.c_int, [C.c_void_p, C.POINTER(BlockParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                  C.POINTER(BlockGrads), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ffno_plan_set_backward_mode": (C.c_int, [C.c_void_p, C.c_int32]),
    "ffno_rel_l2_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
    "ffno_velocity_scratch_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "ffno_velocity_fwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                    C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ffno_plan_set_domain": (C.c_int, [C.c_void_p, C.c_float, C.c_float]),
    "ffno_umma_selftest": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                     C.c_int32, C.c_void_p]),
    "ffno_debug_timeline": (C.c_int, [C.c_int32, C.c_void_p]),
    "ffno_plan_last_launch_count": (C.c_int64, [C.c_void_p]),
    "ffno_plan_graph_active": (C.c_int, [C.c_void_p]),
    "ffno_plan_pipeline_unit": (C.c_int, [C.c_void_p, C.c_int32]),
    "ffno_debug_pipe_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once) and set the prototypes of every export."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing — build it with `python -m fourierflow_b200.build` "
            "(fourierflow_b200 has no CPU or PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)          # AttributeError here == ABI mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.ffno_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libffno_b200 ABI {lib.ffno_abi_version()} != binding ABI {ABI_VERSION}")
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != OK:
        msg = load().ffno_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (status {status}): {msg}")
