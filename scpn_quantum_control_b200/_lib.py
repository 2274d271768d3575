"""ctypes binding of ``libxyk.so`` (C ABI declared in ``include/xyk.h``).

There is no CPU fallback: if the CUDA library is missing or cannot be loaded, importing the
engine raises ``XYKLibraryError`` -- the product path never routes through the oracle.
"""

from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_size_t, c_void_p, c_float

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libxyk.so")

XYK_OK = 0
XYK_ERR_INVALID = 1
XYK_ERR_CUDA = 2
XYK_ERR_NOMEM = 3
XYK_ERR_STATE = 4
XYK_ERR_UNSUPPORTED = 5
XYK_NEED_EXCHANGE = 100
XYK_NEED_BARRIER = 101

ENGINE_AUTO, ENGINE_SIMPLE, ENGINE_TILED = 0, 1, 2
OPT_ENGINE, OPT_FUSE_PHASE, OPT_TILED_OBS, OPT_PEER_EXCHANGE = 1, 2, 3, 4


class XYKLibraryError(RuntimeError):
    """libxyk.so is missing or unusable (build it with ``python __graft_entry__.py``)."""


class XYKStats(ctypes.Structure):
    _fields_ = [
        ("kernel_launches", c_int64),
        ("pass_launches", c_int64),
        ("state_sweeps", c_int64),
        ("last_passes", c_int32),
        ("last_rounds", c_int32),
        ("last_gates", c_int32),
        ("last_layers1", c_int32),
        ("last_pass_ms", c_float),
        ("timing_enabled", c_int32),
        ("pass_kernel_ms", c_double),
        ("pass_kernel_timed", c_int64),
        ("peer_pass_ms", c_double),
        ("peer_passes", c_int64),
    ]


# name -> (restype, argtypes); every symbol include/xyk.h declares
PROTOTYPES = {
    "xyk_version": (c_int, []),
    "xyk_last_error": (c_char_p, [c_void_p]),
    "xyk_device_count": (c_int, [POINTER(c_int)]),
    "xyk_device_memory": (c_int, [c_int, POINTER(c_size_t), POINTER(c_size_t)]),
    "xyk_create": (c_int, [c_int, c_int, c_int, c_int, c_int, POINTER(c_void_p)]),
    "xyk_destroy": (c_int, [c_void_p]),
    "xyk_set_option": (c_int, [c_void_p, c_int, c_int64]),
    "xyk_set_problem": (c_int, [c_void_p, POINTER(c_double), POINTER(c_double)]),
    "xyk_term_counts": (c_int, [c_void_p, POINTER(c_int), POINTER(c_int)]),
    "xyk_init_product_state": (c_int, [c_void_p]),
    "xyk_apply_layers": (c_int, [c_void_p, c_double, c_int, c_int]),
    "xyk_continue": (c_int, [c_void_p]),
    "xyk_exchange_info": (c_int, [c_void_p, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_size_t)]),
    "xyk_exchange_done": (c_int, [c_void_p]),
    "xyk_begin_exchange": (c_int, [c_void_p, POINTER(c_int)]),
    "xyk_ipc_export": (c_int, [c_void_p, c_int, c_char_p]),
    "xyk_ipc_import": (c_int, [c_void_p, c_int, c_int, c_char_p]),
    "xyk_xy_expectations": (c_int, [c_void_p, POINTER(c_double), POINTER(c_double)]),
    "xyk_order_parameter": (c_int, [c_void_p, POINTER(c_double), POINTER(c_double)]),
    "xyk_z_expectations": (c_int, [c_void_p, POINTER(c_double)]),
    "xyk_correlation_matrix_xy": (c_int, [c_void_p, POINTER(c_double)]),
    "xyk_energy": (c_int, [c_void_p, POINTER(c_double)]),
    "xyk_norm2": (c_int, [c_void_p, POINTER(c_double)]),
    "xyk_set_problem_xxz": (c_int, [c_void_p, POINTER(c_double), POINTER(c_double), c_double]),
    "xyk_exact_step": (c_int, [c_void_p, c_double, c_double, POINTER(c_int)]),
    "xyk_xy_expectations_peer": (c_int, [c_void_p, POINTER(c_double), POINTER(c_double)]),
    "xyk_apply_occupation_decay": (c_int, [c_void_p, c_double, c_double, c_double]),
    "xyk_apply_jump": (c_int, [c_void_p, c_int, c_int]),
    "xyk_apply_ry": (c_int, [c_void_p, POINTER(c_double)]),
    "xyk_init_product_angles": (c_int, [c_void_p, POINTER(c_double)]),
    "xyk_run": (c_int, [c_void_p, POINTER(c_double), c_int, c_int, c_int, POINTER(c_double)]),
    "xyk_get_state": (c_int, [c_void_p, c_void_p, c_int]),
    "xyk_set_state": (c_int, [c_void_p, c_void_p, c_int]),
    "xyk_state_buffer": (c_int, [c_void_p, POINTER(c_void_p), POINTER(c_size_t)]),
    "xyk_get_layout": (c_int, [c_void_p, POINTER(c_int)]),
    "xyk_canonicalise": (c_int, [c_void_p]),
    "xyk_synchronize": (c_int, [c_void_p]),
    "xyk_stream": (c_int, [c_void_p, POINTER(c_void_p)]),
    "xyk_get_stats": (c_int, [c_void_p, POINTER(XYKStats)]),
    "xyk_enable_timing": (c_int, [c_void_p, c_int]),
    "xyk_plan_stages": (
        c_int64,
        [POINTER(c_double), POINTER(c_double), c_int, c_double, c_int, c_int, c_int, c_char_p, c_int64],
    ),
    "xyk_plan_describe": (
        c_int64,
        [c_int, c_int, POINTER(c_double), POINTER(c_double), c_int, c_double, c_int, c_int, c_char_p, c_int64],
    ),
    "xyk_batch_run": (
        c_int,
        [c_int, c_int, c_int, POINTER(c_double), POINTER(c_double), POINTER(c_double), c_int, c_int, c_int,
         POINTER(c_double), c_void_p],
    ),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load libxyk.so once and attach the prototypes.  Raises XYKLibraryError when unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise XYKLibraryError(
            f"{LIB_PATH} not found: the CUDA engine has not been built "
            "(run `python __graft_entry__.py` or `make -C scpn_quantum_control_b200/csrc`). "
            "There is no CPU fallback."
        )
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as exc:  # missing libcudart etc.
        raise XYKLibraryError(f"cannot load {LIB_PATH}: {exc}") from exc
    for name, (restype, argtypes) in PROTOTYPES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as exc:
            raise XYKLibraryError(f"{LIB_PATH} does not export {name}") from exc
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error(handle=None) -> str:
    msg = load().xyk_last_error(handle)
    return msg.decode("utf-8", "replace") if msg else ""
