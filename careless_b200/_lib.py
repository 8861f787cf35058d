"""ctypes binding of ``libcareless_b200.so`` (C-ABI declared in ``include/careless_b200.h``).

There is deliberately no fallback: if the shared library has not been built
(``python -m careless_b200.build``) importing the symbols raises, and ``clb_create``
fails on a machine without an sm_100 GPU.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CLB_LIB_PATH") or os.path.join(HERE, "libcareless_b200.so")   # CLB_LIB_PATH: instrumented debug builds (tools/)
ABI_VERSION = 4

# enums of include/careless_b200.h
LIK_NORMAL, LIK_STUDENTT = 0, 1
PRIOR_WILSON, PRIOR_DOUBLE_WILSON = 0, 1
BIJ_EXP, BIJ_SOFTPLUS = 0, 1
ORDER_AUTO, ORDER_REFL, ORDER_SPOT, ORDER_IMAGE, ORDER_NONE = 0, 1, 2, 3, 4
GROUP_SF_LOC, GROUP_SF_SCALE, GROUP_MLP, GROUP_IMAGE_SCALES, GROUP_DW_R, GROUP_IMAGE_LAYERS, GROUP_LIKELIHOOD = 0, 1, 2, 3, 4, 5, 6
GROUPS = {"sf_loc_raw": GROUP_SF_LOC, "sf_scale_raw": GROUP_SF_SCALE, "mlp": GROUP_MLP,
          "image_scales": GROUP_IMAGE_SCALES, "dw_r_logit": GROUP_DW_R, "image_layers": GROUP_IMAGE_LAYERS, "likelihood": GROUP_LIKELIHOOD}


class clb_config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32), ("stream", C.c_void_p),
        ("n_refl", C.c_int64), ("n_refl_total", C.c_int64),
        ("n_meta", C.c_int32), ("mlp_width", C.c_int32), ("mlp_layers", C.c_int32),
        ("n_images", C.c_int32), ("image_scales", C.c_int32), ("mc_samples", C.c_int32),
        ("likelihood", C.c_int32), ("dof", C.c_float), ("laue", C.c_int32),
        ("prior", C.c_int32), ("n_asu", C.c_int32), ("optimize_dw_r", C.c_int32),
        ("scale_bijector", C.c_int32), ("scale_shift", C.c_float), ("epsilon", C.c_float),
        ("use_kl_weight", C.c_int32), ("kl_weight", C.c_float),
        ("learning_rate", C.c_float), ("beta_1", C.c_float), ("beta_2", C.c_float), ("adam_epsilon", C.c_float),
        ("clipnorm", C.c_float), ("clipvalue", C.c_float), ("global_clipnorm", C.c_float),
        ("seed", C.c_uint64), ("rank", C.c_int32), ("world_size", C.c_int32), ("image_layers", C.c_int32), ("deterministic", C.c_int32), ("refine_uncertainties", C.c_int32),
    ]


class clb_metrics(C.Structure):
    _fields_ = [("loss", C.c_double), ("nll", C.c_double), ("kl", C.c_double), ("grad_norm", C.c_double)]


# every symbol include/careless_b200.h declares: name -> (restype, argtypes)
_H = C.c_void_p
_F = C.POINTER(C.c_float)
_I64 = C.POINTER(C.c_int64)
SYMBOLS = {
    "clb_abi_version": (C.c_int, []),
    "clb_last_error": (C.c_char_p, [_H]),
    "clb_create": (C.c_int, [C.POINTER(clb_config), C.POINTER(_H)]),
    "clb_destroy": (None, [_H]),
    "clb_set_observations": (C.c_int, [_H, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "clb_prepare_rows": (C.c_int, [C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_int32, C.c_int32, C.c_int64, _I64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]),
    "clb_download_rows": (C.c_int, [_H, C.c_int64, _I64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "clb_upload_observations": (C.c_int, [_H]),
    "clb_prefetch_observations": (C.c_int, [_H]),
    "clb_set_prior": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_float]),
    "clb_group_size": (C.c_int64, [_H, C.c_int32]),
    "clb_get_params": (C.c_int, [_H, C.c_int32, C.c_void_p, C.c_int64]),
    "clb_set_params": (C.c_int, [_H, C.c_int32, C.c_void_p, C.c_int64]),
    "clb_get_grads": (C.c_int, [_H, C.c_int32, C.c_void_p, C.c_int64]),
    "clb_get_adam_state": (C.c_int, [_H, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, _I64]),
    "clb_set_trainable": (C.c_int, [_H, C.c_int32, C.c_int32]),
    "clb_step": (C.c_int, [_H, C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(clb_metrics), C.POINTER(C.c_int32)]),
    "clb_eval": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.POINTER(clb_metrics)]),
    "clb_step_begin": (C.c_int, [_H, C.c_void_p, C.c_void_p]),
    "clb_step_norms": (C.c_int, [_H]),
    "clb_step_end": (C.c_int, [_H, C.POINTER(clb_metrics)]),
    "clb_reduce_buffers": (C.c_int, [_H, C.POINTER(C.c_void_p), _I64, C.POINTER(C.c_void_p), _I64]),
    "clb_comm_unique_id": (C.c_int, [C.c_void_p]),
    "clb_comm_init": (C.c_int, [_H, C.c_void_p]),
    "clb_get_samples": (C.c_int, [_H, C.c_void_p, C.c_int64]),
    "clb_enable_ipred": (C.c_int, [_H, C.c_int32]),
    "clb_get_ipred": (C.c_int, [_H, C.c_void_p, C.c_int64]),
    "clb_get_results": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "clb_get_scale_moments": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int64]),
    "clb_synchronize": (C.c_int, [_H]),
    "clb_kernel_time_ms": (C.c_int, [_H, C.POINTER(C.c_double), _I64, _I64]),
    "clb_reset_timers": (C.c_int, [_H, C.c_int32]),
}

_lib = None


class LibraryNotBuilt(RuntimeError):
    pass


def load():
    """Load the CUDA library; raise loudly if it is missing (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryNotBuilt(
            f"{LIB_PATH} is missing: build it with `python -m careless_b200.build` "
            "(nvcc, sm_100a).  careless_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    if lib.clb_abi_version() != ABI_VERSION:
        raise RuntimeError(f"ABI mismatch: library {lib.clb_abi_version()} vs binding {ABI_VERSION}")
    _lib = lib
    return lib


class ClbError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"careless_b200 error {code}: {message}")
        self.code = code


def check(rc, handle=None):
    if rc != 0:
        msg = load().clb_last_error(handle)
        raise ClbError(rc, msg.decode() if msg else "unknown error")


def comm_unique_id() -> bytes:
    """128-byte id of a new library-side communicator (rank 0 calls this and hands the bytes to the other ranks)."""
    buf = (C.c_uint8 * 128)()
    check(load().clb_comm_unique_id(buf))
    return bytes(buf)
