"""ctypes binding of libzkmsm.so (include/zkmsm.h).  No fallback: if the CUDA library is missing
or no sm_100 device is usable, importing works but every operation raises ZkmsmError."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ZKMSM_LIB") or os.path.join(_HERE, "libzkmsm.so")   # override: A/B builds of the library

OK = 0
ERR_NAMES = {
    -1: "ZKMSM_ERR_INVALID_ARG", -2: "ZKMSM_ERR_CUDA", -3: "ZKMSM_ERR_SCALAR_RANGE",
    -4: "ZKMSM_ERR_NO_DEVICE", -5: "ZKMSM_ERR_TOO_FEW_POINTS", -6: "ZKMSM_ERR_NOMEM", -7: "ZKMSM_ERR_NOT_IN_SUBGROUP",
}
PRECOMPUTE = 1
SUBGROUP = 2
CHECK_SUBGROUP = 4
G1_WORDS, G2_WORDS = 24, 48
G1_PARTIAL_WORDS, G2_PARTIAL_WORDS = 48, 96
GROTH16_PARTIAL_WORDS = 192


class CrsDesc(ctypes.Structure):
    """zkmsm_crs_desc (include/zkmsm.h)"""
    _fields_ = [("g1_alpha", ctypes.c_void_p), ("g1_beta", ctypes.c_void_p), ("g1_delta", ctypes.c_void_p),
                ("g1_xi", ctypes.c_void_p), ("g1_xi_inf", ctypes.c_void_p), ("n", ctypes.c_size_t),
                ("g1_uvw_wit", ctypes.c_void_p), ("g1_uvw_wit_inf", ctypes.c_void_p), ("n_wit", ctypes.c_size_t),
                ("g1_xt_by_delta", ctypes.c_void_p), ("g1_xt_by_delta_inf", ctypes.c_void_p), ("n_xt", ctypes.c_size_t),
                ("g2_beta", ctypes.c_void_p), ("g2_delta", ctypes.c_void_p), ("g2_xi", ctypes.c_void_p),
                ("g2_xi_inf", ctypes.c_void_p)]


class ZkmsmError(RuntimeError):
    def __init__(self, code, msg=""):
        self.code = code
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")


# every symbol include/zkmsm.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "zkmsm_version", "zkmsm_create", "zkmsm_destroy", "zkmsm_set_stream", "zkmsm_last_error", "zkmsm_set_window",
    "zkmsm_set_option",
    "zkmsm_host_alloc", "zkmsm_host_free",
    "zkmsm_g1_load_points", "zkmsm_g2_load_points", "zkmsm_points_free", "zkmsm_points_len", "zkmsm_points_info", "zkmsm_points_read",
    "zkmsm_g1_msm", "zkmsm_g2_msm", "zkmsm_g1_msm_device", "zkmsm_g2_msm_device",
    "zkmsm_g1_msm_oneshot", "zkmsm_g2_msm_oneshot",
    "zkmsm_g1_msm_enqueue", "zkmsm_g2_msm_enqueue", "zkmsm_g1_msm_begin", "zkmsm_g2_msm_begin", "zkmsm_g1_msm_result", "zkmsm_g2_msm_result",
    "zkmsm_last_launch_count", "zkmsm_profile", "zkmsm_profile_read",
    "zkmsm_g1_msm_partial", "zkmsm_g2_msm_partial", "zkmsm_g1_msm_partial_device", "zkmsm_g2_msm_partial_device",
    "zkmsm_g1_msm_partial_range", "zkmsm_g2_msm_partial_range",
    "zkmsm_g1_msm_partial_range_device", "zkmsm_g2_msm_partial_range_device",
    "zkmsm_g1_combine", "zkmsm_g2_combine", "zkmsm_g1_combine_device", "zkmsm_g2_combine_device",
    "zkmsm_g1_combine_enqueue", "zkmsm_g2_combine_enqueue",
    "zkmsm_g1_mul_base", "zkmsm_g2_mul_base", "zkmsm_g1_points_from_scalars", "zkmsm_g2_points_from_scalars",
    "zkmsm_fr_aggregate", "zkmsm_fr_quotient", "zkmsm_bench_imad",
    "zkmsm_crs_load", "zkmsm_crs_free", "zkmsm_crs_sizes", "zkmsm_groth16_prove", "zkmsm_groth16_prove_partial",
    "zkmsm_groth16_combine",
]

_lib = None


def load():
    """Load the shared library once; raises if it has not been built (python __graft_entry__.py)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ZkmsmError(-2, f"{LIB_PATH} not built; run `make -C zk-toolkit_b200/csrc -j8` "
                             "(there is no CPU implementation to fall back to)")
    lib = ctypes.CDLL(LIB_PATH)
    vp, u32p, u8p, sz, ci, cu = (ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint8),
                                 ctypes.c_size_t, ctypes.c_int, ctypes.c_uint)
    ip = ctypes.POINTER(ctypes.c_int)
    vpp = ctypes.POINTER(ctypes.c_void_p)
    dp = ctypes.POINTER(ctypes.c_double)
    sig = {
        "zkmsm_version": (ctypes.c_char_p, []),
        "zkmsm_create": (ci, [ci, vpp]),
        "zkmsm_destroy": (ci, [vp]),
        "zkmsm_set_stream": (ci, [vp, vp]),
        "zkmsm_last_error": (ctypes.c_char_p, [vp]),
        "zkmsm_set_window": (ci, [vp, cu]),
        "zkmsm_set_option": (ci, [vp, ctypes.c_char_p, ctypes.c_long]),
        "zkmsm_host_alloc": (ci, [sz, vpp]),
        "zkmsm_host_free": (ci, [vp]),
        "zkmsm_g1_load_points": (ci, [vp, vp, vp, sz, cu, vpp]),
        "zkmsm_g2_load_points": (ci, [vp, vp, vp, sz, cu, vpp]),
        "zkmsm_points_free": (ci, [vp, vp]),
        "zkmsm_points_len": (sz, [vp]),
        "zkmsm_points_info": (ci, [vp, ctypes.POINTER(cu), ctypes.POINTER(cu), ip, ip]),
        "zkmsm_points_read": (ci, [vp, vp, sz, sz, vp, vp]),
        "zkmsm_g1_msm": (ci, [vp, vp, vp, sz, vp, ip]),
        "zkmsm_g2_msm": (ci, [vp, vp, vp, sz, vp, ip]),
        "zkmsm_g1_msm_device": (ci, [vp, vp, vp, sz, vp, ip]),
        "zkmsm_g2_msm_device": (ci, [vp, vp, vp, sz, vp, ip]),
        "zkmsm_g1_msm_oneshot": (ci, [vp, vp, vp, vp, sz, vp, ip]),
        "zkmsm_g2_msm_oneshot": (ci, [vp, vp, vp, vp, sz, vp, ip]),
        "zkmsm_g1_msm_begin": (ci, [vp, vp, vp, sz]),
        "zkmsm_g2_msm_begin": (ci, [vp, vp, vp, sz]),
        "zkmsm_g1_msm_enqueue": (ci, [vp, vp, vp, sz]),
        "zkmsm_g2_msm_enqueue": (ci, [vp, vp, vp, sz]),
        "zkmsm_g1_msm_result": (ci, [vp, vp, ip]),
        "zkmsm_g2_msm_result": (ci, [vp, vp, ip]),
        "zkmsm_last_launch_count": (ci, [vp]),
        "zkmsm_profile": (ci, [vp, ci]),
        "zkmsm_profile_read": (ci, [vp, ci, vp, sz, vp, vp]),
        "zkmsm_g1_msm_partial": (ci, [vp, vp, vp, sz, vp]),
        "zkmsm_g2_msm_partial": (ci, [vp, vp, vp, sz, vp]),
        "zkmsm_g1_msm_partial_device": (ci, [vp, vp, vp, sz, vp]),
        "zkmsm_g2_msm_partial_device": (ci, [vp, vp, vp, sz, vp]),
        "zkmsm_g1_msm_partial_range": (ci, [vp, vp, vp, sz, cu, cu, vp]),
        "zkmsm_g2_msm_partial_range": (ci, [vp, vp, vp, sz, cu, cu, vp]),
        "zkmsm_g1_msm_partial_range_device": (ci, [vp, vp, vp, sz, cu, cu, vp]),
        "zkmsm_g2_msm_partial_range_device": (ci, [vp, vp, vp, sz, cu, cu, vp]),
        "zkmsm_g1_combine": (ci, [vp, vp, sz, vp, ip]),
        "zkmsm_g2_combine": (ci, [vp, vp, sz, vp, ip]),
        "zkmsm_g1_combine_device": (ci, [vp, vp, sz, vp, ip]),
        "zkmsm_g2_combine_device": (ci, [vp, vp, sz, vp, ip]),
        "zkmsm_g1_combine_enqueue": (ci, [vp, vp, sz]),
        "zkmsm_g2_combine_enqueue": (ci, [vp, vp, sz]),
        "zkmsm_g1_mul_base": (ci, [vp, vp, vp, sz, vp, vp]),
        "zkmsm_g2_mul_base": (ci, [vp, vp, vp, sz, vp, vp]),
        "zkmsm_g1_points_from_scalars": (ci, [vp, vp, vp, sz, cu, vpp]),
        "zkmsm_g2_points_from_scalars": (ci, [vp, vp, vp, sz, cu, vpp]),
        "zkmsm_fr_aggregate": (ci, [vp, vp, sz, sz, vp, vp]),
        "zkmsm_fr_quotient": (ci, [vp, vp, vp, vp, sz, vp, ip]),
        "zkmsm_bench_imad": (ci, [vp, ci, ci, dp, dp]),
        "zkmsm_crs_load": (ci, [vp, ctypes.POINTER(CrsDesc), cu, vpp]),
        "zkmsm_crs_free": (ci, [vp, vp]),
        "zkmsm_crs_sizes": (ci, [vp, ctypes.POINTER(sz), ctypes.POINTER(sz), ctypes.POINTER(sz)]),
        "zkmsm_groth16_prove": (ci, [vp, vp, vp, vp, vp, vp, vp, vp, vp, ip]),
        "zkmsm_groth16_prove_partial": (ci, [vp, vp, vp, vp, vp, vp, vp, vp, cu, cu, vp]),
        "zkmsm_groth16_combine": (ci, [vp, vp, sz, vp, ip]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def as_u32(a, shape_last=None):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    if shape_last is not None and (a.ndim == 0 or a.shape[-1] != shape_last):
        raise ValueError(f"expected last dimension {shape_last}, got shape {a.shape}")
    return a


def dptr(a):
    """host pointer of a numpy array (or None)"""
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)
