"""ctypes binding of libmrla_b200.so (C ABI declared in include/mrla_b200.h).

There is no CPU fallback and no alternative backend: if the shared library is missing or a
tensor is not on a CUDA device the call raises.
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmrla_b200.so")
ABI_VERSION = 4

F32, BF16, F16 = 0, 1, 2
NCHW, NHWC = 0, 1
ACT_NONE, ACT_GELU = 0, 1
BN_NONE, BN_TRAIN, BN_EVAL = 0, 1, 2

ERRORS = {
    -1: "MRLA_ERR_NULL (a required pointer is NULL)",
    -2: "MRLA_ERR_SHAPE (B,C,H,W,d,k out of the supported range)",
    -3: "MRLA_ERR_ALIGN (pointer / stride / channel count not aligned for the vector width)",
    -4: "MRLA_ERR_UNSUPPORTED (dtype / layout / flag combination not implemented)",
    -5: "MRLA_ERR_WORKSPACE (scratch buffer too small)",
}

_vp, _i32, _i64, _f32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float


class MrlaLightArgs(ctypes.Structure):
    """Mirror of `struct MrlaLightArgs` (include/mrla_b200.h) — field order must match."""
    _fields_ = (
        [(n, _i32) for n in ("B", "C", "H", "W", "dim_perhead", "k_size", "dtype", "layout", "act", "bn_mode",
                             "residual", "update_running", "fuse_relu_bwd", "x_virtual")]
        + [("eps", _f32), ("momentum", _f32)]
        + [(n, _i64) for n in ("bs_x", "bs_o", "bs_y", "bs_dy", "bs_dx", "bs_do")]
        + [(n, _vp) for n in ("x", "o", "y", "wq", "wk", "wv", "lam", "gamma", "beta", "running_mean", "running_var",
                              "drop_scale", "mom", "gate", "mean", "rstd", "coef", "dy", "dx", "dout", "dwq", "dwk",
                              "dwv", "dlam", "dgamma", "dbeta", "gmom", "bcoef", "scratch")]
        + [("scratch_bytes", ctypes.c_size_t), ("z", _vp), ("bs_z", _i64), ("z_coef", _vp), ("dz_sums", _vp)]
    )


class MrlaBaseArgs(ctypes.Structure):
    """Mirror of `struct MrlaBaseArgs` (include/mrla_b200.h) — field order must match."""
    _fields_ = (
        [(n, _i32) for n in ("B", "C", "H", "W", "dim_perhead", "k_size", "dtype", "layout", "t", "t_cap", "bn_mode",
                             "relu", "residual", "update_running", "accumulate")]
        + [("eps", _f32), ("momentum", _f32)]
        + [(n, _i64) for n in ("bs_x", "bs_y", "bs_s", "bs_dy", "bs_dx", "bs_v", "ts_v", "bs_dv", "ts_dv")]
        + [(n, _vp) for n in ("x", "v", "s", "y", "kcache", "wq", "wk", "wv", "gamma", "beta", "running_mean",
                              "running_var", "drop_scale", "sx", "q", "p", "smom", "chan", "dy", "dx", "dv", "dkcache",
                              "dwq", "dwk", "dwv", "dgamma", "dbeta", "gmom", "dpm", "dyc", "scratch")]
        + [("scratch_bytes", ctypes.c_size_t)]
    )


class MrlaBnArgs(ctypes.Structure):
    """Mirror of `struct MrlaBnArgs` (include/mrla_b200.h)."""
    _fields_ = (
        [("M", _i64)]
        + [(n, _i32) for n in ("C", "dtype", "relu", "training", "update_running", "stats_only")]
        + [("eps", _f32), ("momentum", _f32)]
        + [(n, _vp) for n in ("x", "y", "gamma", "beta", "running_mean", "running_var", "stats", "coef", "dy", "dx",
                              "dgamma", "dbeta", "scratch")]
        + [("scratch_bytes", ctypes.c_size_t), ("sums", _vp)]
    )


class MrlaDeitArgs(ctypes.Structure):
    """Mirror of `struct MrlaDeitArgs` (include/mrla_b200.h)."""
    _fields_ = (
        [(n, _i32) for n in ("B", "n", "C", "S", "dim_perhead", "k_size", "dtype", "reserved0")]
        + [("eps", _f32), ("reserved1", _f32)]
        + [(n, _vp) for n in ("x", "o", "out", "normx_w", "normx_b", "normo_w", "normo_b", "wq", "wk", "wv", "lam",
                              "stats_x", "stats_o", "gate", "dout", "dx", "dox", "dparams", "scratch")]
        + [("scratch_bytes", ctypes.c_size_t)]
    )


class MrlaLnArgs(ctypes.Structure):
    """Mirror of `struct MrlaLnArgs` (include/mrla_b200.h)."""
    _fields_ = ([(n, _i32) for n in ("B", "n", "C", "dtype")] + [("eps", _f32), ("reserved0", _f32)]
                + [("x", _vp), ("xn", _vp), ("cls_out", _vp), ("bs_cls", _i64), ("gamma", _vp), ("beta", _vp), ("stats", _vp),
                   ("g_cls", _vp), ("bs_gcls", _i64), ("g_img", _vp), ("bs_gimg", _i64), ("dx", _vp), ("dparams", _vp),
                   ("scratch", _vp), ("scratch_bytes", ctypes.c_size_t)])


_lib = None
_lock = threading.Lock()

EXPORTS = (
    "mrla_abi_version", "mrla_build_info", "mrla_last_launch_count", "mrla_last_error_site", "mrla_sizeof_light_args",
    "mrla_light_bwd_scratch_bytes", "mrla_light_bwd_fuses_relu", "mrla_light_fwd_folds_bn", "mrla_light_virtual_x", "mrla_light_v7_plan", "mrla_light_forward", "mrla_light_backward",
    "mrla_nchw_to_nhwc", "mrla_add_relu", "mrla_sizeof_base_args", "mrla_base_bwd_scratch_bytes", "mrla_base_forward", "mrla_base_backward",
    "mrla_maxpool3x3s2_forward", "mrla_maxpool3x3s2_backward",
    "mrla_sizeof_bn_args", "mrla_bn_scratch_bytes", "mrla_bn_forward", "mrla_bn_backward",
    "mrla_sizeof_deit_args", "mrla_deit_light_supported", "mrla_deit_light_scratch_bytes", "mrla_deit_light_forward",
    "mrla_deit_light_backward",
    "mrla_sizeof_ln_args", "mrla_layernorm_scratch_bytes", "mrla_layernorm_forward", "mrla_layernorm_backward",
)


def lib() -> ctypes.CDLL:
    """Load libmrla_b200.so once; fail loudly if it is absent or has the wrong ABI."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA extension has not been built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C mrla_b200/csrc -j`). "
                "mrla_b200 has no CPU or PyTorch fallback.")
        L = ctypes.CDLL(LIB_PATH)
        L.mrla_abi_version.restype = ctypes.c_int
        if L.mrla_abi_version() != ABI_VERSION:
            raise RuntimeError(f"libmrla_b200.so ABI {L.mrla_abi_version()} != expected {ABI_VERSION}; rebuild")
        L.mrla_build_info.restype = ctypes.c_char_p
        L.mrla_last_launch_count.restype = ctypes.c_int
        L.mrla_last_error_site.restype = ctypes.c_char_p
        L.mrla_sizeof_light_args.restype = ctypes.c_size_t
        if L.mrla_sizeof_light_args() != ctypes.sizeof(MrlaLightArgs):
            raise RuntimeError("MrlaLightArgs layout mismatch between _lib.py and include/mrla_b200.h")
        L.mrla_nchw_to_nhwc.restype = ctypes.c_int
        L.mrla_nchw_to_nhwc.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]
        L.mrla_light_bwd_fuses_relu.restype = ctypes.c_int
        L.mrla_light_bwd_fuses_relu.argtypes = [ctypes.POINTER(MrlaLightArgs)]
        L.mrla_light_fwd_folds_bn.restype = ctypes.c_int
        L.mrla_light_fwd_folds_bn.argtypes = [ctypes.POINTER(MrlaLightArgs)]
        L.mrla_light_virtual_x.restype = ctypes.c_int
        L.mrla_light_virtual_x.argtypes = [ctypes.POINTER(MrlaLightArgs)]
        L.mrla_add_relu.restype = ctypes.c_int
        L.mrla_add_relu.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
                                    ctypes.c_void_p]
        for d in ("forward", "backward"):
            f = getattr(L, f"mrla_maxpool3x3s2_{d}")
            f.restype = ctypes.c_int
            f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                          ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.mrla_sizeof_base_args.restype = ctypes.c_size_t
        if L.mrla_sizeof_base_args() != ctypes.sizeof(MrlaBaseArgs):
            raise RuntimeError("MrlaBaseArgs layout mismatch between _lib.py and include/mrla_b200.h")
        L.mrla_sizeof_bn_args.restype = ctypes.c_size_t
        if L.mrla_sizeof_bn_args() != ctypes.sizeof(MrlaBnArgs):
            raise RuntimeError("MrlaBnArgs layout mismatch between _lib.py and include/mrla_b200.h")
        L.mrla_bn_scratch_bytes.restype = ctypes.c_size_t
        L.mrla_bn_scratch_bytes.argtypes = [ctypes.POINTER(MrlaBnArgs)]
        for d in ("forward", "backward"):
            f = getattr(L, f"mrla_bn_{d}")
            f.restype = ctypes.c_int
            f.argtypes = [ctypes.POINTER(MrlaBnArgs), ctypes.c_void_p]
        L.mrla_sizeof_deit_args.restype = ctypes.c_size_t
        if L.mrla_sizeof_deit_args() != ctypes.sizeof(MrlaDeitArgs):
            raise RuntimeError("MrlaDeitArgs layout mismatch between _lib.py and include/mrla_b200.h")
        L.mrla_deit_light_supported.restype = ctypes.c_int
        L.mrla_deit_light_supported.argtypes = [ctypes.POINTER(MrlaDeitArgs)]
        L.mrla_deit_light_scratch_bytes.restype = ctypes.c_size_t
        L.mrla_deit_light_scratch_bytes.argtypes = [ctypes.POINTER(MrlaDeitArgs)]
        for d in ("forward", "backward"):
            f = getattr(L, f"mrla_deit_light_{d}")
            f.restype = ctypes.c_int
            f.argtypes = [ctypes.POINTER(MrlaDeitArgs), ctypes.c_void_p]
        L.mrla_sizeof_ln_args.restype = ctypes.c_size_t
        if L.mrla_sizeof_ln_args() != ctypes.sizeof(MrlaLnArgs):
            raise RuntimeError("MrlaLnArgs layout mismatch between _lib.py and include/mrla_b200.h")
        L.mrla_layernorm_scratch_bytes.restype = ctypes.c_size_t
        L.mrla_layernorm_scratch_bytes.argtypes = [ctypes.POINTER(MrlaLnArgs)]
        for d in ("forward", "backward"):
            f = getattr(L, f"mrla_layernorm_{d}")
            f.restype = ctypes.c_int
            f.argtypes = [ctypes.POINTER(MrlaLnArgs), ctypes.c_void_p]
        for name, st in (("light", MrlaLightArgs), ("base", MrlaBaseArgs)):
            f = getattr(L, f"mrla_{name}_bwd_scratch_bytes")
            f.restype = ctypes.c_size_t
            f.argtypes = [ctypes.POINTER(st)]
            for d in ("forward", "backward"):
                f = getattr(L, f"mrla_{name}_{d}")
                f.restype = ctypes.c_int
                f.argtypes = [ctypes.POINTER(st), ctypes.c_void_p]
        _lib = L
    return _lib


def check(rc: int, what: str):
    if rc == 0:
        return
    if rc < 0:
        site = lib().mrla_last_error_site().decode() if rc == -4 else ""
        raise RuntimeError(f"{what}: {ERRORS.get(rc, rc)}" + (f" [{site.rsplit('/', 1)[-1]}]" if site else ""))
    raise RuntimeError(f"{what}: CUDA error {rc} at launch")
