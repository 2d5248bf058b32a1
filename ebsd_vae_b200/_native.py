"""ctypes binding of libebsd_b200.so (include/ebsd_b200.h).

There is no CPU or pure-PyTorch fallback: if the shared library is missing or a call fails, the
error is raised to the caller.
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# EBSD_B200_LIB selects another build of the same library (A/B timing of compile-time variants); no fallback either way
LIB_PATH = os.environ.get("EBSD_B200_LIB") or os.path.join(_HERE, "libebsd_b200.so")

OK = 0
PATTERN_U8 = 0
PATTERN_F32 = 1
ANGLE_RADIANS = 0
ANGLE_DEGREES = 1
MAX_TOPK = 32
N_CONV = 10

_c_void_p = ctypes.c_void_p
_i64 = ctypes.c_int64
_int = ctypes.c_int
_size_t = ctypes.c_size_t


class EbsdWeights(ctypes.Structure):
    _fields_ = [
        ("conv_w", _c_void_p * N_CONV),
        ("conv_b", _c_void_p * N_CONV),
        ("mu_w", _c_void_p),
        ("mu_b", _c_void_p),
        ("logvar_w", _c_void_p),
        ("logvar_b", _c_void_p),
    ]


SYMBOLS = {
    "ebsd_abi_version": (_int, []),
    "ebsd_last_error": (ctypes.c_char_p, []),
    "ebsd_launch_count": (ctypes.c_uint64, []),
    "ebsd_quantize_crop": (_int, [_c_void_p, _int, _i64, _int, _int, _int, _int, _int, _int, _int, _int, _c_void_p,
                                  _c_void_p]),
    "ebsd_parse_angle_text": (_i64, [ctypes.c_char_p, _size_t, _c_void_p, _i64]),
    "ebsd_encoder_create": (_int, [ctypes.POINTER(_c_void_p), ctypes.POINTER(EbsdWeights), _int, _c_void_p]),
    "ebsd_encoder_destroy": (None, [_c_void_p]),
    "ebsd_encoder_workspace_bytes": (_size_t, [_c_void_p, _i64]),
    "ebsd_encoder_forward": (_int, [_c_void_p, _c_void_p, _int, _i64, _c_void_p, _c_void_p, _c_void_p, _size_t,
                                    _c_void_p]),
    "ebsd_encoder_block": (_int, [_c_void_p, _int, _int, _c_void_p, _c_void_p, _int, _int, _c_void_p, _c_void_p,
                                  _c_void_p]),
    "ebsd_normalize_rows": (_int, [_c_void_p, _i64, _int, _c_void_p]),
    "ebsd_topk_workspace_bytes": (_size_t, [_i64, _i64, _int]),
    "ebsd_topk": (_int, [_c_void_p, _i64, _i64, _c_void_p, _i64, _int, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                         _size_t, _c_void_p]),
    "ebsd_topk_merge": (_int, [_c_void_p, _c_void_p, _int, _i64, _int, _c_void_p, _c_void_p, _c_void_p, _c_void_p]),
    "ebsd_topk_pack": (_int, [_c_void_p, _c_void_p, _i64, _c_void_p, _c_void_p]),
    "ebsd_topk_merge_packed": (_int, [_c_void_p, _int, _i64, _int, _c_void_p, _c_void_p, _c_void_p, _c_void_p]),
    "ebsd_euler_to_quat": (_int, [_c_void_p, _i64, _c_void_p, _c_void_p]),
    "ebsd_ipf_color": (_int, [_c_void_p, _i64, _int, _c_void_p, _c_void_p]),
    "ebsd_consensus": (_int, [_c_void_p, _c_void_p, _i64, _i64, _c_void_p, _i64, _int, ctypes.c_double, _int, _int,
                              _int, _int, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                              _c_void_p]),
}

_lib = None
_lock = threading.Lock()


class NativeError(RuntimeError):
    """A libebsd_b200 call returned a non-zero status."""


def load() -> ctypes.CDLL:
    """Load libebsd_b200.so (once). Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(or `make -C ebsd_vae_b200/csrc`). ebsd_vae_b200 has no CPU / PyTorch fallback."
                )
            import torch  # noqa: F401  (brings libcudart.so.12 into the process before we resolve it)

            lib = ctypes.CDLL(LIB_PATH)
            for name, (restype, argtypes) in SYMBOLS.items():
                fn = getattr(lib, name)  # AttributeError if the symbol is missing
                fn.restype = restype
                fn.argtypes = argtypes
            _lib = lib
    return _lib


def last_error() -> str:
    return load().ebsd_last_error().decode("utf-8", "replace")


def check(status: int, what: str) -> None:
    if status != OK:
        raise NativeError(f"{what} failed with status {status}: {last_error()}")
