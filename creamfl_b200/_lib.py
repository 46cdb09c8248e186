"""ctypes binding of libcreamfl_b200.so (C ABI declared in include/creamfl_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, a RuntimeError carrying the
library's own message is raised.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "libcreamfl_b200.so"
_lib = None

vp, i32, i64, f32, sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t

# name -> (restype, argtypes); mirrors include/creamfl_b200.h one to one
_SIGNATURES = {
    "creamfl_last_error": (C.c_char_p, []),
    "creamfl_abi_version": (i32, []),
    "creamfl_gemm_bf16": (i32, [vp, i64, i32, vp, i64, i32, i32, i32, i32, vp, i64, i32, vp, vp, i32, f32, vp, i64,
                                i32, vp, i64, i32, vp]),
    "creamfl_rowlse_workspace_bytes": (sz, [i32, i32]),
    "creamfl_infonce_fwd": (i32, [vp, vp, vp, i32, i32, i32, f32, vp, vp, vp, vp, sz, vp]),
    "creamfl_infonce_bwd_workspace_bytes": (sz, [i32, i32]),
    "creamfl_infonce_bwd": (i32, [vp, vp, vp, vp, i32, i32, i32, f32, vp, vp, vp, sz, vp]),
    "creamfl_conw_score": (i32, [vp, vp, i32, i32, vp, vp, sz, vp]),
    "creamfl_conw_reduce": (i32, [C.POINTER(vp), vp, i32, i32, i32, vp, vp, vp]),
    "creamfl_pcme_workspace_bytes": (sz, [i32]),
    "creamfl_pcme_fwd": (i32, [vp, vp, i32, i32, vp, vp, vp, vp, vp, sz, vp]),
    "creamfl_pcme_bwd": (i32, [vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]),
    "creamfl_moon_fwd": (i32, [vp, vp, vp, vp, i32, i32, f32, f32, vp, vp, vp, vp]),
    "creamfl_moon_bwd": (i32, [vp, vp, vp, vp, vp, i32, i32, vp, vp]),
    "creamfl_mse_workspace_bytes": (sz, []),
    "creamfl_mse_gather_fwd": (i32, [vp, vp, vp, i32, i32, vp, vp, sz, vp]),
    "creamfl_mse_gather_bwd": (i32, [vp, vp, vp, vp, i32, i32, vp, vp]),
    "creamfl_l2norm_fwd": (i32, [vp, i32, i32, vp, vp, vp, vp]),
    "creamfl_l2norm_bwd": (i32, [vp, vp, vp, i32, i32, vp, vp]),
    "creamfl_cast_f32_bf16": (i32, [vp, i64, vp, vp]),
    "creamfl_recall_workspace_bytes": (sz, [i32]),
    "creamfl_recall_ranks": (i32, [vp, vp, vp, vp, i32, i32, i32, vp, vp, sz, vp]),
}


def lib_path() -> Path:
    return _LIB_PATH


def load():
    """Load (once) and return the ctypes handle.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise RuntimeError(
            f"{_LIB_PATH} is missing - build it with `python -m creamfl_b200.build` "
            "(there is no CPU or PyTorch fallback for the creamfl_b200 hot path)")
    lib = C.CDLL(str(_LIB_PATH))
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def exported_symbols() -> list[str]:
    return sorted(_SIGNATURES)


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().creamfl_last_error()
        raise RuntimeError(f"{what} failed with status {status}: {msg.decode() if msg else '?'}")
