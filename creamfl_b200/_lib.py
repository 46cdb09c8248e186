"""ctypes binding of libcreamfl_b200.so (C ABI declared in include/creamfl_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, a RuntimeError carrying the
library's own message is raised.
"""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

import os

# CREAMFL_LIB: development aid (A/B timing of two builds inside one GPU lease); the product loads the in-tree library
_LIB_PATH = Path(os.environ.get("CREAMFL_LIB") or Path(__file__).resolve().parent / "libcreamfl_b200.so")
_lib = None

vp, i32, i64, f32, sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t

_HEADER = Path(__file__).resolve().parent.parent / "include" / "creamfl_b200.h"
_SCALARS = {"int": i32, "int64_t": i64, "float": f32, "size_t": sz, "int32_t": i32}
_PROTO = re.compile(r"^(size_t|int|const char\*)\s+(creamfl_\w+)\s*\(([^;{]*)\)\s*;", re.M | re.S)


def _parse_header() -> dict:
    """name -> (restype, argtypes), read from include/creamfl_b200.h so the binding cannot drift from the ABI.
    Every pointer (of any pointee type) travels as c_void_p; scalars map one to one."""
    text = re.sub(r"/\*.*?\*/", "", _HEADER.read_text(), flags=re.S)
    sigs = {}
    for ret, name, args in _PROTO.findall(text):
        res = {"size_t": sz, "int": i32, "const char*": C.c_char_p}[ret]
        argtypes = []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(vp)
                else:
                    argtypes.append(_SCALARS[a.replace("const ", "").split()[0]])
        sigs[name] = (res, argtypes)
    return sigs


_SIGNATURES = _parse_header()


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
