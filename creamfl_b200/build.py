"""Build libcreamfl_b200.so in-tree with nvcc for sm_100a.

The shared library is the whole product boundary (plain C ABI, see include/creamfl_b200.h); Python reaches it
through ctypes (creamfl_b200/_lib.py).  nvcc cross-compiles without a GPU, so this runs on the CPU build box and
the resulting .so travels to the GPU box with the tree.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OBJ_DIR = HERE / "build"
LIB_PATH = HERE / "libcreamfl_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _digest(src: Path) -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(src.read_bytes())
    for hdr in sorted(CSRC.glob("*.cuh")) + sorted((HERE.parent / "include").glob("*.h")):
        h.update(hdr.read_bytes())
    return h.hexdigest()


def _compile_one(src: Path, verbose: bool) -> Path:
    obj = OBJ_DIR / (src.stem + ".o")
    stamp = OBJ_DIR / (src.stem + ".sha")
    dig = _digest(src)
    if obj.exists() and stamp.exists() and stamp.read_text() == dig:
        return obj
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    (OBJ_DIR / (src.stem + ".ptxas.log")).write_text(res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{res.stderr}\n{res.stdout}")
    if verbose:
        sys.stderr.write(res.stderr)
    stamp.write_text(dig)
    return obj


def build(verbose: bool = False, force: bool = False) -> Path:
    """Compile every .cu under csrc/ and link libcreamfl_b200.so next to this file."""
    OBJ_DIR.mkdir(exist_ok=True)
    if force:
        for f in OBJ_DIR.glob("*.sha"):
            f.unlink()
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile_one(s, verbose), srcs))
    newest = max(o.stat().st_mtime for o in objs)
    if not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < newest:
        cmd = [_nvcc(), "-shared", "-o", str(LIB_PATH), *map(str, objs),
               "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print(p)
