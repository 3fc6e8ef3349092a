"""Builds shoulder_b200/libshoulder_b200.so for sm_100a with nvcc (in-tree, no JIT cache).

    python -m shoulder_b200.build [--force]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libshoulder_b200.so"
SOURCES = ["shb_kernels.cu", "shb_features.cu", "shb_meshio.cu", "shb_api.cu"]
HEADERS = [CSRC / "shb_common.cuh", PKG.parent / "include" / "shoulder_b200.h"]
NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--fmad=false",            # fp64 geometry must round like numpy (no contraction)
    "-shared", "-cudart", "static",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; the CUDA extension cannot be built")


def stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(p.stat().st_mtime > t for p in [CSRC / s for s in SOURCES] + HEADERS + [Path(__file__)])


def build(force: bool = False, verbose: bool = False) -> Path:
    """One object per source (compiled in parallel, only the stale ones), then the shared library."""
    if not force and not stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    extra = os.environ.get("SHB_NVCC_EXTRA", "").split()          # experiments only
    out = os.environ.get("SHB_BUILD_OUT", str(LIB))                   # experiments only: a second library beside the product
    objdir = PKG / "csrc" / ("_obj" if not extra else "_obj_x")
    objdir.mkdir(exist_ok=True)
    compile_flags = [f for f in NVCC_FLAGS if f not in ("-shared",)]
    i = compile_flags.index("-cudart")
    del compile_flags[i:i + 2]
    newest_hdr = max(p.stat().st_mtime for p in HEADERS + [Path(__file__)])

    def one(src: str):
        obj = objdir / (src[:-3] + ".o")
        s = CSRC / src
        if not force and not extra and obj.exists() and obj.stat().st_mtime > max(s.stat().st_mtime, newest_hdr):
            return obj, ""
        cmd = [nvcc_path(), *compile_flags, *extra, "-c", "-o", str(obj), str(s)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n" + res.stdout + res.stderr)
        return obj, res.stderr

    with ThreadPoolExecutor(len(SOURCES)) as ex:
        done = list(ex.map(one, SOURCES))
    if verbose:
        print("".join(d[1] for d in done))
    link = [nvcc_path(), "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
            "-o", out, *[str(d[0]) for d in done]]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
