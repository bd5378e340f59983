"""Builds libdrtb.so (CUDA kernels + C ABI) in-tree for sm_100a with nvcc.

    python differentiable-renderer_b200/build.py [--force] [--verbose]

The .so lands in differentiable-renderer_b200/lib/ (git-ignored, shipped to the
GPU box by gpurun).  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB = PKG / "lib" / "libdrtb.so"
SOURCES = [CSRC / "drtb.cu"]
HEADERS = [CSRC / "path.cuh", CSRC / "real.cuh", CSRC / "rng.cuh", ROOT / "include" / "drtb.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
    "-cudart", "static",
    "-Xptxas", "-v",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(p.stat().st_mtime > t for p in SOURCES + HEADERS + [Path(__file__)])


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    LIB.parent.mkdir(parents=True, exist_ok=True)
    # the image exports CC=/opt/gcc/bin/gcc; nvcc wants the system host compiler
    cmd = [nvcc_path(), *NVCC_FLAGS, "-ccbin", "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++",
           "-I", str(ROOT / "include"), "-o", str(LIB), *map(str, SOURCES)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    (PKG / "lib" / "build.log").write_text(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libdrtb.so (see lib/build.log)")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv))
