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
# one translation unit per kernel group, compiled in parallel (host.hpp says what each one exports)
SOURCES = [CSRC / "drtb.cu", CSRC / "render_f64.cu", CSRC / "render_f32.cu", CSRC / "mesh.cu", CSRC / "multi.cu"]
HEADERS = sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.hpp")) + [ROOT / "include" / "drtb.h"]
OBJ = PKG / "lib" / "obj"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    # Resident blocks of 128 threads per SM for the double kernels.  The first kernel of round 1 (141 registers
    # uncapped) gained from a cap of 96 (5 blocks) and nothing beyond; with today's kernel the picture is different:
    # 5 / 6 / 7 / 8 blocks (96 / 80 / 72 / 64 registers) -> 42.9 / 40.7 / 38.8 / 39.1 ms on the headline workload
    # (profiles/README.md, round 2).  72 registers are free of spills once the pixel accumulators live in shared
    # memory and the adjoint seed is fetched in the sweep (render_kernels.cuh); 64 registers spill 28 bytes.
    "-DDRTB_MIN_BLOCKS=7",
    # the float instantiation needs fewer registers: resident blocks 5 / 6 / 7 / 8 -> 34.6 / 33.1 / 31.2 / 31.1 ms;
    # 7 (72 registers) is the most that stays free of spills in every all-diffuse variant
    "-DDRTB_MIN_BLOCKS_F32=7",
    # mesh kernels are latency bound (BVH node fetches): more resident warps beat more registers
    "-DDRTB_MESH_MIN_BLOCKS=6",
    # the wavefront traversal: 7 resident blocks (72 registers) over 6: +1.5 % (profiles/README.md, round 2)
    "-DDRTB_WF_MIN_BLOCKS=7",
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


def _stale(obj: Path, src: Path) -> bool:
    if not obj.exists():
        return True
    t = obj.stat().st_mtime
    return any(p.stat().st_mtime > t for p in [src, Path(__file__)] + HEADERS)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu to an object (only the stale ones, in parallel), link libdrtb.so."""
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    OBJ.mkdir(parents=True, exist_ok=True)
    # the image exports CC=/opt/gcc/bin/gcc; nvcc wants the system host compiler
    ccbin = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
    nvcc = nvcc_path()

    def compile_one(src: Path):
        obj = OBJ / (src.stem + ".o")
        if not force and not _stale(obj, src):
            return src, None, 0
        cmd = [nvcc, *NVCC_FLAGS, "-ccbin", ccbin, "-I", str(ROOT / "include"), "-c", "-o", str(obj), str(src)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        (OBJ / (src.stem + ".log")).write_text(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        return src, r, r.returncode

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    failed = [(src, r) for src, r, rc in results if rc != 0]
    # lib/build.log: the ptxas reports of every object (tests read the register budgets from it)
    (PKG / "lib" / "build.log").write_text("".join((OBJ / (s.stem + ".log")).read_text() for s in SOURCES
                                                   if (OBJ / (s.stem + ".log")).exists()))
    for src, r in failed:
        sys.stderr.write(r.stdout + r.stderr)
    if failed:
        raise RuntimeError("nvcc failed on " + ", ".join(s.name for s, _ in failed) + " (see lib/obj/*.log)")
    if verbose:
        sys.stderr.write((PKG / "lib" / "build.log").read_text())
    cmd = [nvcc, "-shared", "-cudart", "static", "-ccbin", ccbin, "-o", str(LIB)] + [str(OBJ / (s.stem + ".o")) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("linking libdrtb.so failed")
    return LIB


def build_example(force: bool = False) -> Path:
    """examples/render.cpp (the reference application on the new include/drt
    headers) -> build/render, linked against libdrtb.so."""
    exe = ROOT / "build" / "render"
    src = ROOT / "examples" / "render.cpp"
    hdrs = list((ROOT / "include" / "drt").glob("*.hpp")) + [ROOT / "include" / "drtb.h", src, src.parent / "write.hpp"]
    if not force and exe.exists() and all(h.stat().st_mtime <= exe.stat().st_mtime for h in hdrs) \
            and LIB.stat().st_mtime <= exe.stat().st_mtime:
        return exe
    exe.parent.mkdir(parents=True, exist_ok=True)
    gxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
    cmd = [gxx, "-std=c++17", "-O2", "-Wall", "-Wextra", "-I", str(ROOT / "include"), str(src), "-o", str(exe),
           "-L", str(LIB.parent), "-ldrtb", "-Wl,-rpath,$ORIGIN/../differentiable-renderer_b200/lib"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed building examples/render.cpp")
    return exe


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv))
    print(build_example(force="--force" in sys.argv))
