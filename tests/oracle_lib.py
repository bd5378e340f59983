"""Loader for the CHECKERS under oracle/ (test infrastructure only).

`restate`  = oracle/liboracle_restate.so  (plain-C restatement, always buildable)
`ref`      = oracle/_ref/libdrt_ref.so    (unmodified reference headers; built
             where /root/reference exists, shipped prebuilt to the GPU box)
"""
from __future__ import annotations

import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import drt_b200 as drt  # noqa: E402
from drt_b200 import abi  # noqa: E402

ORACLE_DIR = ROOT / "oracle"
RESTATE_SO = ORACLE_DIR / "liboracle_restate.so"
REF_SO = ORACLE_DIR / "_ref" / "libdrt_ref.so"
_dp = C.POINTER(C.c_double)


def _make(target: str):
    subprocess.run(["make", "-C", str(ORACLE_DIR), target], check=True, capture_output=True)


def load_restate() -> C.CDLL:
    if not RESTATE_SO.exists():
        _make("restate")
    lib = C.CDLL(str(RESTATE_SO))
    lib.drt_oracle_render.restype = C.c_int
    lib.drt_oracle_render.argtypes = [C.POINTER(abi.Scene), C.POINTER(abi.RenderOpts), _dp, _dp, _dp,
                                      C.c_int, C.POINTER(abi.Stats)]
    lib.drt_oracle_render_mesh.restype = C.c_int
    lib.drt_oracle_render_mesh.argtypes = [C.POINTER(abi.Scene), C.POINTER(abi.Mesh), C.POINTER(abi.RenderOpts), _dp, _dp,
                                           _dp, C.c_int, C.POINTER(abi.Stats)]
    lib.drt_oracle_render_gimg.restype = C.c_int
    lib.drt_oracle_render_gimg.argtypes = [C.POINTER(abi.Scene), C.POINTER(abi.Mesh), C.POINTER(abi.RenderOpts), _dp, _dp,
                                           _dp, C.c_int, _dp, C.c_int, C.POINTER(abi.Stats)]
    lib.drt_oracle_trace_rays_mesh.restype = C.c_int
    lib.drt_oracle_trace_rays_mesh.argtypes = [C.POINTER(abi.Scene), C.POINTER(abi.Mesh), C.POINTER(abi.RenderOpts),
                                               C.c_int64, _dp, _dp, C.POINTER(C.c_uint64), _dp, _dp]
    lib.drt_oracle_trace_rays.restype = C.c_int
    lib.drt_oracle_trace_rays.argtypes = [C.POINTER(abi.Scene), C.POINTER(abi.RenderOpts), C.c_int64,
                                          _dp, _dp, C.POINTER(C.c_uint64), _dp, _dp]
    lib.drt_oracle_stream_draw.restype = C.c_uint32
    lib.drt_oracle_stream_draw.argtypes = [C.c_uint64, C.c_uint32]
    lib.drt_oracle_max_threads.restype = C.c_int
    return lib


def have_ref() -> bool:
    if REF_SO.exists():
        return True
    if Path("/root/reference/include/drt").is_dir():
        _make("ref")
        return REF_SO.exists()
    return False


def load_ref() -> C.CDLL:
    assert have_ref(), "oracle/_ref/libdrt_ref.so missing and /root/reference absent"
    lib = C.CDLL(str(REF_SO))
    lib.drt_ref_render.restype = C.c_int
    lib.drt_ref_render.argtypes = [C.POINTER(abi.Scene), C.POINTER(abi.RenderOpts), _dp, _dp, _dp,
                                   C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    lib.drt_ref_render_mesh.restype = C.c_int
    lib.drt_ref_render_mesh.argtypes = [C.POINTER(abi.Scene), C.POINTER(abi.Mesh), C.POINTER(abi.RenderOpts), _dp, _dp,
                                        _dp, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    lib.drt_ref_render_gimg.restype = C.c_int
    lib.drt_ref_render_gimg.argtypes = [C.POINTER(abi.Scene), C.POINTER(abi.Mesh), C.POINTER(abi.RenderOpts), _dp, _dp,
                                        _dp, C.c_int, _dp, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    lib.drt_ref_trace_ray.restype = C.c_int
    lib.drt_ref_trace_ray.argtypes = [C.POINTER(abi.Scene), C.POINTER(abi.RenderOpts), _dp, _dp,
                                      C.c_uint64, _dp, _dp]
    lib.drt_ref_max_threads.restype = C.c_int
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _rows(scene, opts):
    H = scene.camera.height
    if opts.shard_count <= 1:
        return H
    band = max(1, opts.band_rows)
    return sum(1 for y in range(H) if (y // band) % opts.shard_count == opts.shard_index)


def restate_render(scene, opts, seed_img=None, threads=1, want_stats=False, grad_image_of=None):
    """grad_image_of = parameter index: also returns the per-pixel gradient image (last)."""
    lib = load_restate()
    sc = scene.flatten()
    mesh = scene.flatten_mesh()
    rows = _rows(scene, opts)
    img = np.zeros((rows, scene.camera.width, 3))
    grad = np.zeros((scene.n_params, 3))
    gimg = np.zeros((rows, scene.camera.width, 3)) if grad_image_of is not None else None
    st = abi.Stats()
    if seed_img is not None:
        seed_img = np.ascontiguousarray(seed_img, dtype=np.float64)
    rc = lib.drt_oracle_render_gimg(C.byref(sc), C.byref(mesh) if mesh is not None else None, C.byref(opts),
                                    _ptr(seed_img), _ptr(img), _ptr(grad),
                                    -1 if grad_image_of is None else int(grad_image_of), _ptr(gimg), threads, C.byref(st))
    assert rc == 0
    out = (img, grad, st) if want_stats else (img, grad)
    return out + (gimg,) if gimg is not None else out


def ref_render(scene, opts, seed_img=None, threads=1, rand_mode=0, want_draws=False, grad_image_of=None):
    lib = load_ref()
    sc = scene.flatten()
    mesh = scene.flatten_mesh()
    rows = _rows(scene, opts)
    img = np.zeros((rows, scene.camera.width, 3))
    grad = np.zeros((scene.n_params, 3))
    gimg = np.zeros((rows, scene.camera.width, 3)) if grad_image_of is not None else None
    draws = C.c_uint64()
    if seed_img is not None:
        seed_img = np.ascontiguousarray(seed_img, dtype=np.float64)
    rc = lib.drt_ref_render_gimg(C.byref(sc), C.byref(mesh) if mesh is not None else None, C.byref(opts),
                                 _ptr(seed_img), _ptr(img), _ptr(grad),
                                 -1 if grad_image_of is None else int(grad_image_of), _ptr(gimg), threads, rand_mode,
                                 C.byref(draws))
    assert rc == 0
    out = (img, grad, draws.value) if want_draws else (img, grad)
    return out + (gimg,) if gimg is not None else out


def rel_err(a: np.ndarray, b: np.ndarray, floor: float | None = None) -> np.ndarray:
    """|a-b| / max(|b|, floor); floor defaults to 1e-6 * mean|b| (SURVEY §7.3:
    exact zeros must match exactly-ish, not blow the ratio up)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if floor is None:
        floor = 1e-6 * float(np.mean(np.abs(b))) + 1e-300
    return np.abs(a - b) / np.maximum(np.abs(b), floor)
