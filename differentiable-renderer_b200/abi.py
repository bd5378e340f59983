"""ctypes mirror of include/drtb.h (the C ABI) and the loader for libdrtb.so.

There is deliberately no fallback: if the CUDA library has not been built the
import of the product path fails loudly (`DrtbLibraryMissing`).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "lib" / "libdrtb.so"

ABI_VERSION = 3

OK, ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_UNSUPPORTED, ERR_NOMEM = 0, -1, -2, -3, -4, -5
SPHERE, PLANE = 0, 1
DIFFUSE, SPECULAR = 0, 1
F64, F32, MIXED = 0, 1, 2
FLAG_IMAGE, FLAG_GRAD, FLAG_STATS, FLAG_NO_BVH, FLAG_DETERMINISTIC = 1, 2, 4, 8, 16

STREAM_KEY_MUL = 0x9E3779B97F4A7C15
IPC_HANDLE_BYTES = 64
MAX_PEERS = 8


class Prim(C.Structure):
    _fields_ = [("type", C.c_int32), ("material", C.c_int32), ("emission", C.c_int32),
                ("reserved", C.c_int32), ("v", C.c_double * 4)]


class Material(C.Structure):
    _fields_ = [("type", C.c_int32), ("color", C.c_int32), ("exponent", C.c_double)]


class Camera(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("vfov", C.c_double),
                ("eye", C.c_double * 3), ("forward", C.c_double * 3),
                ("right", C.c_double * 3), ("up", C.c_double * 3)]


class Scene(C.Structure):
    _fields_ = [("prims", C.POINTER(Prim)), ("n_prims", C.c_int32),
                ("materials", C.POINTER(Material)), ("n_materials", C.c_int32),
                ("params", C.POINTER(C.c_double)), ("n_params", C.c_int32),
                ("camera", Camera)]


class RenderOpts(C.Structure):
    _fields_ = [("spp", C.c_int32), ("min_bounces", C.c_int32), ("absorb", C.c_double),
                ("seed", C.c_uint64), ("precision", C.c_int32), ("flags", C.c_uint32),
                ("shard_index", C.c_int32), ("shard_count", C.c_int32),
                ("band_rows", C.c_int32), ("max_depth", C.c_int32),
                ("seed_scale", C.c_double), ("adjoint_seed", C.c_uint64)]


class Stats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("segments", C.c_uint64), ("lit_paths", C.c_uint64),
                ("truncated_paths", C.c_uint64), ("retraced_paths", C.c_uint64),
                ("bvh_nodes", C.c_uint64), ("tri_tests", C.c_uint64),
                ("kernel_ms", C.c_double)]


class Mesh(C.Structure):
    _fields_ = [("vertices", C.POINTER(C.c_double)), ("n_vertices", C.c_int64),
                ("indices", C.POINTER(C.c_int32)), ("n_triangles", C.c_int64),
                ("color", C.POINTER(C.c_int32)), ("emission", C.POINTER(C.c_int32))]


# every symbol include/drtb.h declares: (name, restype, argtypes)
_dp = C.POINTER(C.c_double)
SYMBOLS = [
    ("drtb_abi_version", C.c_int, []),
    ("drtb_device_count", C.c_int, []),
    ("drtb_struct_size", C.c_size_t, [C.c_int]),
    ("drtb_create", C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    ("drtb_destroy", None, [C.c_void_p]),
    ("drtb_last_error", C.c_char_p, [C.c_void_p]),
    ("drtb_scene_upload", C.c_int, [C.c_void_p, C.POINTER(Scene)]),
    ("drtb_mesh_upload", C.c_int, [C.c_void_p, C.POINTER(Mesh)]),
    ("drtb_mesh_build_ms", C.c_double, [C.c_void_p]),
    ("drtb_set_params", C.c_int, [C.c_void_p, _dp, C.c_int32]),
    ("drtb_set_params_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    ("drtb_shard_rows", C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    ("drtb_render", C.c_int, [C.c_void_p, C.POINTER(RenderOpts), _dp, _dp, _dp, C.POINTER(Stats)]),
    ("drtb_reserve", C.c_int, [C.c_void_p, C.POINTER(RenderOpts)]),
    ("drtb_render_device", C.c_int, [C.c_void_p, C.POINTER(RenderOpts), C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p]),
    ("drtb_render_grad_image", C.c_int, [C.c_void_p, C.POINTER(RenderOpts), C.c_int32, _dp, _dp, _dp, _dp,
                                         C.POINTER(Stats)]),
    ("drtb_render_grad_image_device", C.c_int, [C.c_void_p, C.POINTER(RenderOpts), C.c_int32, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("drtb_set_image_peers", C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32]),
    ("drtb_grad_exchange_bytes", C.c_size_t, [C.c_int32, C.c_int32]),
    ("drtb_set_grad_peers", C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.c_int32]),
    ("drtb_host_alloc", C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    ("drtb_host_free", C.c_int, [C.c_void_p]),
    ("drtb_ipc_alloc", C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.c_void_p]),
    ("drtb_ipc_open", C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    ("drtb_ipc_close", C.c_int, [C.c_void_p, C.c_void_p]),
    ("drtb_ipc_free", C.c_int, [C.c_void_p, C.c_void_p]),
    ("drtb_multi_create", C.c_int, [C.POINTER(C.c_int), C.c_int32, C.POINTER(C.c_void_p)]),
    ("drtb_multi_destroy", None, [C.c_void_p]),
    ("drtb_multi_last_error", C.c_char_p, [C.c_void_p]),
    ("drtb_multi_device_count", C.c_int32, [C.c_void_p]),
    ("drtb_multi_scene_upload", C.c_int, [C.c_void_p, C.POINTER(Scene)]),
    ("drtb_multi_mesh_upload", C.c_int, [C.c_void_p, C.POINTER(Mesh)]),
    ("drtb_multi_set_params", C.c_int, [C.c_void_p, _dp, C.c_int32]),
    ("drtb_multi_render", C.c_int, [C.c_void_p, C.POINTER(RenderOpts), _dp, _dp, _dp, C.POINTER(Stats)]),
    ("drtb_trace_rays", C.c_int, [C.c_void_p, C.POINTER(RenderOpts), C.c_int64, _dp, _dp,
                                  C.POINTER(C.c_uint64), _dp, _dp]),
    ("drtb_fma_peak", C.c_int, [C.c_void_p, C.c_int32, _dp]),
    ("drtb_chunk_plan", C.c_int, [C.c_int64, C.c_int32, C.c_int64, C.c_int32, C.POINTER(C.c_int64)]),
    ("drtb_launch_count", C.c_uint64, [C.c_void_p]),
    ("drtb_stream_draw", C.c_uint32, [C.c_uint64, C.c_uint32]),
]


class DrtbLibraryMissing(ImportError):
    pass


class DrtbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"drtb error {code}: {msg}")
        self.code = code


_lib = None


def load_library(path: os.PathLike | None = None) -> C.CDLL:
    """dlopen libdrtb.so and bind every symbol of include/drtb.h."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    if path is None and os.environ.get("DRTB_LIB"):          # development: A/B a variant build
        path = os.environ["DRTB_LIB"]
    p = Path(path) if path else LIB_PATH
    if not p.exists():
        raise DrtbLibraryMissing(
            f"{p} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the render path)")
    lib = C.CDLL(str(p))
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)          # AttributeError == missing export
        fn.restype = res
        fn.argtypes = args
    if lib.drtb_abi_version() != ABI_VERSION:
        raise DrtbLibraryMissing(f"{p}: ABI {lib.drtb_abi_version()} != {ABI_VERSION}")
    for which, st in enumerate((Prim, Material, Camera, Scene, RenderOpts, Stats, Mesh)):
        if lib.drtb_struct_size(which) != C.sizeof(st):
            raise DrtbLibraryMissing(f"{p}: layout of {st.__name__} differs from the ctypes mirror")
    if path is None:
        _lib = lib
    return lib
