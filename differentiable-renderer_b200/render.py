"""Thin Python host over the C ABI (include/drtb.h): the call a user makes in
place of the pixel loop of src/render.cpp:72-86.  Nothing here computes
radiance; every number comes out of libdrtb.so's CUDA kernels.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import abi
from .scene import SceneDesc, make_opts

_dp = C.POINTER(C.c_double)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_dp)


def shard_rows(height: int, shard_index: int = 0, shard_count: int = 1, band_rows: int = 8) -> int:
    return int(abi.load_library().drtb_shard_rows(height, shard_index, shard_count, band_rows))


def stream_draw(key: int, slot: int) -> int:
    return int(abi.load_library().drtb_stream_draw(key & (2**64 - 1), slot))


class Context:
    """Owns a drtb_ctx: one CUDA device, one uploaded scene."""

    def __init__(self, device: int = 0):
        self._lib = abi.load_library()
        h = C.c_void_p()
        rc = self._lib.drtb_create(int(device), C.byref(h))
        if rc != abi.OK:
            msg = self._lib.drtb_last_error(None)
            raise abi.DrtbError(rc, msg.decode() if msg else "drtb_create failed")
        self._h = h
        self.device = int(device)
        self.scene: Optional[SceneDesc] = None
        self._abi_scene = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.drtb_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int):
        if rc != abi.OK:
            msg = self._lib.drtb_last_error(self._h)
            raise abi.DrtbError(rc, msg.decode() if msg else "")

    # -- scene ----------------------------------------------------------------
    def upload(self, scene: SceneDesc):
        sc = scene.flatten()
        self._check(self._lib.drtb_scene_upload(self._h, C.byref(sc)))
        self.scene, self._abi_scene = scene, sc
        mesh = scene.flatten_mesh()
        if mesh is not None:                      # GPU LBVH build happens here
            self._check(self._lib.drtb_mesh_upload(self._h, C.byref(mesh)))
        return self

    def set_params(self, values: np.ndarray):
        v = np.ascontiguousarray(values, dtype=np.float64).reshape(-1, 3)
        self._check(self._lib.drtb_set_params(self._h, _ptr(v), v.shape[0]))

    def set_params_device(self, d_params: int, n_params: int, stream: int = 0):
        """drtb_set_params_device: n_params x 3 doubles already on the device, copied on `stream`."""
        self._check(self._lib.drtb_set_params_device(self._h, d_params, int(n_params), stream or None))

    # -- hot path, host buffers --------------------------------------------------
    def render(self, opts: abi.RenderOpts, seed_img: Optional[np.ndarray] = None,
               *, stats: bool = False):
        """Returns (img[rows,W,3], grad[P,3]) (+ abi.Stats when stats=True)."""
        assert self.scene is not None, "upload a scene first"
        cam = self.scene.camera
        rows = shard_rows(cam.height, opts.shard_index, max(1, opts.shard_count),
                          max(1, opts.band_rows))
        img = np.empty((rows, cam.width, 3), dtype=np.float64) if opts.flags & abi.FLAG_IMAGE else None
        grad = np.empty((self.scene.n_params, 3), dtype=np.float64) if opts.flags & abi.FLAG_GRAD else None
        if seed_img is not None:
            seed_img = np.ascontiguousarray(seed_img, dtype=np.float64)
            assert seed_img.shape == (rows, cam.width, 3)
        st = abi.Stats() if stats else None     # drtb_render turns the statistics on when it gets a drtb_stats
        self._check(self._lib.drtb_render(self._h, C.byref(opts), _ptr(seed_img), _ptr(img),
                                          _ptr(grad), C.byref(st) if stats else None))
        return (img, grad, st) if stats else (img, grad)

    def render_grad_image(self, opts: abi.RenderOpts, param, seed_img: Optional[np.ndarray] = None):
        """drtb_render_grad_image: (img, grad, grad_img[rows,W,3]) where grad_img is
        parameter `param`'s (a `Param` or an index) share of the gradient per pixel
        (README.md:138-145)."""
        assert self.scene is not None, "upload a scene first"
        k = param if isinstance(param, int) else param.index
        cam = self.scene.camera
        rows = shard_rows(cam.height, opts.shard_index, max(1, opts.shard_count), max(1, opts.band_rows))
        img = np.empty((rows, cam.width, 3), dtype=np.float64) if opts.flags & abi.FLAG_IMAGE else None
        grad = np.empty((self.scene.n_params, 3), dtype=np.float64)
        gimg = np.empty((rows, cam.width, 3), dtype=np.float64)
        if seed_img is not None:
            seed_img = np.ascontiguousarray(seed_img, dtype=np.float64)
            assert seed_img.shape == (rows, cam.width, 3)
        self._check(self._lib.drtb_render_grad_image(self._h, C.byref(opts), int(k), _ptr(seed_img), _ptr(img),
                                                     _ptr(grad), _ptr(gimg), None))
        return img, grad, gimg

    def render_host_ptrs(self, opts: abi.RenderOpts, seed_ptr: int, img_ptr: int, grad_ptr: int,
                         st: Optional[abi.Stats] = None):
        """drtb_render on caller-owned (e.g. pinned) host memory, by address."""
        self._check(self._lib.drtb_render(
            self._h, C.byref(opts), C.cast(seed_ptr, _dp) if seed_ptr else None,
            C.cast(img_ptr, _dp) if img_ptr else None, C.cast(grad_ptr, _dp) if grad_ptr else None,
            C.byref(st) if st is not None else None))

    def reserve(self, opts: abi.RenderOpts):
        """drtb_reserve: size scratch and load kernels for renders with these options."""
        self._check(self._lib.drtb_reserve(self._h, C.byref(opts)))

    # -- hot path, device buffers ------------------------------------------------
    def render_device(self, opts: abi.RenderOpts, d_seed_img: int, d_img: int, d_grad: int,
                      d_stats: int = 0, stream: int = 0):
        """Raw device addresses (e.g. torch.Tensor.data_ptr()) and a cudaStream_t
        handle; asynchronous."""
        self._check(self._lib.drtb_render_device(self._h, C.byref(opts), d_seed_img or None,
                                                 d_img or None, d_grad or None, d_stats or None,
                                                 stream or None))

    # -- multi-GPU: the image all-gather fused into the render ------------------------
    def set_image_peers(self, full_images):
        """drtb_set_image_peers: device addresses of the FULL image (H x W x 3 doubles) on
        every GPU of the job; sharded renders then store their pixels into all of them.
        An empty list switches it off."""
        n = len(full_images)
        arr = (C.c_void_p * max(n, 1))(*[int(p) for p in full_images])
        self._check(self._lib.drtb_set_image_peers(self._h, arr, n))

    def set_grad_peers(self, exchange, rank: int = 0):
        """drtb_set_grad_peers: device addresses of every rank's gradient exchange buffer (rank order); renders
        with FLAG_GRAD then return the sum over the ranks.  An empty list switches it off."""
        n = len(exchange)
        arr = (C.c_void_p * max(n, 1))(*[int(p) for p in exchange])
        self._check(self._lib.drtb_set_grad_peers(self._h, arr, n, int(rank)))

    def grad_exchange_bytes(self, n_ranks: int) -> int:
        return int(self._lib.drtb_grad_exchange_bytes(int(n_ranks), int(self.scene.n_params)))

    def ipc_alloc(self, nbytes: int):
        """(device address, 64-byte handle) of memory a peer process can map."""
        ptr = C.c_void_p()
        handle = (C.c_ubyte * abi.IPC_HANDLE_BYTES)()
        self._check(self._lib.drtb_ipc_alloc(self._h, int(nbytes), C.byref(ptr), handle))
        return int(ptr.value), bytes(handle)

    def ipc_open(self, handle: bytes) -> int:
        ptr = C.c_void_p()
        buf = (C.c_ubyte * abi.IPC_HANDLE_BYTES).from_buffer_copy(handle)
        self._check(self._lib.drtb_ipc_open(self._h, buf, C.byref(ptr)))
        return int(ptr.value)

    def ipc_close(self, ptr: int):
        self._check(self._lib.drtb_ipc_close(self._h, C.c_void_p(ptr)))

    def ipc_free(self, ptr: int):
        self._check(self._lib.drtb_ipc_free(self._h, C.c_void_p(ptr)))

    def trace_rays(self, opts: abi.RenderOpts, orig: np.ndarray, dirs: np.ndarray,
                   keys: np.ndarray, *, jac: bool = True) -> Tuple[np.ndarray, Optional[np.ndarray]]:
        """Pathtracer::trace on explicit rays (pathtracer.hpp:121-136)."""
        orig = np.ascontiguousarray(orig, dtype=np.float64).reshape(-1, 3)
        dirs = np.ascontiguousarray(dirs, dtype=np.float64).reshape(-1, 3)
        keys = np.ascontiguousarray(keys, dtype=np.uint64).reshape(-1)
        n = orig.shape[0]
        rad = np.empty((n, 3), dtype=np.float64)
        J = np.empty((n, self.scene.n_params, 3), dtype=np.float64) if jac else None
        self._check(self._lib.drtb_trace_rays(self._h, C.byref(opts), n, _ptr(orig), _ptr(dirs),
                                              keys.ctypes.data_as(C.POINTER(C.c_uint64)),
                                              _ptr(rad), _ptr(J)))
        return rad, J

    def fma_peak(self, precision: int) -> float:
        out = C.c_double()
        self._check(self._lib.drtb_fma_peak(self._h, precision, C.byref(out)))
        return out.value

    @property
    def mesh_build_ms(self) -> float:
        return float(self._lib.drtb_mesh_build_ms(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._lib.drtb_launch_count(self._h))


def render(scene: SceneDesc, spp: int, min_bounces: int = 1, absorb: float = 0.5, *,
           device: int = 0, **kw):
    """One-shot convenience: the whole of src/render.cpp:62-86 for `scene`.
    Adds the gradients into each `Param.grad` like VariableNode::backward does
    (vector.hpp:185-188) and returns (img, grad)."""
    with Context(device) as ctx:
        ctx.upload(scene)
        img, grad = ctx.render(make_opts(spp, min_bounces, absorb, **kw))
    if grad is not None:
        for p in scene.params:
            p.grad = p.grad + grad[p.index]
    return img, grad
