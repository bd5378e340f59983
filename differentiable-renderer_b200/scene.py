"""Host-side scene description: the Python mirror of the object graph a
reference program builds at src/render.cpp:26-65 (parameters -> materials ->
shapes -> Scene<T> -> Camera<T>), flattened to the PODs of include/drtb.h.

Names and argument meaning follow the reference (`Sphere(center, radius, bxdf,
emitter)`, `Plane(normal, offset, bxdf, emitter)`, `Camera(w, h, vfov).look_at`)
so that parity tests read like a reference program.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import abi


@dataclass
class Param:
    """A differentiable RGB parameter: `Vector<T,3,true>(v, requires_grad=true)`
    (vector.hpp:228-234).  Copies of the handle alias one gradient accumulator
    (vector.hpp:317); here that is one index."""
    value: np.ndarray
    name: str = ""
    index: int = -1
    grad: np.ndarray = field(default_factory=lambda: np.zeros(3))


@dataclass
class DiffuseBxDF:
    """DiffuseBxDF(color), bxdf.hpp:56-83."""
    color: Param
    index: int = -1


@dataclass
class AreaEmitter:
    """AreaEmitter(emission), emitter.hpp:15-25."""
    emission: Param


@dataclass
class Sphere:
    """Sphere(center, radius, bxdf=None, emitter=None), shape.hpp:66-111."""
    center: Sequence[float]
    radius: float
    bxdf: Optional[DiffuseBxDF] = None
    emitter: Optional[AreaEmitter] = None


@dataclass
class Plane:
    """Plane(normal, offset, bxdf=None, emitter=None), shape.hpp:37-64.  The
    normal is kept exactly as given (never normalised, shape.hpp:58-59)."""
    normal: Sequence[float]
    offset: float
    bxdf: Optional[DiffuseBxDF] = None
    emitter: Optional[AreaEmitter] = None


def _normalize(v):
    v = np.asarray(v, dtype=np.float64)
    return v / math.sqrt(float(((0.0 + v[0] * v[0]) + v[1] * v[1]) + v[2] * v[2]))


def _cross(a, b):
    return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2],
                     a[0] * b[1] - a[1] * b[0]], dtype=np.float64)


class Camera:
    """Camera(width, height, vfov=1.3963, ...), camera.hpp:13-37."""

    def __init__(self, width: int, height: int, vfov: float = 1.3963,
                 eye=(0, 0, 0), forward=(0, 0, -1), right=(1, 0, 0), up=(0, 1, 0)):
        self.width, self.height, self.vfov = int(width), int(height), float(vfov)
        self.eye = np.asarray(eye, dtype=np.float64)
        self.forward = np.asarray(forward, dtype=np.float64)
        self.right = np.asarray(right, dtype=np.float64)
        self.up = np.asarray(up, dtype=np.float64)

    def look_at(self, eye, at, up=(0, 1, 0)):
        """camera.hpp:29-37."""
        self.eye = np.asarray(eye, dtype=np.float64)
        self.forward = _normalize(np.asarray(at, dtype=np.float64) - self.eye)
        self.right = _normalize(_cross(self.forward, np.asarray(up, dtype=np.float64)))
        self.up = _cross(self.right, self.forward)
        return self

    def to_abi(self) -> abi.Camera:
        c = abi.Camera()
        c.width, c.height, c.vfov = self.width, self.height, self.vfov
        for name in ("eye", "forward", "right", "up"):
            arr = getattr(c, name)
            for i in range(3):
                arr[i] = float(getattr(self, name)[i])
        return c


class SceneDesc:
    """`Scene<T>` (pathtracer.hpp:12-13): shapes in push_back order, which is
    the closest-hit tie-break priority (pathtracer.hpp:78-87)."""

    def __init__(self):
        self.shapes: List[object] = []
        self.params: List[Param] = []
        self.materials: List[DiffuseBxDF] = []
        self.camera: Optional[Camera] = None
        self._keep = None

    def push_back(self, shape):
        self.shapes.append(shape)
        return self

    # -- flattening ---------------------------------------------------------
    def _param_index(self, p: Param) -> int:
        for q in self.params:
            if q is p:
                return q.index
        p.index = len(self.params)
        self.params.append(p)
        return p.index

    def _material_index(self, m: DiffuseBxDF) -> int:
        for q in self.materials:
            if q is m:
                return q.index
        m.index = len(self.materials)
        self.materials.append(m)
        self._param_index(m.color)
        return m.index

    def flatten(self):
        prims = (abi.Prim * max(1, len(self.shapes)))()
        for i, s in enumerate(self.shapes):
            pr = prims[i]
            pr.material = self._material_index(s.bxdf) if s.bxdf is not None else -1
            pr.emission = self._param_index(s.emitter.emission) if s.emitter is not None else -1
            if isinstance(s, Sphere):
                pr.type = abi.SPHERE
                vals = [*s.center, s.radius]
            elif isinstance(s, Plane):
                pr.type = abi.PLANE
                vals = [*s.normal, s.offset]
            else:
                raise TypeError(f"unsupported shape {type(s).__name__}")
            for j in range(4):
                pr.v[j] = float(vals[j])
        mats = (abi.Material * max(1, len(self.materials)))()
        for m in self.materials:
            mats[m.index].type = abi.DIFFUSE
            mats[m.index].color = m.color.index
            mats[m.index].exponent = 0.0
        pvals = (C.c_double * max(3, 3 * len(self.params)))()
        for p in self.params:
            for c in range(3):
                pvals[3 * p.index + c] = float(p.value[c])
        sc = abi.Scene()
        sc.prims, sc.n_prims = prims, len(self.shapes)
        sc.materials, sc.n_materials = mats, len(self.materials)
        sc.params, sc.n_params = pvals, len(self.params)
        if self.camera is None:
            raise ValueError("scene has no camera")
        sc.camera = self.camera.to_abi()
        self._keep = (prims, mats, pvals)        # keep the arrays alive
        return sc

    def param_values(self) -> np.ndarray:
        return np.array([p.value for p in self.params], dtype=np.float64).reshape(-1, 3)


def cornell_box(width: int, height: int, *, red=(0.5, 0, 0), green=(0, 0.5, 0),
                white=(0.5, 0.5, 0.5), emission=(1, 1, 1)) -> SceneDesc:
    """The test scene of src/render.cpp:26-65, value for value, in scene order."""
    P = lambda v, n: Param(np.asarray(v, dtype=np.float64), n)
    red, green, white, emission = P(red, "red"), P(green, "green"), P(white, "white"), P(emission, "emission")
    diffuse_red, diffuse_green, diffuse_white = DiffuseBxDF(red), DiffuseBxDF(green), DiffuseBxDF(white)
    emitter = AreaEmitter(emission)
    sc = SceneDesc()
    sc.push_back(Sphere((0., 0., 3.), 1., diffuse_white))           # sphere_front
    sc.push_back(Sphere((-1., 1., 4.5), 1., diffuse_white))         # sphere_back
    sc.push_back(Plane((-1., 0., 0.), -3., diffuse_red))            # left_plane
    sc.push_back(Plane((1., 0., 0.1), -3., diffuse_green))          # right_plane (non-unit normal)
    sc.push_back(Plane((0., 0., -1.), -6., diffuse_white))          # back_plane
    sc.push_back(Plane((0., 0., 1.), 0., diffuse_white))            # front_plane
    sc.push_back(Plane((0., 1., 0.), -3., diffuse_white))           # ground_plane
    sc.push_back(Plane((0., -1., 0.), -3., diffuse_white))          # ceiling_plane
    sc.push_back(Sphere((0., 3., 3.), 1., None, emitter))           # light
    sc.camera = Camera(width, height).look_at((0, 0, 0), (0, 0, 1))
    # register parameters in declaration order red, green, white, emission
    for p in (red, green, white, emission):
        sc._param_index(p)
    return sc


def make_opts(spp: int, min_bounces: int = 1, absorb: float = 0.5, *, seed: int = 0,
              precision: int = abi.F64, flags: int = abi.FLAG_IMAGE | abi.FLAG_GRAD,
              shard_index: int = 0, shard_count: int = 1, band_rows: int = 8,
              max_depth: int = 0, seed_scale: float = 1.0, adjoint_seed: int = 0) -> abi.RenderOpts:
    """Defaults are the reference CLI's (-b 1 -p 0.5, args.hpp:44-59)."""
    o = abi.RenderOpts()
    o.spp, o.min_bounces, o.absorb, o.seed = spp, min_bounces, absorb, seed
    o.precision, o.flags = precision, flags
    o.shard_index, o.shard_count, o.band_rows = shard_index, shard_count, band_rows
    o.max_depth, o.seed_scale, o.adjoint_seed = max_depth, seed_scale, adjoint_seed
    return o
