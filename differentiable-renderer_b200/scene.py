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
class SpecularBxDF:
    """SpecularBxDF(color, exponent), bxdf.hpp:85-124: normalised Blinn-Phong
    lobe, half-vector sampling."""
    color: Param
    exponent: float
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


class TriangleMesh:
    """Indexed triangle mesh (NEW relative to the reference; semantics in
    include/drtb.h).  `albedo` is either one shared `Param` or an (m, 3) array
    of per-triangle differentiable albedos; `emissive` marks triangles that
    carry `emission` (a `Param`) -- those get a null BxDF, like the reference's
    light sphere (src/render.cpp:47)."""

    def __init__(self, vertices, indices, albedo, emission: Optional[Param] = None, emissive=None):
        self.vertices = np.ascontiguousarray(vertices, dtype=np.float64).reshape(-1, 3)
        self.indices = np.ascontiguousarray(indices, dtype=np.int32).reshape(-1, 3)
        m = self.indices.shape[0]
        self.albedo = albedo if isinstance(albedo, Param) else np.ascontiguousarray(albedo, dtype=np.float64).reshape(m, 3)
        self.emission = emission
        self.emissive = np.zeros(m, dtype=bool) if emissive is None else np.asarray(emissive, dtype=bool)
        self.param_base = -1                      # index of triangle 0's albedo in params[] (per-triangle mode)

    @property
    def n_triangles(self) -> int:
        return self.indices.shape[0]


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
        self.mesh: Optional[TriangleMesh] = None
        self.block_values: Optional[np.ndarray] = None    # per-triangle albedos appended after the object params
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

    def _material_index(self, m) -> int:
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
            spec = isinstance(m, SpecularBxDF)
            mats[m.index].type = abi.SPECULAR if spec else abi.DIFFUSE
            mats[m.index].color = m.color.index
            mats[m.index].exponent = float(m.exponent) if spec else 0.0
        mesh = self.mesh
        if mesh is not None:                      # register what the mesh references before sizing params[]
            if isinstance(mesh.albedo, Param):
                self._param_index(mesh.albedo)
            if mesh.emission is not None:
                self._param_index(mesh.emission)
        n_obj = len(self.params)
        self.block_values = None
        if mesh is not None and not isinstance(mesh.albedo, Param):
            mesh.param_base = n_obj
            self.block_values = mesh.albedo
        values = self.param_values()
        pvals = np.ascontiguousarray(values if values.size else np.zeros((1, 3)))
        sc = abi.Scene()
        sc.prims, sc.n_prims = prims, len(self.shapes)
        sc.materials, sc.n_materials = mats, len(self.materials)
        sc.params, sc.n_params = pvals.ctypes.data_as(C.POINTER(C.c_double)), values.shape[0]
        if self.camera is None:
            raise ValueError("scene has no camera")
        sc.camera = self.camera.to_abi()
        self._keep = (prims, mats, pvals)        # keep the arrays alive
        return sc

    def flatten_mesh(self) -> Optional[abi.Mesh]:
        """drtb_mesh for the attached TriangleMesh (call after flatten())."""
        mesh = self.mesh
        if mesh is None:
            return None
        m = mesh.n_triangles
        if isinstance(mesh.albedo, Param):
            color = np.full(m, mesh.albedo.index, dtype=np.int32)
        else:
            color = (mesh.param_base + np.arange(m)).astype(np.int32)
        emis = np.full(m, -1, dtype=np.int32)
        if mesh.emission is not None:
            emis[mesh.emissive] = mesh.emission.index
            color[mesh.emissive] = -1            # emitters have a null BxDF
        am = abi.Mesh()
        am.vertices = mesh.vertices.ctypes.data_as(C.POINTER(C.c_double)); am.n_vertices = mesh.vertices.shape[0]
        am.indices = mesh.indices.ctypes.data_as(C.POINTER(C.c_int32)); am.n_triangles = m
        am.color = color.ctypes.data_as(C.POINTER(C.c_int32)); am.emission = emis.ctypes.data_as(C.POINTER(C.c_int32))
        self._keep_mesh = (color, emis)
        return am

    @property
    def n_params(self) -> int:
        return len(self.params) + (0 if self.block_values is None else self.block_values.shape[0])

    def param_values(self) -> np.ndarray:
        obj = np.array([p.value for p in self.params], dtype=np.float64).reshape(-1, 3)
        if self.block_values is None:
            return obj
        return np.concatenate([obj, self.block_values], axis=0)


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


def specular_box(width: int, height: int, *, exponent_ball: float = 20.0, exponent_wall: float = 4.0,
                 gloss=(0.8, 0.7, 0.6)) -> SceneDesc:
    """The Cornell box of src/render.cpp:26-65 with the reference's SpecularBxDF
    (bxdf.hpp:85-124; src/render.cpp:35 builds one and never attaches it) on the
    front sphere and on the back wall, sharing one differentiable tint `gloss`.
    Everything else, including the scene order, is the Cornell box's."""
    sc = cornell_box(width, height)
    tint = Param(np.asarray(gloss, dtype=np.float64), "gloss")
    sc.shapes[0].bxdf = SpecularBxDF(tint, exponent_ball)          # sphere_front
    sc.shapes[4].bxdf = SpecularBxDF(tint, exponent_wall)          # back_plane, unit normal
    sc._param_index(tint)
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


def tessellated_room(wall_grid: int = 8, sphere_segments: int = 16, *, width: int = 256, height: int = 256,
                     seed: int = 1, displacement: float = 0.06) -> SceneDesc:
    """Config 4 of BASELINE.json, procedurally: a closed room [-3,3] x [-3,3] x [0,6]
    (the Cornell box's extent) whose six walls are `wall_grid` x `wall_grid` quads,
    a hash-displaced tessellated sphere, and an emissive quad pair under the ceiling.
    Every non-emissive triangle has its OWN differentiable albedo ~ U(0.2, 0.8).
    12 * wall_grid^2 + 4 * sphere_segments^2 - 4 * sphere_segments + 2 triangles;
    deterministic in `seed`.  Walls face inward, the sphere outward."""
    rng = np.random.default_rng(seed)
    verts, tris = [], []

    def quad_grid(origin, du, dv, n):            # n x n quads; normal = normalize(cross(du, dv))
        base = len(verts)
        for j in range(n + 1):
            for i in range(n + 1):
                verts.append(origin + du * (i / n) + dv * (j / n))
        for j in range(n):
            for i in range(n):
                a = base + j * (n + 1) + i
                b, c, d = a + 1, a + n + 1, a + n + 2
                tris.append((a, b, d)); tris.append((a, d, c))

    v = lambda *x: np.array(x, dtype=np.float64)
    g = wall_grid
    quad_grid(v(-3, -3, 6), v(6, 0, 0), v(0, 6, 0), g)       # back   z=6, normal -z ... see orientation fix below
    quad_grid(v(-3, -3, 0), v(0, 6, 0), v(6, 0, 0), g)       # front  z=0
    quad_grid(v(-3, -3, 0), v(6, 0, 0), v(0, 0, 6), g)       # floor  y=-3
    quad_grid(v(-3, 3, 0), v(0, 0, 6), v(6, 0, 0), g)        # ceiling y=3
    quad_grid(v(-3, -3, 0), v(0, 0, 6), v(0, 6, 0), g)       # x=-3
    quad_grid(v(3, -3, 0), v(0, 6, 0), v(0, 0, 6), g)        # x=+3
    n_wall = len(tris)
    # sphere: stacks x slices, outward winding, radial hash displacement
    s = sphere_segments
    c0, r0 = v(0.4, -1.2, 3.6), 1.4
    base = len(verts)
    disp = 1.0 + displacement * (rng.random((s + 1, 2 * s)) * 2 - 1)
    for j in range(s + 1):
        th = math.pi * j / s
        for i in range(2 * s):
            ph = 2 * math.pi * i / (2 * s)
            rr = r0 * (disp[j, i] if 0 < j < s else 1.0)
            verts.append(c0 + rr * v(math.sin(th) * math.cos(ph), math.cos(th), math.sin(th) * math.sin(ph)))
    for j in range(s):
        for i in range(2 * s):
            a = base + j * 2 * s + i
            b = base + j * 2 * s + (i + 1) % (2 * s)
            c = a + 2 * s
            d = b + 2 * s
            if j > 0: tris.append((a, b, c))
            if j < s - 1: tris.append((b, d, c))
    n_geo = len(tris)
    # light: a quad just under the ceiling, facing down
    quad_grid(v(-1, 2.9, 2), v(2, 0, 0), v(0, 0, 2), 1)
    verts = np.array(verts); tris = np.array(tris, dtype=np.int32)
    # orientation: walls + light must face the room centre, the sphere away from its centre
    e1 = verts[tris[:, 1]] - verts[tris[:, 0]]; e2 = verts[tris[:, 2]] - verts[tris[:, 0]]
    nrm = np.cross(e1, e2); cen = verts[tris].mean(1)
    room_c = v(0, 0, 3)
    want = np.where(np.arange(len(tris))[:, None] < n_wall, room_c - cen, cen - c0)
    want[n_geo:] = room_c - cen[n_geo:]
    flip = (nrm * want).sum(1) < 0
    tris[flip] = tris[flip][:, [0, 2, 1]]
    emissive = np.zeros(len(tris), dtype=bool); emissive[n_geo:] = True
    albedo = rng.uniform(0.2, 0.8, size=(len(tris), 3))
    sc = SceneDesc()
    sc.mesh = TriangleMesh(verts, tris, albedo, emission=Param(np.array([12.0, 12.0, 12.0]), "emission"), emissive=emissive)
    sc.camera = Camera(width, height).look_at((0, 0, 0.2), (0, -0.3, 3.5))
    return sc
