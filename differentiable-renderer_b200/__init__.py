"""differentiable-renderer_b200 — B200-native hot path of
thalesfm/differentiable-renderer: forward path tracer + radiative-backprop
adjoint as hand-written sm_100a CUDA behind the C ABI of include/drtb.h.

The directory name carries a hyphen (it is the name the build contract asks
for); import it through the `drt_b200` shim at the repo root.
"""
from . import abi
from .abi import (DrtbError, DrtbLibraryMissing, F32, F64, MIXED, FLAG_GRAD, FLAG_IMAGE,
                  FLAG_NO_BVH, FLAG_STATS, FLAG_DETERMINISTIC, load_library)
from .scene import (AreaEmitter, Camera, DiffuseBxDF, Param, Plane, SceneDesc, SpecularBxDF, Sphere, TriangleMesh,
                    cornell_box, make_opts, specular_box, tessellated_room)
from .render import Context, render, shard_rows, stream_draw
