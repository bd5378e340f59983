"""BVH build time of the 1 M-triangle scene, three uploads in one context (development aid): the first one
carries the module load and first-use costs of the build kernels."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import drt_b200 as drt
scene = drt.tessellated_room(204, 362, width=256, height=256)
with drt.Context(0) as ctx:
    for i in range(4):
        t0 = time.time(); ctx.upload(scene); wall = time.time() - t0
        print(f"upload {i}: device {ctx.mesh_build_ms:.2f} ms, wall (copy + build) {wall * 1e3:.1f} ms")
