"""Short mesh target for ncu: one render of the 1 M-triangle scene at a reduced image size (development aid)."""
import argparse, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import drt_b200 as drt

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--spp", type=int, default=32)
ap.add_argument("--bounces", type=int, default=8)
ap.add_argument("--precision", default="f64")
ap.add_argument("--reps", type=int, default=1)
a = ap.parse_args()
scene = drt.tessellated_room(204, 362, width=a.size, height=a.size)
with drt.Context(0) as ctx:
    ctx.upload(scene)
    for _ in range(a.reps):
        img, grad, st = ctx.render(drt.make_opts(a.spp, a.bounces, 1.0,
                                                 precision=drt.F64 if a.precision == "f64" else drt.F32), stats=True)
    print(a.precision, st.kernel_ms, "ms", st.segments / st.kernel_ms / 1e3, "Msegments/s",
          st.bvh_nodes / st.segments, "nodes/seg", st.tri_tests / st.segments, "tests/seg")
