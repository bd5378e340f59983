"""Short target for ncu: a few renders of one configuration (development aid)."""
import argparse, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import drt_b200 as drt

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=1024)
ap.add_argument("--spp", type=int, default=256)
ap.add_argument("--bounces", type=int, default=8)
ap.add_argument("--absorb", type=float, default=1.0)
ap.add_argument("--precision", default="f64")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
with drt.Context(0) as ctx:
    ctx.upload(drt.cornell_box(a.size, a.size))
    for _ in range(a.reps):
        img, grad, st = ctx.render(drt.make_opts(a.spp, a.bounces, a.absorb,
                                                 precision=drt.F64 if a.precision == "f64" else drt.F32), stats=True)
    print(a.precision, st.kernel_ms, "ms", st.paths / st.kernel_ms / 1e3, "Mpaths/s")
