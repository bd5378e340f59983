"""Several analytic-scene configurations on one build (development aid, used with DRTB_LIB for A/B runs):
fixed-length paths at B = 1..16, Russian roulette at the reference defaults, small spp, many parameters."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import drt_b200 as drt

cfgs = [(1024, 256, 8, 1.0), (1024, 256, 1, 1.0), (1024, 256, 4, 1.0), (1024, 128, 16, 1.0), (1024, 256, 1, 0.5), (1024, 16, 1, 0.5),
        (1024, 16, 8, 1.0), (256, 16, 8, 1.0)]
out = []
with drt.Context(0) as ctx:
    for prec, name in ((drt.F64, "f64"), (drt.F32, "f32")):
        for (size, spp, mb, ab) in cfgs:
            ctx.upload(drt.cornell_box(size, size))
            best = None
            for _ in range(3):
                img, grad, st = ctx.render(drt.make_opts(spp, mb, ab, precision=prec), stats=True)
                best = st.kernel_ms if best is None else min(best, st.kernel_ms)
            out.append(f"{name} {size}^2 spp={spp} b={mb} p={ab}: {best:.3f} ms")
print("\n".join(out))
