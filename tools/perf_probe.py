"""Quick device-side timing probe (development aid, not the bench)."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import drt_b200 as drt

def main():
    cfgs = [(256, 256, 16, 1, 0.5), (256, 256, 16, 8, 1.0), (1024, 1024, 32, 8, 1.0)]
    if "--full" in sys.argv:
        cfgs.append((1024, 1024, 256, 8, 1.0))
    with drt.Context(0) as ctx:
        for prec, name in ((drt.F64, "f64"), (drt.F32, "f32")):
            print(name, "fma peak TFLOP/s:", round(ctx.fma_peak(prec), 2))
        for (W, H, spp, mb, ab) in cfgs:
            scene = drt.cornell_box(W, H)
            ctx.upload(scene)
            for prec, name in ((drt.F64, "f64"), (drt.F32, "f32")):
                best = None
                for rep in range(3):
                    img, grad, st = ctx.render(drt.make_opts(spp, mb, ab, precision=prec), stats=True)
                    if best is None or st.kernel_ms < best.kernel_ms:
                        best = st
                print(f"{W}x{H} spp={spp} mb={mb} p={ab} {name}: {best.kernel_ms:.2f} ms "
                      f"{best.paths / best.kernel_ms / 1e3:.1f} Mpaths/s seg/path={best.segments / best.paths:.2f} "
                      f"lit={best.lit_paths / best.paths:.3f} mean={img.reshape(-1,3).mean(0)} grad0={grad[0]}")

if __name__ == "__main__":
    main()
