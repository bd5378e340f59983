"""How the analytic megakernel behaves as the number of differentiable parameters grows
(development aid): N spheres with their own albedo in a box, 512^2, 64 spp, 6 bounces."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import drt_b200 as drt


def scene(n_spheres, size):
    P = lambda v, n: drt.Param(np.asarray(v, dtype=np.float64), n)
    rng = np.random.default_rng(1)
    sc = drt.SceneDesc()
    for i in range(n_spheres):
        sc.push_back(drt.Sphere((-2.4 + 4.8 * (i % 6) / 5, -2.2 + 1.1 * (i // 6), 3.0 + 0.4 * (i % 3)), 0.4,
                                drt.DiffuseBxDF(P(rng.uniform(0.2, 0.9, 3), f"c{i}"))))
    wall = drt.DiffuseBxDF(P((0.5, 0.5, 0.5), "wall"))
    for nrm, off in (((-1., 0., 0.), -3.), ((1., 0., 0.), -3.), ((0., 0., -1.), -6.), ((0., 0., 1.), 0.),
                     ((0., 1., 0.), -3.), ((0., -1., 0.), -3.)):
        sc.push_back(drt.Plane(nrm, off, wall))
    sc.push_back(drt.Sphere((0., 3., 3.), 1., None, drt.AreaEmitter(P((1, 1, 1), "emission"))))
    sc.camera = drt.Camera(size, size).look_at((0, 0, 0), (0, 0, 1))
    return sc


def main():
    size, spp = 512, 64
    with drt.Context(0) as ctx:
        for n in (2, 6, 7, 12, 20):
            sc = scene(n, size)
            ctx.upload(sc)
            for prec, name in ((drt.F64, "f64"),):
                best = None
                for _ in range(3):
                    _, grad, st = ctx.render(drt.make_opts(spp, 6, 1.0, precision=prec), stats=True)
                    best = st if best is None or st.kernel_ms < best.kernel_ms else best
                _, _, st_nograd = ctx.render(drt.make_opts(spp, 6, 1.0, precision=prec, flags=drt.FLAG_IMAGE), stats=True)
                print(f"spheres={n:3d} params={sc.n_params:3d} {name}: {best.kernel_ms:8.2f} ms with grad, "
                      f"{st_nograd.kernel_ms:8.2f} ms image only, {best.paths / best.kernel_ms / 1e3:8.1f} Mpaths/s, "
                      f"seg/path {best.segments / best.paths:.2f} lit {best.lit_paths / best.paths:.3f}")


if __name__ == "__main__":
    main()
