"""Config 4 of BASELINE.json: ~1 M procedurally tessellated triangles, GPU-built
BVH, 1024x1024, 64 spp, 8 bounces, gradient w.r.t. per-triangle albedo.
Reports build time, throughput, traversal work per ray and the algorithmic-bytes
roofline of SURVEY.md §8(d): ceil(log2(N/4)) * 64 B + 4 * 48 B + 32 B per segment.

    python tools/mesh_bench.py [--grid 204 --sphere 362] [--size 1024] [--spp 64] [--out file.json]
"""
import argparse, json, math, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import drt_b200 as drt

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=204)
ap.add_argument("--sphere", type=int, default=362)
ap.add_argument("--size", type=int, default=1024)
ap.add_argument("--spp", type=int, default=64)
ap.add_argument("--bounces", type=int, default=8)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--out", default="")
a = ap.parse_args()
t0 = time.time()
scene = drt.tessellated_room(a.grid, a.sphere, width=a.size, height=a.size)
n = scene.mesh.n_triangles
gen_s = time.time() - t0
res = {"workload": f"tessellated_room: {n} triangles, {a.size}x{a.size}, {a.spp} spp, min_bounces={a.bounces}, absorb=1, "
                   f"per-triangle albedo gradient ({scene.mesh.n_triangles * 3} scalars)", "host_generate_s": gen_s}
peaks = {}
try:
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
except Exception:
    pass
hbm = float(peaks.get("hbm_gbs", 6650.0))
with drt.Context(0) as ctx:
    t0 = time.time(); ctx.upload(scene); res["upload_and_build_s"] = time.time() - t0
    res["bvh_build_ms_device"] = ctx.mesh_build_ms
    for prec, name in ((drt.F64, "f64"), (drt.F32, "f32")):
        best = None
        for _ in range(a.reps):
            img, grad, st = ctx.render(drt.make_opts(a.spp, a.bounces, 1.0, precision=prec), stats=True)
            if best is None or st.kernel_ms < best.kernel_ms:
                best = st
        bytes_per_seg = math.ceil(math.log2(n / 4)) * 64 + 4 * 48 + 32
        gbs = best.segments * bytes_per_seg / (best.kernel_ms * 1e-3) / 1e9
        res[name] = {"kernel_ms": best.kernel_ms, "Mpaths_per_s": best.paths / best.kernel_ms / 1e3,
                     "Msegments_per_s": best.segments / best.kernel_ms / 1e3,
                     "segments_per_path": best.segments / best.paths, "lit_fraction": best.lit_paths / best.paths,
                     "bvh_nodes_per_segment": best.bvh_nodes / best.segments, "tri_tests_per_segment": best.tri_tests / best.segments,
                     "algorithmic_bytes_per_segment": bytes_per_seg, "algorithmic_GBps": gbs, "hbm_peak_GBps": hbm,
                     "frac_of_hbm_peak": gbs / hbm,
                     "actual_bytes_per_segment_est": best.bvh_nodes / best.segments * 128 + best.tri_tests / best.segments * 48,
                     "image_mean": float(img.mean()), "nonzero_triangle_grads": int((abs(grad[scene.mesh.param_base:]).sum(1) > 0).sum())}
        print(name, json.dumps(res[name]))
print(json.dumps({k: v for k, v in res.items() if k not in ("f64", "f32")}))
if a.out:
    Path(a.out).write_text(json.dumps(res, indent=1) + "\n")
