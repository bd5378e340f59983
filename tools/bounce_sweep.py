"""Config 5 of BASELINE.json: max-bounce sweep B = 1..16, Cornell box 2048x2048,
128 spp (divergence stress).  One GPU; writes a JSON table.

    python tools/bounce_sweep.py [--size 2048] [--spp 128] [--precision f64] [--out profiles/x.json]
"""
import argparse, json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import drt_b200 as drt

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=2048)
ap.add_argument("--spp", type=int, default=128)
ap.add_argument("--precision", default="f64")
ap.add_argument("--out", default="")
ap.add_argument("--bounces", default="1,2,3,4,6,8,12,16")
a = ap.parse_args()
prec = drt.F64 if a.precision == "f64" else drt.F32
rows = []
with drt.Context(0) as ctx:
    ctx.upload(drt.cornell_box(a.size, a.size))
    for B in [int(x) for x in a.bounces.split(",")]:
        best = None
        for _ in range(3):
            _, _, st = ctx.render(drt.make_opts(a.spp, B, 1.0, precision=prec), stats=True)
            if best is None or st.kernel_ms < best.kernel_ms:
                best = st
        row = {"bounces": B, "kernel_ms": best.kernel_ms, "Mpaths_per_s": best.paths / best.kernel_ms / 1e3,
               "Msegments_per_s": best.segments / best.kernel_ms / 1e3, "segments_per_path": best.segments / best.paths,
               "lit_fraction": best.lit_paths / best.paths}
        rows.append(row)
        print(json.dumps(row))
out = {"workload": f"cornell_box {a.size}x{a.size}, {a.spp} spp, min_bounces=B, absorb=1", "precision": a.precision,
       "timing": "drtb_stats.kernel_ms (CUDA events around render + reduce kernels), best of 3", "rows": rows}
if a.out:
    Path(a.out).write_text(json.dumps(out, indent=1) + "\n")
