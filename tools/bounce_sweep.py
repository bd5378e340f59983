"""Config 5 of BASELINE.json: max-bounce sweep B = 1..16, Cornell box 2048x2048, 128 spp, on 1 / 2 / 4 / 8 GPUs
(divergence and scaling stress).

    python tools/bounce_sweep.py [--out profiles/x.json]                                   # 1 GPU
    python -m torch.distributed.run --nproc-per-node N ... tools/bounce_sweep.py --gpus N  # N GPUs, one rank each
    ncu --metrics smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum --clock-control none \
        -k regex:render_kernel --csv --log-file gpurun_out/sweep_warp.csv python tools/bounce_sweep.py --once
    python tools/bounce_sweep.py --merge-ncu gpurun_out/sweep_warp.csv --out profiles/x.json  # adds lanes/instruction

Per B: a step = one render of the whole image (rows in interleaved bands over the ranks, the image assembled by the
kernels' peer stores, gradients summed by the peer exchange kernel -- the bench's N > 1 path), device-resident,
CUDA events on the launching stream, max over ranks, best of 3 after a warm-up.
"""
import argparse, csv, json, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=2048)
ap.add_argument("--spp", type=int, default=128)
ap.add_argument("--precision", default="f64")
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--out", default="")
ap.add_argument("--bounces", default="1,2,3,4,6,8,12,16")
ap.add_argument("--once", action="store_true", help="one render per B, no timing loop (the ncu target)")
ap.add_argument("--merge-ncu", default="", help="ncu --csv log of a --once run: adds lanes per instruction to --out")
a = ap.parse_args()
BS = [int(x) for x in a.bounces.split(",")]

if a.merge_ncu:
    rows = [r for r in csv.reader(open(a.merge_ncu)) if len(r) > 5]
    hdr = next(r for r in rows if "Metric Name" in r)
    iname, ival, ikern, igrid = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Kernel Name"), hdr.index("Grid Size")
    lanes, ms = [], []
    for r in rows:
        if r is hdr or len(r) <= ival or "render_kernel" not in r[ikern] or r[igrid].replace(" ", "") == "(1,1,1)":
            continue                                          # (1,1,1): the first-use launch of drtb_reserve, no work
        if r[iname].startswith("smsp__thread_inst_executed_per_inst_executed"):
            lanes.append(float(r[ival].replace(",", "")))
        if r[iname].startswith("gpu__time_duration"):
            ms.append(float(r[ival].replace(",", "")) / 1e6)
    out = json.loads(Path(a.out).read_text())
    assert len(lanes) == len(out["rows"]), (len(lanes), len(out["rows"]))
    for row, l, t in zip(out["rows"], lanes, ms):
        row["lanes_per_instruction"] = l                      # smsp__thread_inst_executed_per_inst_executed.ratio, of 32
        row["warp_execution_efficiency"] = l / 32.0
        row["kernel_ms_under_ncu"] = t
    out["warp_efficiency_source"] = "ncu --metrics smsp__thread_inst_executed_per_inst_executed.ratio, 1 GPU, one render per B"
    Path(a.out).write_text(json.dumps(out, indent=1) + "\n")
    for row in out["rows"]:
        print(row["bounces"], row["lanes_per_instruction"])
    sys.exit(0)

import torch
import drt_b200 as drt

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist = None
if world > 1:
    import torch.distributed as dist
    from differentiable_renderer_b200 import sharding
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
prec = drt.F64 if a.precision == "f64" else drt.F32
W = H = a.size
band = 8
rows_tbl = []
with drt.Context(local) as ctx:
    scene = drt.cornell_box(W, H)
    ctx.upload(scene)
    P = len(scene.params)
    stream = torch.cuda.current_stream()
    d_grad = torch.empty((P, 3), dtype=torch.float64, device=dev)
    d_stats = torch.zeros(8, dtype=torch.int64, device=dev)
    n_rows = drt.shard_rows(H, rank, world, band)
    peer = pgrad = None
    if world > 1:
        peer = sharding.PeerImage(ctx, H, W, dist)
        pgrad = sharding.PeerGrad(ctx, dist)
        d_img = None
    else:
        d_img = torch.empty((H, W, 3), dtype=torch.float64, device=dev)
    for B in BS:
        o = drt.make_opts(a.spp, B, 1.0, precision=prec, shard_index=rank, shard_count=world, band_rows=band)
        so = drt.make_opts(a.spp, B, 1.0, precision=prec, shard_index=rank, shard_count=world, band_rows=band,
                           flags=drt.FLAG_IMAGE | drt.FLAG_GRAD | drt.FLAG_STATS)
        img_ptr = 0 if world > 1 else d_img.data_ptr()
        ctx.reserve(so)
        ctx.render_device(so, 0, img_ptr, d_grad.data_ptr(), d_stats.data_ptr(), stream.cuda_stream)     # warm + counters
        torch.cuda.synchronize()
        if a.once:
            continue
        tot = torch.tensor([n_rows * W * a.spp, int(d_stats[1].item()), int(d_stats[2].item())], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tot)
        paths, segs, lit = (float(x) for x in tot.tolist())
        best = None
        for _ in range(3):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.render_device(o, 0, img_ptr, d_grad.data_ptr(), 0, stream.cuda_stream)
            e1.record(stream)
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = float(t.item()) if best is None else min(best, float(t.item()))
        row = {"bounces": B, "ms_per_step": best, "Mpaths_per_s": paths / best / 1e3, "Msegments_per_s": segs / best / 1e3,
               "segments_per_path": segs / paths, "lit_fraction": lit / paths}
        rows_tbl.append(row)
        if rank == 0:
            print(json.dumps(row), flush=True)
    if pgrad is not None:
        torch.cuda.synchronize(); pgrad.close()
    if peer is not None:
        torch.cuda.synchronize(); peer.close()
if rank == 0 and a.out and not a.once:
    out = {"workload": f"cornell_box {W}x{H}, {a.spp} spp, min_bounces=B, absorb=1", "precision": a.precision, "n_gpus": world,
           "timing": "CUDA events around the render (+ gradient reduction and peer exchange) on the launching stream, max over "
                     "ranks, best of 3", "rows": rows_tbl}
    Path(a.out).write_text(json.dumps(out, indent=1) + "\n")
if world > 1:
    dist.barrier(); dist.destroy_process_group()
