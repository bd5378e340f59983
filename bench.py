#!/usr/bin/env python
"""bench.py — Mpaths/s, forward + adjoint, Cornell box 1024x1024, 256 spp, 8 bounces.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path

A step is one full render of the workload: every camera sample of every pixel
traced to termination plus its gradient contribution (BASELINE.json `metric`).
`value`  : device-resident throughput (outputs stay in HBM), CUDA-event timed.
`e2e`    : the same through the C ABI call with HOST buffers (drtb_set_params +
           drtb_render), parameter H2D and image/gradient D2H inside the timing.
N > 1    : launched under torchrun, one rank per GPU; image rows are sharded in
           interleaved bands; the image gather is fused into the render kernel (every
           rank's kernel stores its pixels into all ranks' full images over NVLink,
           drtb_set_image_peers) and the gradient sum is a one-block peer-store kernel
           behind the render (drtb_set_grad_peers); both are checked once against NCCL
           (all-gather, all-reduce) before timing, and NCCL stays the fallback; max over
           ranks; the workload is fixed, so "scaling" is "strong".  e2e at N > 1 lands the
           ASSEMBLED image in one pinned host buffer on rank 0 every step.
Extra keys (N = 1): "mesh" = config 4 of BASELINE.json (1 M triangles, GPU-built BVH, wavefront
           integrator), "f32" / "mixed" = the other arithmetic modes on the headline workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "Mpaths/s fwd+adjoint (Cornell 1024^2, 8 bounces)"
UNIT = "Mpaths/s"
FLOP_PER_SEGMENT = 280.0      # algorithmic FLOP per ray segment
FLOP_PER_PATH = 30.0          # + camera ray per path


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--spp", type=int, default=256)
    ap.add_argument("--bounces", type=int, default=8)
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"])
    ap.add_argument("--band-rows", type=int, default=1)   # N > 1: 1-row bands balance the ranks best (5.472 vs 5.495 ms at 8 GPUs)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the mesh / f32 / mixed blocks (N = 1)")
    return ap.parse_args()


def workload_name(a):
    return (f"cornell_box {a.width}x{a.height}, {a.spp} spp, min_bounces={a.bounces}, absorb=1 "
            f"(= {a.bounces} bounces), grad wrt red/green/white albedo + emission, seed (1,1,1)")


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons, power = [], None, set(), []
        for ts, line in self.rows:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax = float(f[2]); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


# --------------------------------------------------------------------------- CPU legs
def kernel_source_sha() -> str:
    """Hash of the sources the analytic render kernels are compiled from: profiles/ncu_render_kernel.json carries
    the hash it was captured at, so that a stale ncu figure is not reported for a kernel that has changed."""
    import hashlib
    h = hashlib.sha256()
    csrc = ROOT / "differentiable-renderer_b200" / "csrc"
    for name in ("render_kernels.cuh", "path.cuh", "real.cuh", "rng.cuh", "sinks.cuh"):
        h.update((csrc / name).read_bytes())
    return h.hexdigest()[:16]


def cpu_one_core_as_shipped(a, seconds_budget: float = 4.0):
    """the reference AS SHIPPED -- one thread, the sequential unseeded glibc rand() of
    random.hpp:9, the loop order of src/render.cpp:72-76 -- on a small sample of the same scene and settings."""
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_lib
    import drt_b200 as drt
    if not oracle_lib.have_ref():
        return None
    side, spp = 48, min(a.spp, 16)
    t = time.perf_counter()
    oracle_lib.ref_render(drt.cornell_box(side, side), drt.make_opts(spp, a.bounces, 1.0), threads=1, rand_mode=1)
    dt = time.perf_counter() - t
    if dt < seconds_budget / 4:
        side = int(min(256, side * (seconds_budget / max(dt, 1e-3)) ** 0.5)) // 8 * 8
        t = time.perf_counter()
        oracle_lib.ref_render(drt.cornell_box(side, side), drt.make_opts(spp, a.bounces, 1.0), threads=1, rand_mode=1)
        dt = time.perf_counter() - t
    return {"value": side * side * spp / dt / 1e6, "unit": UNIT, "cores": 1,
            "sample": f"{side}x{side}, {spp} spp, min_bounces={a.bounces}, absorb=1, sequential glibc rand(), 1 thread"}


def cpu_render_rate(a, seconds_budget: float, threads: int | None = None):
    """Times the reference's CPU path (oracle/_ref when built, else the C port)
    on a bounded sample of the SAME workload (same scene, mb, absorb, stream);
    Mpaths/s is resolution independent, so the sample is a smaller image."""
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_lib
    import drt_b200 as drt
    kind = "reference" if oracle_lib.have_ref() else "port"
    # every host core this process may run on: torchrun exports OMP_NUM_THREADS=1, which would turn the
    # CPU arm into a single-thread run (the thread count is passed to the library explicitly)
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    if kind == "reference":
        lib = oracle_lib.load_ref()
        nthr = threads or max(cores, lib.drt_ref_max_threads())
        run = lambda sc, o: oracle_lib.ref_render(sc, o, threads=nthr)
    else:
        lib = oracle_lib.load_restate()
        nthr = threads or max(cores, lib.drt_oracle_max_threads())
        run = lambda sc, o: oracle_lib.restate_render(sc, o, threads=nthr)
    spp = min(a.spp, 16)
    # pilot to size the sample
    sc = drt.cornell_box(64, 64)
    o = drt.make_opts(spp, a.bounces, 1.0)
    side = 64
    for _ in range(4):                     # grow the sample until one step fills the budget
        t = time.perf_counter(); run(sc, o); pilot = time.perf_counter() - t
        if pilot >= seconds_budget / 2 or side >= 1024:
            break
        rate = side * side * spp / max(pilot, 1e-6)
        side = int(min(1024, max(64, (rate * seconds_budget / spp) ** 0.5))) // 8 * 8
        sc = drt.cornell_box(side, side)

    def step():
        t = time.perf_counter(); run(sc, o); dt = time.perf_counter() - t
        return side * side * spp, dt
    sample = f"{side}x{side}, {spp} spp, min_bounces={a.bounces}, absorb=1, same stream, {nthr} OpenMP threads"
    return kind, nthr, sample, step


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind, nthr, sample, step = cpu_render_rate(a, seconds_budget=6.0)
    for _ in range(a.warmup):
        step()
    paths, secs = 0, 0.0
    for _ in range(a.steps):
        p, dt = step(); paths += p; secs += dt
    v = paths / secs / 1e6
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": secs / a.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "sample_per_step": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": nthr, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), file=_REAL_STDOUT, flush=True)


# --------------------------------------------------------------------------- GPU arm
def timed_renders(ctx, torch, opts, d_img, d_grad, flush, reps=3):
    """Best CUDA-event time (ms) of `reps` device-resident renders, L2 flushed before each."""
    stream = torch.cuda.current_stream()
    best = None
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.render_device(opts, 0, d_img.data_ptr(), d_grad.data_ptr(), 0, stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return best


def extra_blocks(a, drt, ctx, scene, torch, dev, hbm_peak):
    """Separately keyed blocks beside the headline (N = 1): the other arithmetic modes on the headline workload and
    config 4 of BASELINE.json.  Same timing rules: device-resident, CUDA events, L2 flushed, warm."""
    import math
    out = {}
    W, H, spp, B = a.width, a.height, a.spp, a.bounces
    P = len(scene.params)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    d_img = torch.empty((H, W, 3), dtype=torch.float64, device=dev)
    d_grad = torch.empty((P, 3), dtype=torch.float64, device=dev)
    paths = W * H * spp
    for name, prec in (("f32", drt.F32), ("mixed", drt.MIXED)):
        try:
            o = drt.make_opts(spp, B, 1.0, precision=prec)
            ctx.reserve(o)
            ms = timed_renders(ctx, torch, o, d_img, d_grad, flush)
            out[name] = {"value": paths / (ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms,
                         "workload": "the headline workload in this arithmetic mode"}
            if prec == drt.MIXED:
                img, grad, st = ctx.render(o, stats=True)
                out[name]["retraced_path_fraction"] = st.retraced_paths / st.paths
        except drt.DrtbError as e:
            out[name] = {"unavailable": str(e)}
    # ---- config 4: ~1 M procedurally tessellated triangles, GPU-built BVH, 1024 x 1024, 64 spp, 8 bounces,
    # gradient w.r.t. per-triangle albedo (3 M scalars)
    try:
        t0 = time.time()
        mscene = drt.tessellated_room(204, 362, width=1024, height=1024)
        gen_s = time.time() - t0
        n = mscene.mesh.n_triangles
        with drt.Context(dev.index or 0) as mctx:
            t0 = time.time(); mctx.upload(mscene); up_s = time.time() - t0
            Pm = mscene.n_params
            m_img = torch.empty((1024, 1024, 3), dtype=torch.float64, device=dev)
            m_grad = torch.empty((Pm, 3), dtype=torch.float64, device=dev)
            bytes_per_seg = math.ceil(math.log2(n / 4)) * 64 + 4 * 48 + 32
            blk = {"workload": f"tessellated_room: {n} triangles, 1024x1024, 64 spp, min_bounces=8, absorb=1, per-triangle "
                               f"albedo gradient ({n * 3} scalars)", "bvh_build_ms_device": mctx.mesh_build_ms,
                   "host_generate_s": gen_s, "upload_and_build_s": up_s, "algorithmic_bytes_per_segment": bytes_per_seg}
            for name, prec in (("f64", drt.F64), ("f32", drt.F32)):
                o = drt.make_opts(64, 8, 1.0, precision=prec)
                _, _, st = mctx.render(o, stats=True)                            # warm + counters
                ms = timed_renders(mctx, torch, o, m_img, m_grad, flush, reps=2)
                gbs = st.segments * bytes_per_seg / (ms * 1e-3) / 1e9
                blk[name] = {"ms_per_step": ms, "Mpaths_per_s": st.paths / ms / 1e3, "Msegments_per_s": st.segments / ms / 1e3,
                             "segments_per_path": st.segments / st.paths, "bvh_nodes_per_segment": st.bvh_nodes / st.segments,
                             "tri_tests_per_segment": st.tri_tests / st.segments,
                             "node_and_triangle_bytes_per_segment": st.bvh_nodes / st.segments * 128 + st.tri_tests / st.segments * 64,
                             "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                                          "kernel": "wf_traverse (85 % of the step)"}}
            out["mesh"] = blk
    except Exception as e:                                                     # noqa: BLE001 -- reported in the line
        out["mesh"] = {"unavailable": repr(e)}
    return out


def run_b200_arm(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    import drt_b200 as drt
    from differentiable_renderer_b200 import sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.gpus != world:
        if world == 1 and a.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched under torchrun (one rank per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    W, H, spp, B = a.width, a.height, a.spp, a.bounces
    prec = drt.F64 if a.precision == "f64" else drt.F32
    scene = drt.cornell_box(W, H)
    ctx = drt.Context(local)
    ctx.upload(scene)
    P = len(scene.params)
    band = a.band_rows
    rows = drt.shard_rows(H, rank, world, band)
    equal = H % (band * world) == 0
    stream = torch.cuda.current_stream()

    opts = drt.make_opts(spp, B, 1.0, precision=prec, shard_index=rank, shard_count=world, band_rows=band)
    d_img = torch.empty((rows, W, 3), dtype=torch.float64, device=dev)
    d_grad = torch.empty((P, 3), dtype=torch.float64, device=dev)
    d_full = torch.empty((H, W, 3), dtype=torch.float64, device=dev) if world > 1 else d_img
    perm = sharding.deinterleave_index(H, world, band, dev) if (world > 1 and equal) else None
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2

    # ---- N > 1: the image gather is fused into the render kernel (peer stores over NVLink into every
    # rank's full image, drtb_set_image_peers); checked once against NCCL all-gather + re-interleave.
    # If the peers cannot be mapped the NCCL gather stays in the step, and the JSON line says which ran.
    peer, gather = None, "none (1 GPU)"
    if world > 1:
        gather = "NCCL all-gather of the row bands"
        try:
            peer = sharding.PeerImage(ctx, H, W, dist)
            ok = torch.ones(1, device=dev)
        except Exception as e:                                     # noqa: BLE001 -- reported, not hidden
            print(f"[bench] rank {rank}: peer image unavailable ({e}); using the NCCL gather", file=sys.stderr)
            ok = torch.zeros(1, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if (ok.item() == 0 or not equal) and peer is not None:     # ragged shards keep the NCCL-free compact path
            peer.close(); peer = None
        if peer is not None:
            full_p = peer.tensor(dev)
            full_p.fill_(float("nan"))
            dist.barrier(); torch.cuda.synchronize()
            ctx.render_device(opts, 0, 0, d_grad.data_ptr(), 0, stream.cuda_stream)       # peers on: no compact image
            dist.all_reduce(d_grad, op=dist.ReduceOp.SUM)
            torch.cuda.synchronize(); dist.barrier()
            ctx.set_image_peers([])
            ctx.render_device(opts, 0, d_img.data_ptr(), d_grad.data_ptr(), 0, stream.cuda_stream)
            dist.all_gather_into_tensor(d_full.view(world, rows, W, 3), d_img)
            ref = d_full.view(H, W, 3)[perm]
            same = torch.tensor([1.0 if torch.equal(full_p, ref) else 0.0], device=dev)
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            if same.item() != 1.0:
                raise SystemExit("peer-filled image differs from the NCCL all-gather")
            ctx.set_image_peers(peer.ptrs)
            gather = "fused: render kernel stores every pixel into all ranks' full images (NVLink peer stores), verified bit-equal to an NCCL all-gather"
    use_peer = peer is not None
    img_arg = 0 if use_peer else d_img.data_ptr()

    # ---- N > 1: the gradient sum as a peer-store kernel behind the render (drtb_set_grad_peers), checked once
    # against ncclAllReduce; NCCL stays in the step if the exchange buffers cannot be mapped.
    pgrad, grad_sum = None, "none (1 GPU)"
    if world > 1:
        grad_sum = f"NCCL all-reduce of {P * 3} doubles"
        try:
            pgrad = sharding.PeerGrad(ctx, dist)
            ok = torch.ones(1, device=dev)
        except Exception as e:                                     # noqa: BLE001 -- reported, not hidden
            print(f"[bench] rank {rank}: peer gradient exchange unavailable ({e}); using NCCL", file=sys.stderr)
            ok = torch.zeros(1, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0 and pgrad is not None:
            pgrad.close(); pgrad = None
        if pgrad is not None:
            dist.barrier(); torch.cuda.synchronize()
            ctx.render_device(opts, 0, img_arg, d_grad.data_ptr(), 0, stream.cuda_stream)        # exchange on: d_grad = the sum
            torch.cuda.synchronize()
            summed = d_grad.clone()
            ctx.set_grad_peers([])
            ctx.render_device(opts, 0, img_arg, d_grad.data_ptr(), 0, stream.cuda_stream)
            dist.all_reduce(d_grad, op=dist.ReduceOp.SUM)
            torch.cuda.synchronize()
            close = (summed - d_grad).abs().max() <= 1e-12 * d_grad.abs().max()              # NCCL adds in its own order
            ranks = [torch.empty_like(summed) for _ in range(world)]
            dist.all_gather(ranks, summed)
            same = torch.tensor([1.0 if (close and all(torch.equal(r, ranks[0]) for r in ranks)) else 0.0], device=dev)
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            if same.item() != 1.0:
                raise SystemExit("peer gradient exchange differs from ncclAllReduce")
            ctx.set_grad_peers(pgrad.ptrs, rank)
            grad_sum = (f"fused: one-block kernel behind the render stores {P * 3} doubles into every rank's exchange buffer "
                        "(NVLink peer stores) and adds them in rank order; verified against ncclAllReduce, bit-identical on all ranks")
    use_pgrad = pgrad is not None

    def step_device():
        ctx.render_device(opts, 0, img_arg, d_grad.data_ptr(), 0, stream.cuda_stream)        # + gradient exchange when on
        if world > 1:
            if not use_pgrad:
                dist.all_reduce(d_grad, op=dist.ReduceOp.SUM)      # the one collective of the path
            if equal and not use_peer:
                dist.all_gather_into_tensor(d_full.view(world, rows, W, 3), d_img)

    # ---- stats run (untimed): segments/path for the algorithmic FLOP count
    d_stats = torch.zeros(8, dtype=torch.int64, device=dev)      # sizeof(drtb_stats) = 64
    so = drt.make_opts(spp, B, 1.0, precision=prec, shard_index=rank, shard_count=world, band_rows=band,
                       flags=drt.FLAG_IMAGE | drt.FLAG_GRAD | drt.FLAG_STATS)
    ctx.render_device(so, 0, d_img.data_ptr(), d_grad.data_ptr(), d_stats.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    segs_local = int(d_stats[1].item()); lit_local = int(d_stats[2].item())
    paths_local = rows * W * spp
    tot = torch.tensor([paths_local, segs_local, lit_local], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot)
    paths_total, segs_total, lit_total = (float(x) for x in tot.tolist())

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing
    for _ in range(max(3, a.warmup)):
        step_device()
    sync_all()
    sampler = ClockSampler(local); sampler.start(); time.sleep(0.25)
    l0 = ctx.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
           torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    t0 = time.time()
    sync_all()
    for s0, s1, s2 in ev:
        flush.zero_()                                  # evict L2 between timed iterations (untimed)
        s0.record(stream)
        ctx.render_device(opts, 0, img_arg, d_grad.data_ptr(), 0, stream.cuda_stream)
        s1.record(stream)                              # render + gradient reduction (+ peer exchange) kernels
        if world > 1:
            if not use_pgrad:
                dist.all_reduce(d_grad, op=dist.ReduceOp.SUM)
            if equal and not use_peer:
                dist.all_gather_into_tensor(d_full.view(world, rows, W, 3), d_img)
        s2.record(stream)
    sync_all()
    t1 = time.time()
    launches = ctx.launch_count - l0
    clocks = sampler.stop(t0, t1)
    step_ms = sum(s0.elapsed_time(s2) for s0, s1, s2 in ev)
    kern_ms = sum(s0.elapsed_time(s1) for s0, s1, s2 in ev) / a.steps
    tt = torch.tensor([step_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms = float(tt.item())
    ms_per_step = total_ms / a.steps
    value = paths_total / (ms_per_step * 1e-3) / 1e6

    # ---- end to end: this step's inputs from pinned host memory, its results back in pinned host memory
    h_img = torch.empty((rows if world == 1 else H, W, 3), dtype=torch.float64).pin_memory()
    h_grad = torch.empty((P, 3), dtype=torch.float64).pin_memory()
    g_dev = torch.empty((P, 3), dtype=torch.float64, device=dev)
    pvals = scene.param_values()
    e2e_api = ("drtb_set_params + drtb_render (host buffers, pinned; the kernel stores the image straight into the pinned "
               "host buffer over PCIe, the gradients follow by one D2H copy)")
    peer2 = None
    if world > 1 and use_peer and use_pgrad:
        # N > 1: the ASSEMBLED image must land in ONE host buffer.  Every rank's kernel stores its pixels into rank
        # 0's full image, the gradient exchange orders "all stores have landed", and rank 0 copies the full image
        # out -- on a copy stream, so that the copy of step k runs under the render of step k + 1.  THREE full images
        # (and three pinned host images) rotate: image j is stored into again at step k + 3, which a fast rank can
        # begin as soon as the exchange of step k + 2 has completed everywhere; rank 0 makes its stream wait for the
        # copy of step k before it launches render k + 2 (whose exchange is what releases the others), a wait that
        # is always already satisfied.  (drtb.h: the caller double-buffers the full images.)
        peer2 = [sharding.PeerImage(ctx, H, W, dist), sharding.PeerImage(ctx, H, W, dist)]
        bufs = [(peer, peer.tensor(dev))] + [(q, q.tensor(dev)) for q in peer2]
        h_imgs = [h_img] + [torch.empty((H, W, 3), dtype=torch.float64).pin_memory() if rank == 0 else h_img for _ in range(2)]
        h_grads = [h_grad] + [torch.empty((P, 3), dtype=torch.float64).pin_memory() for _ in range(2)]
        copy_stream = torch.cuda.Stream(device=dev)
        copied = [None, None, None]
        e2e_api = ("drtb_set_params + drtb_render_device (image peers + gradient peers) + one D2H of the assembled "
                   "full image on rank 0 (pinned, on a copy stream under the next render; three rotating images), "
                   "gradients D2H on every rank")
        flip = [0]

        def step_e2e():
            k = flip[0]; flip[0] += 1
            j = k % 3
            pi, full = bufs[j]
            ctx.set_params(pvals)                                            # H2D: this step's inputs
            ctx.set_image_peers(pi.ptrs)
            if copied[(k - 2) % 3] is not None:
                stream.wait_event(copied[(k - 2) % 3])                       # see above: satisfied long ago
            ctx.render_device(opts, 0, 0, g_dev.data_ptr(), 0, stream.cuda_stream)
            h_grads[j].copy_(g_dev, non_blocking=True)                       # D2H: the summed gradients
            done = torch.cuda.Event(); done.record(stream)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done)
                if rank == 0:
                    h_imgs[j].copy_(full, non_blocking=True)                 # D2H: the whole image, assembled by the kernels
                ev_c = torch.cuda.Event(); ev_c.record(copy_stream)
            copied[j] = ev_c
    else:
        if use_peer:
            ctx.set_image_peers([])                                          # the host-buffer API returns this rank's rows
        if use_pgrad:
            ctx.set_grad_peers([])
        h_img = torch.empty((rows, W, 3), dtype=torch.float64).pin_memory()

        def step_e2e():
            ctx.set_params(pvals)                                            # H2D: this step's inputs
            ctx.render_host_ptrs(opts, 0, h_img.data_ptr(), h_grad.data_ptr())   # render + D2H image, gradients
            if world > 1:
                g_dev.copy_(h_grad, non_blocking=True)
                dist.all_reduce(g_dev, op=dist.ReduceOp.SUM)
                h_grad.copy_(g_dev)
    for _ in range(2):
        step_e2e()
    sync_all()
    te = time.perf_counter()
    for _ in range(a.steps):
        step_e2e()
    sync_all()
    e2e_s = time.perf_counter() - te
    te_t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te_t, op=dist.ReduceOp.MAX)
    e2e_value = paths_total / (float(te_t.item()) / a.steps) / 1e6
    h2d = P * 3 * 8 * world
    d2h = (H * W * 3 * 8) + P * 3 * 8 * world
    e2e_check = None
    if peer2 is not None and rank == 0:
        # the image in the host buffer is the whole picture: compare with the device image it was copied from
        last = (flip[0] - 1) % 3
        e2e_check = bool(torch.equal(h_imgs[last], bufs[last][1].cpu())) and bool(torch.isfinite(h_imgs[last]).all())

    if pgrad is not None:
        torch.cuda.synchronize()
        pgrad.close()                                                    # collective
    if peer2 is not None:
        torch.cuda.synchronize()
        for q in peer2:
            q.close()                                                    # collective
    if peer is not None:
        torch.cuda.synchronize()
        peer.close()                                                     # collective
    if rank != 0:
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (render_kernel), rank 0's launch
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    fma_peak = ctx.fma_peak(prec)                                       # TFLOP/s, measured live
    ctx_sm_count = torch.cuda.get_device_properties(local).multi_processor_count
    flop_launch = FLOP_PER_PATH * paths_local + FLOP_PER_SEGMENT * segs_local
    achieved_tf = flop_launch / (kern_ms * 1e-3) / 1e12
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    alg_bytes = rows * W * 3 * 8 + P * 3 * 8                            # the outputs; the scene is 0.4 KB
    # DRAM bytes per launch come from an ncu capture (counters are not readable from inside the run); the capture
    # records the hash of the kernel sources it was taken at and a figure from other sources is NOT reported
    traffic, traffic_note, issue = None, "no ncu capture for this workload under profiles/", None
    prof = ROOT / "profiles" / "ncu_render_kernel.json"
    if prof.exists():
        try:
            ent = json.loads(prof.read_text()).get(f"{a.precision}_{W}x{H}_{spp}spp_b{B}", {})
            if ent.get("source_sha") == kernel_source_sha():
                traffic, traffic_note = ent.get("dram_bytes_per_launch"), f"ncu --set full, {ent.get('source')}"
                if ent.get("warp_instructions") and world == 1:
                    # The bound this kernel actually runs against: one instruction per cycle and
                    # scheduler, an FP64 instruction holding the dispatch port for two.  Instruction counts from the
                    # same ncu capture as `traffic`, time and clock from this run.
                    slots = ent["warp_instructions"] + ent.get("fp64_warp_instructions", 0)
                    cycles = kern_ms * 1e-3 * 1.965e9 * ctx_sm_count * 4
                    issue = {"warp_instructions": ent["warp_instructions"], "fp64_warp_instructions": ent.get("fp64_warp_instructions"),
                             "issue_slots_needed": slots, "issue_cycles_available": cycles, "frac": slots / cycles,
                             "note": "instructions + FP64 instructions (two dispatch cycles each) over kernel time x 1.965 GHz x 4 schedulers x SMs"}
            elif ent:
                traffic_note = (f"stale: {ent.get('source')} was captured at kernel sources {ent.get('source_sha')}, "
                                f"this build is {kernel_source_sha()} (tools/ncu_traffic.sh regenerates it)")
        except Exception:
            pass
    nominal_tf = ctx_sm_count * (64 if a.precision == "f64" else 128) * 2 * 1.965e9 / 1e12
    roofline = {
        "bound": "fma_" + a.precision, "kernel": "render_kernel",
        "achieved": achieved_tf, "peak": fma_peak, "unit": "TFLOP/s", "frac": achieved_tf / fma_peak,
        "traffic": traffic, "traffic_source": traffic_note,
        "peak_source": "measured live: drtb_fma_peak (dependent-free FMA chains, all SMs); "
                       "MEASURED_PEAKS.json has no non-tensor FMA figure",
        "peak_nominal": nominal_tf, "frac_of_nominal": achieved_tf / nominal_tf,
        "peak_nominal_source": f"{ctx_sm_count} SMs x {64 if a.precision == 'f64' else 128} FMA/clk x 2 x 1.965 GHz",
        "algorithmic_flop_per_launch": flop_launch, "kernel_ms": kern_ms, "issue_bound": issue,
        "hbm": {"bound": "hbm", "achieved": alg_bytes / (kern_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": alg_bytes / (kern_ms * 1e-3) / 1e9 / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                "note": "outputs only (24 B/pixel); the path is FMA-issue bound, not HBM bound"},
    }

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(3, a.warmup),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": a.precision, "data": "synthetic",
        "config": {"workload": workload_name(a), "width": W, "height": H, "spp": spp, "bounces": B,
                   "paths_per_step": int(paths_total), "segments_per_path": segs_total / paths_total,
                   "lit_path_fraction": lit_total / paths_total,
                   "parallelism": f"pixel-band dp{world} (bands of {band} rows)",
                   "image_gather": gather, "gradient_sum": grad_sum,
                   "l2": "flushed between timed iterations (256 MiB memset, untimed)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": e2e_api, "assembled_image_on_rank0": e2e_check},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
    }

    if world == 1 and not a.no_extras:
        out.update(extra_blocks(a, drt, ctx, scene, torch, dev, hbm_peak))
    if world == 1 and not a.no_cpu_baseline:
        kind, nthr, sample, step = cpu_render_rate(a, seconds_budget=12.0)
        p, dt = step()
        out["cpu_baseline"] = {"value": p / dt / 1e6, "unit": UNIT, "cores": nthr, "kind": kind, "sample": sample,
                               "one_core_as_shipped": cpu_one_core_as_shipped(a)}
    print(json.dumps(out), file=_REAL_STDOUT, flush=True)
    ctx.close()
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


def main():
    a = parse_args()
    # stdout carries the one JSON line and nothing else: libraries that print there (NCCL's version
    # banner under NCCL_DEBUG=VERSION) are sent to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_b200_arm(a)


if __name__ == "__main__":
    main()
