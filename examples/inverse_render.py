#!/usr/bin/env python
"""Config 3 of BASELINE.json: recover the Cornell box's wall albedos from a
target image by gradient descent on the rendered image's MSE.

    python examples/inverse_render.py                      # 1 GPU
    torchrun --nproc-per-node 8 examples/inverse_render.py # image rows sharded over 8 GPUs

Per iteration and per rank (no collective inside the renders):
  1. render this rank's row bands with the current albedos          (image only)
  2. seed = dLoss/dImage = 2 (I - I*) / (3 W H) on the device
  3. adjoint pass with that per-pixel seed, seed_scale = 1/spp, on a DIFFERENT
     sample stream (decorrelated image / gradient estimates, cf. the reference
     README's remark on biased gradients and integrate.hpp:39-52)
  4. the 12 gradient scalars are summed over the ranks by the one-block peer-store kernel that
     closes the adjoint render (drtb_set_grad_peers: no collective library on the loop's critical
     path; NCCL all-reduce is the fallback), then a clamped gradient-descent step ON THE DEVICE,
     pushed with drtb_set_params_device on the render stream: the loop never synchronises with
     the host.  The loss history is summed over the ranks once, after the loop.
"""
from __future__ import annotations

import argparse
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import drt_b200 as drt  # noqa: E402

TRUE = dict(red=(0.5, 0.0, 0.0), green=(0.0, 0.5, 0.0), white=(0.5, 0.5, 0.5))


def fit(width=256, height=256, spp=64, bounces=4, iters=100, lr=None, start=0.3, precision=drt.F64,
        band_rows=8, verbose=False, peer_exchange=True):
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    stream = torch.cuda.current_stream().cuda_stream

    ctx = drt.Context(local)
    scene = drt.cornell_box(width, height, **TRUE)
    ctx.upload(scene)
    rows = drt.shard_rows(height, rank, world, band_rows)
    P = len(scene.params)
    shard = dict(shard_index=rank, shard_count=world, band_rows=band_rows, precision=precision)
    img = torch.empty((rows, width, 3), dtype=torch.float64, device=dev)
    target = torch.empty_like(img)
    seed = torch.empty_like(img)
    grad = torch.empty((P, 3), dtype=torch.float64, device=dev)

    # gradient sum over the ranks inside the render (peer stores over NVLink); NCCL if the buffers cannot be mapped
    pgrad = None
    if world > 1 and peer_exchange:
        from differentiable_renderer_b200 import sharding
        try:
            pgrad = sharding.PeerGrad(ctx, dist)
            ok = torch.ones(1, device=dev)
        except Exception as e:                                                   # noqa: BLE001 -- reported, not hidden
            print(f"[inverse_render] rank {rank}: peer gradient exchange unavailable ({e}); using NCCL", file=sys.stderr)
            ok = torch.zeros(1, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0 and pgrad is not None:
            pgrad.close(); pgrad = None
    # the target: the true scene on its own stream
    ctx.render_device(drt.make_opts(spp * 4, bounces, 1.0, seed=1000, flags=drt.FLAG_IMAGE, **shard),
                      0, target.data_ptr(), 0, 0, stream)
    # the whole loop stays on the device: parameters, update step and loss history are tensors,
    # drtb_set_params_device copies on the render stream, nothing returns to the host until the end
    theta = torch.tensor([[start] * 3, [start] * 3, [start] * 3, [1.0, 1.0, 1.0]], dtype=torch.float64, device=dev)  # emission is known
    if lr is None:
        lr = 4.0
    # one row per iteration: [12 gradient scalars | squared error]; the adjoint render writes the gradients
    # straight into the row, the all-reduce sums the row, the loss history is the last column
    rows_it = torch.zeros((iters + 1, P * 3 + 1), dtype=torch.float64, device=dev)
    diff = torch.empty_like(img)
    theta3 = theta[:3]
    sq = torch.empty_like(img)
    scale = 2.0 / (3.0 * width * height)

    def iteration(it):
        row = rows_it[it]
        ctx.set_params_device(theta.data_ptr(), P, stream)
        ctx.render_device(drt.make_opts(spp, bounces, 1.0, seed=2 * it + 1, flags=drt.FLAG_IMAGE, **shard),
                          0, img.data_ptr(), 0, 0, stream)
        torch.sub(img, target, out=diff)
        torch.mul(diff, scale, out=seed)
        torch.mul(diff, diff, out=sq)
        torch.sum(sq.view(-1), dim=0, out=row[-1])                               # squared error of this rank's rows
        ctx.render_device(drt.make_opts(spp, bounces, 1.0, seed=2 * it + 2, flags=drt.FLAG_GRAD,
                                        seed_scale=1.0 / spp, **shard),
                          seed.data_ptr(), 0, row.data_ptr(), 0, stream)
        if world > 1 and pgrad is None:
            dist.all_reduce(row)                                                 # fallback: the one collective
        g = row[:9].view(3, 3)
        step = lr * 0.02 * (0.97 ** it)
        theta3.addcdiv_(g, g.abs().max().clamp_min_(1e-30).expand_as(g), value=-step).clamp_(0.0, 1.0)

    # two untimed iterations: NCCL builds its communicator and torch loads its kernels on first use
    theta0 = theta.clone()
    for it in range(2):
        iteration(it)
    theta.copy_(theta0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for it in range(iters):
        iteration(it)
        if verbose and rank == 0 and (it % 10 == 0 or it == iters - 1):
            th = theta.cpu().numpy()
            print(f"it {it:3d} loss {float(rows_it[it, -1]) / (3.0 * width * height):.3e} red {th[0].round(3)} green {th[1].round(3)} white {th[2].round(3)}")
    torch.cuda.synchronize()
    secs = time.perf_counter() - t0
    if pgrad is not None:
        losses = rows_it[:, -1].contiguous()                                     # the squared errors of the other ranks' rows
        dist.all_reduce(losses)
        rows_it[:, -1] = losses
        pgrad.close()
    ctx.close()
    true = np.array([TRUE["red"], TRUE["green"], TRUE["white"]])
    theta = theta.cpu().numpy()
    history = (rows_it[:iters, -1] / (3.0 * width * height)).cpu().tolist()
    return theta[:3], float(np.abs(theta[:3] - true).max()), history, iters / secs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--spp", type=int, default=64)
    ap.add_argument("--bounces", type=int, default=4)
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--nccl", action="store_true", help="sum the gradients with ncclAllReduce instead of the peer exchange (A/B)")
    ap.add_argument("--quiet", action="store_true")
    a = ap.parse_args()
    import torch.distributed as dist
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl")
    theta, err, hist, ips = fit(a.size, a.size, a.spp, a.bounces, a.iters, verbose=not a.quiet, peer_exchange=not a.nccl)
    if not dist.is_initialized() or dist.get_rank() == 0:
        print(f"final max |albedo - true| = {err:.4f}; loss {hist[0]:.3e} -> {hist[-1]:.3e}; {ips:.1f} iterations/s")
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
