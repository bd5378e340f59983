"""Data-parallel sharding of the pixel loop (src/render.cpp:72-76) across ranks.

Every (pixel, sample) is independent once the stream is counter-based, so the
image is cut into bands of `band_rows` rows and band b goes to rank b % world
(interleaved so the cheap, directly-lit top rows are spread over all ranks).
There is no data-path collective inside the render; after it
  * the parameter gradients are summed with ONE all-reduce (NCCL on GPUs), and
  * the disjoint image bands are gathered and re-interleaved.
The transport is torch.distributed (plumbing); `render_shard` is whatever
renders one shard -- the CUDA context in production, the CPU oracle in the
world_size-2 gloo tests.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import numpy as np


def shard_row_indices(height: int, index: int, count: int, band_rows: int) -> np.ndarray:
    """Image rows (increasing y) owned by shard `index` of `count`."""
    ys = np.arange(height)
    if count <= 1:
        return ys
    return ys[(ys // max(1, band_rows)) % count == index]


def assemble_image(shards: List[np.ndarray], height: int, band_rows: int) -> np.ndarray:
    """Re-interleave the compact per-shard images into the full image."""
    count = len(shards)
    width = shards[0].shape[1]
    full = np.empty((height, width, 3), dtype=shards[0].dtype)
    for r, s in enumerate(shards):
        ys = shard_row_indices(height, r, count, band_rows)
        assert s.shape[0] >= len(ys)
        full[ys] = s[:len(ys)]
    return full


def max_shard_rows(height: int, count: int, band_rows: int) -> int:
    return max(len(shard_row_indices(height, r, count, band_rows)) for r in range(max(1, count)))


def render_distributed(render_shard: Callable[[int, int, int], Tuple[np.ndarray, np.ndarray]],
                       height: int, band_rows: int = 8, *, dist=None, device=None,
                       gather_image: bool = True) -> Tuple[Optional[np.ndarray], np.ndarray]:
    """Run `render_shard(index, count, band_rows) -> (img_shard, grad)` on this
    rank, all-reduce the gradients, gather the image.  Returns (full image,
    summed gradients) on every rank.  `dist` is torch.distributed (already
    initialised) or None for a single process."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        img, grad = render_shard(0, 1, band_rows)
        return img, grad
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    img, grad = render_shard(rank, world, band_rows)
    dev = device if device is not None else torch.device("cpu")
    g = torch.from_numpy(np.ascontiguousarray(grad)).to(dev)
    dist.all_reduce(g, op=dist.ReduceOp.SUM)                      # the one collective
    full = None
    if gather_image:
        rows = max_shard_rows(height, world, band_rows)           # ragged last band: pad
        pad = np.zeros((rows,) + img.shape[1:], dtype=img.dtype)
        pad[:img.shape[0]] = img
        mine = torch.from_numpy(pad).to(dev)
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        full = assemble_image([p.cpu().numpy() for p in parts], height, band_rows)
    return full, g.cpu().numpy()


def deinterleave_index(height: int, count: int, band_rows: int, device=None):
    """Row permutation that turns the rank-major concatenation of the compact
    shards into the full image: full = cat(shards)[perm].  Needs equal shards
    (height % (band_rows * count) == 0); used on-device by bench.py."""
    import torch
    assert height % (band_rows * count) == 0
    order = np.concatenate([shard_row_indices(height, r, count, band_rows) for r in range(count)])
    perm = np.empty(height, dtype=np.int64)
    perm[order] = np.arange(height)
    return torch.from_numpy(perm).to(device) if device is not None else torch.from_numpy(perm)


class PeerImage:
    """The full image (H x W x 3 doubles) replicated on every rank and filled by the
    render kernels themselves: each rank's kernel stores its pixels into all ranks'
    buffers over NVLink (drtb_set_image_peers), so no image gather follows the render.

    Host-side plumbing only: every rank allocates its buffer with `ctx.ipc_alloc`,
    the 64-byte CUDA IPC handles travel through `dist.all_gather_object`, peers are
    mapped with `ctx.ipc_open`, and the rank-ordered pointer list goes to
    `ctx.set_image_peers`.  `ctx` is a `render.Context` (or, in the CPU tests, any
    object with the same four methods)."""

    def __init__(self, ctx, height: int, width: int, dist):
        self.ctx, self.height, self.width = ctx, height, width
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.nbytes = height * width * 3 * 8
        self.local_ptr, handle = ctx.ipc_alloc(self.nbytes)
        handles = [None] * self.world
        dist.all_gather_object(handles, handle)
        self.ptrs = [self.local_ptr if r == self.rank else ctx.ipc_open(handles[r]) for r in range(self.world)]
        ctx.set_image_peers(self.ptrs)
        self._dist = dist

    # numba-style interface so torch (or cupy) can wrap the local buffer without a copy
    @property
    def __cuda_array_interface__(self):
        return {"shape": (self.height, self.width, 3), "typestr": "<f8", "data": (self.local_ptr, False),
                "version": 3, "strides": None}

    def tensor(self, device):
        import torch
        return torch.as_tensor(self, device=device)

    def close(self):
        """Collective: every rank unmaps its peers before anyone frees."""
        if self.ctx is None:
            return
        self.ctx.set_image_peers([])
        for r, p in enumerate(self.ptrs):
            if r != self.rank:
                self.ctx.ipc_close(p)
        self._dist.barrier()
        self.ctx.ipc_free(self.local_ptr)
        self.ctx = None


class PeerGrad:
    """The gradient sum over the ranks without a collective library (drtb_set_grad_peers): every rank allocates a
    small exchange buffer (zero-filled by `ctx.ipc_alloc`), the IPC handles travel through
    `dist.all_gather_object`, and from then on every render with FLAG_GRAD on `ctx` ends with a one-block kernel
    that stores this rank's gradients into all ranks' buffers over NVLink, waits for theirs and adds them in rank
    order -- `grad` comes back summed over the job, bit-identical on every rank.  Host-side plumbing only."""

    def __init__(self, ctx, dist):
        self.ctx = ctx
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.nbytes = ctx.grad_exchange_bytes(self.world)
        self.local_ptr, handle = ctx.ipc_alloc(self.nbytes)
        handles = [None] * self.world
        dist.all_gather_object(handles, handle)
        self.ptrs = [self.local_ptr if r == self.rank else ctx.ipc_open(handles[r]) for r in range(self.world)]
        dist.barrier()                                   # every buffer is zeroed and mapped before the first exchange
        ctx.set_grad_peers(self.ptrs, self.rank)
        self._dist = dist

    def close(self):
        """Collective: every rank unmaps its peers before anyone frees."""
        if self.ctx is None:
            return
        self.ctx.set_grad_peers([])
        for r, p in enumerate(self.ptrs):
            if r != self.rank:
                self.ctx.ipc_close(p)
        self._dist.barrier()
        self.ctx.ipc_free(self.local_ptr)
        self.ctx = None
