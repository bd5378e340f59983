"""The N > 1 path on CPU: two gloo ranks shard the pixel loop in interleaved row
bands, all-reduce the gradients, gather the image.  The shard renderer here is
the CPU oracle (test infrastructure); on GPUs it is the CUDA context and NCCL."""
import os
import socket

import numpy as np
import pytest

import oracle_lib
from oracle_lib import drt, rel_err, restate_render


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, H, W, band, out):
    import torch.distributed as dist
    from differentiable_renderer_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene = drt.cornell_box(W, H)

    def render_shard(index, count, band_rows):
        return restate_render(scene, drt.make_opts(3, 2, 0.4, shard_index=index, shard_count=count, band_rows=band_rows))
    img, grad = sharding.render_distributed(render_shard, H, band, dist=dist)
    np.savez(out + f".{rank}.npz", img=img, grad=grad)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("H,band", [(32, 4), (22, 4)])          # even bands, ragged last band
def test_two_rank_gloo_render_equals_single_process(tmp_path, H, band):
    import torch.multiprocessing as mp
    W, world = 24, 2
    out = str(tmp_path / "r")
    mp.spawn(_worker, args=(world, _free_port(), H, W, band, out), nprocs=world, join=True)
    full, grad = restate_render(drt.cornell_box(W, H), drt.make_opts(3, 2, 0.4))
    for r in range(world):
        z = np.load(out + f".{r}.npz")
        assert np.array_equal(z["img"], full)                     # pixels are independent: bit-equal
        assert rel_err(z["grad"], grad).max() < 1e-13             # summation order differs


def test_deinterleave_index_inverts_the_band_layout():
    from differentiable_renderer_b200 import sharding
    H, count, band = 64, 4, 8
    full = np.arange(H)
    cat = np.concatenate([full[sharding.shard_row_indices(H, r, count, band)] for r in range(count)])
    perm = sharding.deinterleave_index(H, count, band).numpy()
    assert np.array_equal(cat[perm], full)


class _FakeCtx:
    """Stands in for render.Context in the handle exchange: `ipc_alloc` hands out a made-up
    address and a handle that encodes it, `ipc_open` decodes a peer's handle."""
    def __init__(self, rank):
        self.rank, self.peers, self.opened, self.closed, self.freed = rank, None, [], [], []

    def ipc_alloc(self, nbytes):
        ptr = 0x7000_0000_0000 + self.rank * 0x1000_0000
        return ptr, ptr.to_bytes(8, "little") + bytes([self.rank]) * 56

    def ipc_open(self, handle):
        assert len(handle) == 64
        p = int.from_bytes(handle[:8], "little") + 1       # a peer mapping lives at another address
        self.opened.append(p)
        return p

    def set_image_peers(self, ptrs):
        self.peers = list(ptrs)

    def set_grad_peers(self, ptrs, rank=0):
        self.grad_peers, self.grad_rank = list(ptrs), rank

    def grad_exchange_bytes(self, n_ranks):
        return (2 * n_ranks * 12 + n_ranks) * 8

    def ipc_close(self, ptr):
        self.closed.append(ptr)

    def ipc_free(self, ptr):
        self.freed.append(ptr)


def _peer_worker(rank, world, port, out):
    import torch.distributed as dist
    from differentiable_renderer_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = _FakeCtx(rank)
    pi = sharding.PeerImage(ctx, 16, 8, dist)
    peers = list(ctx.peers)
    iface = pi.__cuda_array_interface__
    pi.close()
    np.savez(out + f".{rank}.npz", peers=np.array(peers, dtype=np.uint64), local=pi.local_ptr, nbytes=pi.nbytes,
             closed=np.array(ctx.closed, dtype=np.uint64), freed=np.array(ctx.freed, dtype=np.uint64),
             after=len(ctx.peers), shape=np.array(iface["shape"]))
    dist.destroy_process_group()


def test_peer_image_handle_exchange_on_two_gloo_ranks(tmp_path):
    """Host logic of the fused image gather: rank-ordered pointer list, own buffer by its
    local address, peers by their mapped address, everything unmapped on close."""
    import torch.multiprocessing as mp
    world = 2
    out = str(tmp_path / "p")
    mp.spawn(_peer_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    for r in range(world):
        z = np.load(out + f".{r}.npz")
        base = [0x7000_0000_0000 + q * 0x1000_0000 for q in range(world)]
        want = [base[q] if q == r else base[q] + 1 for q in range(world)]
        assert z["peers"].tolist() == want and int(z["local"]) == base[r]
        assert int(z["nbytes"]) == 16 * 8 * 3 * 8 and z["shape"].tolist() == [16, 8, 3]
        assert z["closed"].tolist() == [w for q, w in enumerate(want) if q != r]
        assert z["freed"].tolist() == [base[r]] and int(z["after"]) == 0


def _peer_grad_worker(rank, world, port, out):
    import torch.distributed as dist
    from differentiable_renderer_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = _FakeCtx(rank)
    pg = sharding.PeerGrad(ctx, dist)
    peers, grank = list(ctx.grad_peers), ctx.grad_rank
    pg.close()
    np.savez(out + f".{rank}.npz", peers=np.array(peers, dtype=np.uint64), rank=grank, nbytes=pg.nbytes,
             closed=np.array(ctx.closed, dtype=np.uint64), freed=np.array(ctx.freed, dtype=np.uint64), after=len(ctx.grad_peers))
    dist.destroy_process_group()


def test_peer_gradient_exchange_plumbing_on_two_gloo_ranks(tmp_path):
    """Host logic of the NCCL-free gradient sum (sharding.PeerGrad): buffer size from the ABI, rank-ordered
    pointer list with the own buffer by its local address, rank passed on, everything unmapped on close."""
    import torch.multiprocessing as mp
    world = 2
    out = str(tmp_path / "g")
    mp.spawn(_peer_grad_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    for r in range(world):
        z = np.load(out + f".{r}.npz")
        base = [0x7000_0000_0000 + q * 0x1000_0000 for q in range(world)]
        want = [base[q] if q == r else base[q] + 1 for q in range(world)]
        assert z["peers"].tolist() == want and int(z["rank"]) == r and int(z["nbytes"]) == (2 * 2 * 12 + 2) * 8
        assert z["closed"].tolist() == [w for q, w in enumerate(want) if q != r]
        assert z["freed"].tolist() == [base[r]] and int(z["after"]) == 0
