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
