"""The image all-gather fused into the render (drtb_set_image_peers): sharded renders
store their pixels straight into the full image of every peer.  On one GPU the "peers"
are two buffers of the same device, which exercises the kernel's store path; the
cross-process mapping (CUDA IPC) is exercised by bench.py at N > 1, which checks the
peer-filled image against an NCCL all-gather before it times anything."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("spp,mb,absorb", [(40, 3, 0.3), (6, 4, 1.0), (32, 8, 1.0)])
@pytest.mark.parametrize("count,band", [(3, 4), (2, 8)])
def test_sharded_renders_fill_every_peer_image(drt, ctx, spp, mb, absorb, count, band):
    import torch
    W, H = 40, 28                                   # ragged last band for both band sizes
    ctx.upload(drt.cornell_box(W, H))
    ref_img, ref_grad = ctx.render(drt.make_opts(spp, mb, absorb))
    dev = torch.device("cuda", 0)
    # poisoned: every pixel must be written by some shard
    bufs = [torch.full((H, W, 3), float("nan"), dtype=torch.float64, device=dev) for _ in range(2)]
    grads = torch.zeros((count,) + ref_grad.shape, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    try:
        ctx.set_image_peers([b.data_ptr() for b in bufs])
        for r in range(count):
            o = drt.make_opts(spp, mb, absorb, shard_index=r, shard_count=count, band_rows=band)
            ctx.render_device(o, 0, 0, grads[r].data_ptr(), 0, stream)     # no compact shard image at all
        torch.cuda.synchronize()
    finally:
        ctx.set_image_peers([])
    for b in bufs:
        assert np.array_equal(b.cpu().numpy(), ref_img)       # pixels do not depend on the sharding: bit-equal
    g = grads.sum(0).cpu().numpy()
    assert np.abs(g - ref_grad).max() <= 1e-12 * np.abs(ref_grad).max()


def test_ipc_alloc_round_trip_and_validation(drt, ctx):
    ctx.upload(drt.cornell_box(8, 8))
    ptr, handle = ctx.ipc_alloc(8 * 8 * 3 * 8)
    assert ptr != 0 and len(handle) == 64 and any(handle)
    import torch
    oi = drt.make_opts(4, 1, 0.5, flags=drt.FLAG_IMAGE)
    try:
        ctx.set_image_peers([ptr])
        ctx.render_device(oi, 0, 0, 0, 0, 0)        # image-only render into the exportable buffer
        torch.cuda.synchronize()
    finally:
        ctx.set_image_peers([])
        ctx.ipc_free(ptr)
    with pytest.raises(drt.DrtbError):
        ctx.set_image_peers([1] * 9)                # more than 8 GPUs in a box
    with pytest.raises(drt.DrtbError):
        ctx.set_image_peers([0])                    # NULL image
    with pytest.raises(drt.DrtbError):              # peers off: an image render still needs d_img
        ctx.render_device(oi, 0, 0, 0, 0, 0)


@pytest.mark.parametrize("spp,mb,absorb", [(8, 4, 1.0), (32, 1, 0.5)])
def test_peer_gradient_exchange_sums_the_ranks(drt, spp, mb, absorb):
    """drtb_set_grad_peers: two contexts play two ranks (here on one GPU, each on its own stream); every render ends
    with the one-block exchange kernel, and both come back with the SUM of the shards' gradients -- equal to the
    unsharded render's, bit-identical on both ranks, call after call (the slots alternate parity)."""
    import torch
    W, H = 48, 32
    dev = torch.device("cuda", 0)
    with drt.Context(0) as a, drt.Context(0) as b:
        ctxs = [a, b]
        for c in ctxs:
            c.upload(drt.cornell_box(W, H))
        ref_img, ref_grad = a.render(drt.make_opts(spp, mb, absorb))
        nbytes = a.grad_exchange_bytes(2)
        assert nbytes == (2 * 2 * 12 + 2) * 8
        bufs = [torch.zeros(nbytes // 8, dtype=torch.float64, device=dev) for _ in range(2)]
        streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
        grads = [torch.zeros((4, 3), dtype=torch.float64, device=dev) for _ in range(2)]
        imgs = [torch.zeros((H // 2, W, 3), dtype=torch.float64, device=dev) for _ in range(2)]
        torch.cuda.synchronize()
        try:
            for r, c in enumerate(ctxs):
                c.set_grad_peers([x.data_ptr() for x in bufs], r)
            for rep in range(5):                                    # several calls: epochs and parities advance together
                for r, c in enumerate(ctxs):
                    o = drt.make_opts(spp, mb, absorb, shard_index=r, shard_count=2, band_rows=8)
                    c.render_device(o, 0, imgs[r].data_ptr(), grads[r].data_ptr(), 0, streams[r].cuda_stream)
                torch.cuda.synchronize()
                g0, g1 = grads[0].cpu().numpy(), grads[1].cpu().numpy()
                assert np.isfinite(g0).all(), "the exchange timed out"
                assert np.array_equal(g0, g1)                       # same order of addition on every rank
                assert np.abs(g0 - ref_grad).max() <= 1e-12 * np.abs(ref_grad).max()
        finally:
            for c in ctxs:
                c.set_grad_peers([])
        # off again: a sharded render returns its own share only
        o = drt.make_opts(spp, mb, absorb, shard_index=0, shard_count=2, band_rows=8)
        a.render_device(o, 0, imgs[0].data_ptr(), grads[0].data_ptr(), 0, 0)
        torch.cuda.synchronize()
        assert np.abs(grads[0].cpu().numpy()).sum() < np.abs(ref_grad).sum()
        with pytest.raises(drt.DrtbError):
            a.set_grad_peers([1, 2], 2)                             # rank out of range
