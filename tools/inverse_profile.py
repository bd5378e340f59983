"""GPU time of the pieces of one inverse-rendering iteration (development aid): image-only render,
adjoint render with a seed image, both in one call, and the tensor arithmetic between them."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch, numpy as np
import drt_b200 as drt
dev = torch.device("cuda", 0); stream = torch.cuda.current_stream().cuda_stream
W = H = 256; spp = 64; B = 4
with drt.Context(0) as ctx:
    sc = drt.cornell_box(W, H); ctx.upload(sc)
    img = torch.empty((H, W, 3), dtype=torch.float64, device=dev); seed = torch.randn_like(img); grad = torch.empty((4, 3), dtype=torch.float64, device=dev)
    def timeit(f, n=50):
        for _ in range(5): f()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); 
        for _ in range(n): f()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    oi = drt.make_opts(spp, B, 1.0, flags=drt.FLAG_IMAGE)
    og = drt.make_opts(spp, B, 1.0, flags=drt.FLAG_GRAD, seed_scale=1.0 / spp)
    ob = drt.make_opts(spp, B, 1.0)
    print("image only  ms", timeit(lambda: ctx.render_device(oi, 0, img.data_ptr(), 0, 0, stream)))
    print("grad only   ms", timeit(lambda: ctx.render_device(og, seed.data_ptr(), 0, grad.data_ptr(), 0, stream)))
    print("image+grad  ms", timeit(lambda: ctx.render_device(ob, 0, img.data_ptr(), grad.data_ptr(), 0, stream)))
    def torch_part():
        diff = img - seed
        torch.mul(diff, 2.0, out=seed)
        x = (diff * diff).sum() / 3.0
    print("torch part  ms", timeit(torch_part))
