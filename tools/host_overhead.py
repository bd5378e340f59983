"""Host-side cost of one drtb_render_device call (development aid): a tiny image, so the kernels
take microseconds and the wall time per call is launch path + driver."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import drt_b200 as drt

dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream
with drt.Context(0) as ctx:
    ctx.upload(drt.cornell_box(16, 16))
    img = torch.empty((16, 16, 3), dtype=torch.float64, device=dev)
    grad = torch.empty((4, 3), dtype=torch.float64, device=dev)
    for flags, name in ((drt.FLAG_IMAGE, "image only"), (drt.FLAG_IMAGE | drt.FLAG_GRAD, "image + gradients")):
        o = drt.make_opts(32, 4, 1.0, flags=flags)
        for _ in range(20):
            ctx.render_device(o, 0, img.data_ptr(), grad.data_ptr(), 0, stream)
        torch.cuda.synchronize()
        n = 500
        t = time.perf_counter()
        for _ in range(n):
            ctx.render_device(o, 0, img.data_ptr(), grad.data_ptr(), 0, stream)
        t_issue = time.perf_counter() - t
        torch.cuda.synchronize()
        t_all = time.perf_counter() - t
        print(f"{name}: {t_issue / n * 1e6:.1f} us per call to enqueue, {t_all / n * 1e6:.1f} us per call including the GPU")
    t = time.perf_counter()
    for _ in range(n):
        o = drt.make_opts(32, 4, 1.0, flags=flags, seed=3)
    print(f"make_opts: {(time.perf_counter() - t) / n * 1e6:.1f} us")
