"""Small mesh + analytic renders for compute-sanitizer (development aid): the wavefront with its compacting adjoint,
both precisions, the deterministic sink, and a pinned-image analytic render."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import drt_b200 as drt

with drt.Context(0) as ctx:
    scene = drt.tessellated_room(6, 12, width=48, height=40)
    ctx.upload(scene)
    for prec in (drt.F64, drt.F32):
        for spp, mb, ab in ((8, 3, 1.0), (40, 2, 0.5), (3, 4, 1.0)):
            img, grad, st = ctx.render(drt.make_opts(spp, mb, ab, precision=prec), stats=True)
            print("mesh", prec, spp, mb, ab, float(img.mean()), st.segments, st.lit_paths)
    det = drt.FLAG_IMAGE | drt.FLAG_GRAD | drt.FLAG_DETERMINISTIC
    img, grad = ctx.render(drt.make_opts(8, 3, 1.0, flags=det))
    print("deterministic", float(np.abs(grad).sum()))
    import torch
    ctx.upload(drt.cornell_box(48, 40))
    h_img = torch.empty((40, 48, 3), dtype=torch.float64).pin_memory()
    h_grad = torch.empty((4, 3), dtype=torch.float64).pin_memory()
    for o in (drt.make_opts(32, 4, 1.0), drt.make_opts(8, 1, 0.5)):
        ctx.render_host_ptrs(o, 0, h_img.data_ptr(), h_grad.data_ptr())
        print("pinned", float(h_img.mean()), float(h_grad.sum()))
