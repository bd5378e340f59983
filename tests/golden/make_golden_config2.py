"""Golden vector of BASELINE.json's config 2 AT FULL SIZE: Cornell box 1024x1024, 256 spp,
8 bounces (min_bounces = 8, absorb = 1), seed (1,1,1), produced once by the UNMODIFIED
reference headers (oracle/_ref/libdrt_ref.so, /root/reference/include/drt compiled by
oracle/Makefile) driven like src/render.cpp:72-86 with `.backward` enabled.

    python tests/golden/make_golden_config2.py [--threads N]      # ~10 minutes on 8 cores

268 435 456 paths.  The full image is 25 MB of doubles, too large for a fixture, so the file
keeps (all in double):
  grad      the 12 gradient scalars (4 parameters x RGB), unnormalised sums (render.cpp:78-82)
  sub       every 4th pixel of every 4th row, img[::4, ::4]  (256 x 256 x 3): per-pixel check
  rows      per-row sums      img.sum(axis=1)  (1024 x 3)
  cols      per-column sums   img.sum(axis=0)  (1024 x 3)
  tiles     sums over 16 x 16 pixel tiles      (64 x 64 x 3): every pixel is in exactly one
  total     img.sum((0, 1))
The CUDA path is compared per pixel on `sub` (1e-4 relative, the north-star tolerance) and on
every aggregate; a wrong pixel anywhere moves its tile / row / column sum.
"""
import sys
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import oracle_lib  # noqa: E402
from oracle_lib import drt  # noqa: E402

W = H = 1024
SPP, MB, AB = 256, 8, 1.0
NAME = "cbox_1024x1024_256spp_b8_p1_config2"


def reduce_image(img: np.ndarray) -> dict:
    return dict(sub=img[::4, ::4].copy(), rows=img.sum(axis=1), cols=img.sum(axis=0),
                tiles=img.reshape(H // 16, 16, W // 16, 16, 3).sum(axis=(1, 3)), total=img.sum(axis=(0, 1)))


def main():
    assert oracle_lib.have_ref(), "needs /root/reference (build container only)"
    threads = int(sys.argv[sys.argv.index("--threads") + 1]) if "--threads" in sys.argv else 8
    t = time.time()
    img, grad = oracle_lib.ref_render(drt.cornell_box(W, H), drt.make_opts(SPP, MB, AB), threads=threads)
    dt = time.time() - t
    assert np.isfinite(img).all() and np.isfinite(grad).all()
    np.savez_compressed(HERE / f"{NAME}.npz", grad=grad, **reduce_image(img),
                        meta=np.array([W, H, SPP, MB, AB, 0, 0], dtype=np.float64))
    print(f"{NAME}: {W * H * SPP / dt / 1e6:.3f} Mpaths/s on {threads} threads ({dt:.0f} s)")
    print("mean RGB", img.reshape(-1, 3).mean(0))
    print("grad", grad)


if __name__ == "__main__":
    main()
