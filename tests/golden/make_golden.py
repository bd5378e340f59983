"""Generates the golden vectors under tests/golden/ from the UNMODIFIED
reference (oracle/_ref/libdrt_ref.so = /root/reference/include/drt compiled by
oracle/Makefile).  Run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

The reference ships no tests or fixtures of its own, so these
files -- produced by the reference's code itself -- are what pins the oracle
restatement and the CUDA path on machines without /root/reference.
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import oracle_lib  # noqa: E402
from oracle_lib import drt  # noqa: E402

CASES = {
    # name: (W, H, spp, min_bounces, absorb, seed, rand_mode)
    "cbox_48x32_8spp_b8_p1":     (48, 32, 8, 8, 1.0, 0, 0),
    "cbox_48x32_8spp_b1_p05":    (48, 32, 8, 1, 0.5, 0, 0),
    "cbox_48x32_8spp_b3_p03":    (48, 32, 8, 3, 0.3, 0, 0),
    "cbox_40x24_5spp_b0_p025_s7": (40, 24, 5, 0, 0.25, 7, 0),
    "cbox_32x32_40spp_b4_p1":    (32, 32, 40, 4, 1.0, 0, 0),
    # as-shipped behaviour: sequential unseeded glibc rand(), loop order of src/render.cpp:72-76
    "cbox_64x64_16spp_b1_p05_libc": (64, 64, 16, 1, 0.5, 0, 1),
}


def specular_and_grad_image():
    """The reference's own SpecularBxDF (bxdf.hpp:85-124) in the Cornell box, and the
    per-pixel gradient image of one parameter (README.md:138-145)."""
    for name, (W, H, spp, mb, ab, seed) in {"specbox_40x28_6spp_b4_p1": (40, 28, 6, 4, 1.0, 2),
                                            "specbox_40x28_6spp_b1_p05": (40, 28, 6, 1, 0.5, 2),
                                            "specbox_24x16_40spp_b3_p03": (24, 16, 40, 3, 0.3, 5)}.items():
        scene = drt.specular_box(W, H)
        img, grad, gimg = oracle_lib.ref_render(scene, drt.make_opts(spp, mb, ab, seed=seed), grad_image_of=4)
        assert np.isfinite(img).all() and np.isfinite(grad).all()
        np.savez_compressed(HERE / f"{name}.npz", img=img, grad=grad, gimg_gloss=gimg,
                            meta=np.array([W, H, spp, mb, ab, seed, 0], dtype=np.float64))
        print(name, img.reshape(-1, 3).mean(0), grad[4])
    # gradient image of the red wall's albedo in the plain Cornell box (the README figure)
    W, H, spp, mb, ab = 48, 32, 8, 8, 1.0
    img, grad, gimg = oracle_lib.ref_render(drt.cornell_box(W, H), drt.make_opts(spp, mb, ab, seed_scale=1.0 / spp),
                                            grad_image_of=0)
    np.savez_compressed(HERE / "cbox_48x32_8spp_b8_p1_gimg_red.npz", img=img, grad=grad, gimg=gimg,
                        meta=np.array([W, H, spp, mb, ab, 0, 0], dtype=np.float64))
    print("gimg red", gimg.sum((0, 1)), grad[0])


def main():
    assert oracle_lib.have_ref(), "needs /root/reference (build container only)"
    if "--only-specular" in sys.argv:
        return specular_and_grad_image()
    for name, (W, H, spp, mb, ab, seed, mode) in CASES.items():
        scene = drt.cornell_box(W, H)
        opts = drt.make_opts(spp, mb, ab, seed=seed)
        img, grad = oracle_lib.ref_render(scene, opts, rand_mode=mode)
        np.savez_compressed(HERE / f"{name}.npz", img=img, grad=grad,
                            meta=np.array([W, H, spp, mb, ab, seed, mode], dtype=np.float64))
        print(name, img.reshape(-1, 3).mean(0), grad[0])
    # per-pixel adjoint seed image + seed_scale
    W, H, spp = 32, 24, 6
    rng = np.random.default_rng(1234)
    seed_img = rng.uniform(-1, 1, size=(H, W, 3))
    scene = drt.cornell_box(W, H)
    opts = drt.make_opts(spp, 2, 0.4, seed_scale=1.0 / spp)
    img, grad = oracle_lib.ref_render(scene, opts, seed_img=seed_img)
    np.savez_compressed(HERE / "cbox_32x24_6spp_b2_p04_seedimg.npz", img=img, grad=grad, seed_img=seed_img,
                        meta=np.array([W, H, spp, 2, 0.4, 0, 0], dtype=np.float64))
    # explicit rays through Pathtracer::trace
    import ctypes as C
    lib = oracle_lib.load_ref()
    n = 64
    scene = drt.cornell_box(8, 8)
    sc = scene.flatten()
    opts = drt.make_opts(1, 3, 0.3)
    orig = rng.uniform(-1.5, 1.5, size=(n, 3)) + np.array([0, 0, 2.0])
    dirs = rng.normal(size=(n, 3)); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    dirs[::7] *= 1.7                                   # non-unit directions are legal input
    keys = rng.integers(0, 2**62, size=n, dtype=np.uint64)
    rad = np.zeros((n, 3)); jac = np.zeros((n, len(scene.params), 3))
    dp = C.POINTER(C.c_double)
    for i in range(n):
        rc = lib.drt_ref_trace_ray(C.byref(sc), C.byref(opts), orig[i].ctypes.data_as(dp), dirs[i].ctypes.data_as(dp),
                                   int(keys[i]), rad[i].ctypes.data_as(dp), jac[i].ctypes.data_as(dp))
        assert rc == 0
    np.savez_compressed(HERE / "rays_64_b3_p03.npz", orig=orig, dirs=dirs, keys=keys, radiance=rad, jac=jac)
    print("rays lit:", int((rad.sum(1) > 0).sum()), "of", n)
    # triangle mesh through the reference's raycast/scatter/tape (test-only Triangle<T> shape)
    scene = drt.tessellated_room(2, 4, width=32, height=24)
    img, grad = oracle_lib.ref_render(scene, drt.make_opts(4, 4, 1.0))
    np.savez_compressed(HERE / "mesh_room_98tri_32x24_4spp_b4.npz", img=img, grad=grad)
    print("mesh", img.reshape(-1, 3).mean(0), np.count_nonzero(grad))
    specular_and_grad_image()


if __name__ == "__main__":
    main()
