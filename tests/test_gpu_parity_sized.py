"""CUDA-vs-reference parity AT BASELINE.json's SIZES (VERDICT r1 "weak 1"): the small cases of
test_gpu_parity.py never reach the chunk planner's large plans, the two-pass gradient reduction
(> 4096 partial rows) or stream keys near (y W + x) spp ~ 2.7e8.

  config 1 exactly : 256 x 256, 16 spp, `-b 1 -p 0.5` (the app's defaults) and 8 bounces, against the
                     UNMODIFIED reference headers (oracle/_ref/libdrt_ref.so) and the C restatement
  reduced config 2 : 1024 x 1024 at 4 spp and 256 x 256 at 256 spp, same key formula
  full config 2    : 1024 x 1024, 256 spp, 8 bounces against the committed golden vector that
                     libdrt_ref.so produced once (tests/golden/make_golden_config2.py)
  config 4         : an 8 114-triangle tessellation against the oracle's Triangle shape, 2^16 rays
                     BVH-vs-linear-scan on the 1 M-triangle mesh, 3-sigma statistics at 1 M

Tolerances are the north-star's (image 1e-4 relative per pixel, gradients 1e-3 per parameter); the
double instantiation is additionally held to 1e-9.
"""
import os

import numpy as np
import pytest

import oracle_lib
from oracle_lib import rel_err, restate_render

pytestmark = pytest.mark.gpu

IMG_TOL, GRAD_TOL = 1e-4, 1e-3
GOLDEN = oracle_lib.ROOT / "tests" / "golden"
THREADS = max(1, len(os.sched_getaffinity(0)))


def check(img, grad, ref_img, ref_grad, tight=1e-9):
    e_img, e_grad = rel_err(img, ref_img).max(), rel_err(grad, ref_grad).max()
    assert e_img <= IMG_TOL and e_grad <= GRAD_TOL, (e_img, e_grad)
    assert e_img <= tight and e_grad <= tight, (e_img, e_grad)
    assert np.array_equal(img == 0.0, ref_img == 0.0)          # exact zeros stay exact zeros


@pytest.mark.parametrize("mb,absorb", [(1, 0.5), (8, 1.0)])
def test_config1_exact_size_against_the_reference(drt, ctx, mb, absorb):
    """BASELINE.json configs[0]: 256 x 256, 16 spp; src/render.cpp:72-86 with .backward enabled."""
    scene = drt.cornell_box(256, 256)
    ctx.upload(scene)
    img, grad, st = ctx.render(drt.make_opts(16, mb, absorb), stats=True)
    r_img, r_grad, r_st = restate_render(scene, drt.make_opts(16, mb, absorb), threads=THREADS, want_stats=True)
    assert st.paths == r_st.paths == 256 * 256 * 16
    assert st.segments == r_st.segments and st.lit_paths == r_st.lit_paths and st.truncated_paths == 0
    check(img, grad, r_img, r_grad)
    if oracle_lib.have_ref():                                  # the reference's own headers, unmodified
        f_img, f_grad = oracle_lib.ref_render(scene, drt.make_opts(16, mb, absorb), threads=THREADS)
        check(img, grad, f_img, f_grad)
    # the float instantiation on the same stream: a few paths per million take another hit sequence
    img32, grad32 = ctx.render(drt.make_opts(16, mb, absorb, precision=drt.F32))
    assert (rel_err(img32, r_img) > IMG_TOL).any(axis=-1).mean() <= 1e-3
    assert rel_err(grad32, r_grad).max() <= GRAD_TOL


@pytest.mark.parametrize("w,h,spp", [(1024, 1024, 4), (256, 256, 256)])
def test_reduced_config2_against_the_oracle(drt, ctx, w, h, spp):
    """config 2's integrator (8 bounces) at sizes the CPU finishes in seconds; 1024^2 x 4 spp
    has config 2's pixel count (262 144 warp tasks' worth of keys), 256^2 x 256 spp its sample count."""
    scene = drt.cornell_box(w, h)
    ctx.upload(scene)
    img, grad, st = ctx.render(drt.make_opts(spp, 8, 1.0), stats=True)
    r_img, r_grad, r_st = restate_render(scene, drt.make_opts(spp, 8, 1.0), threads=THREADS, want_stats=True)
    assert st.segments == r_st.segments and st.lit_paths == r_st.lit_paths
    check(img, grad, r_img, r_grad)


def test_reduced_config2_russian_roulette_at_size(drt, ctx):
    """The regenerating kernel at 1024^2 (16 384 pixel chunks, two-pass gradient reduction)."""
    scene = drt.cornell_box(1024, 1024)
    ctx.upload(scene)
    img, grad, st = ctx.render(drt.make_opts(4, 1, 0.5), stats=True)
    r_img, r_grad, r_st = restate_render(scene, drt.make_opts(4, 1, 0.5), threads=THREADS, want_stats=True)
    assert st.segments == r_st.segments and st.lit_paths == r_st.lit_paths
    check(img, grad, r_img, r_grad)


def test_full_config2_against_the_reference_golden(drt, ctx):
    """BASELINE.json configs[1] at FULL size, 268 435 456 paths, against numbers the unmodified reference
    headers produced (about 10 core-minutes x 8, generated once in the build container)."""
    z = np.load(GOLDEN / "cbox_1024x1024_256spp_b8_p1_config2.npz")
    W, H, spp, mb, ab = (int(z["meta"][0]), int(z["meta"][1]), int(z["meta"][2]), int(z["meta"][3]), float(z["meta"][4]))
    assert (W, H, spp, mb, ab) == (1024, 1024, 256, 8, 1.0)
    ctx.upload(drt.cornell_box(W, H))
    img, grad = ctx.render(drt.make_opts(spp, mb, ab))
    e_grad = rel_err(grad, z["grad"]).max()
    assert e_grad <= GRAD_TOL and e_grad <= 1e-9, e_grad
    sub = img[::4, ::4]
    e_sub = rel_err(sub, z["sub"]).max()                       # 65 536 pixels, one by one
    assert e_sub <= IMG_TOL and e_sub <= 1e-9, e_sub
    assert np.array_equal(sub == 0.0, z["sub"] == 0.0)
    # every pixel is in one tile, one row and one column sum
    tiles = img.reshape(H // 16, 16, W // 16, 16, 3).sum(axis=(1, 3))
    assert rel_err(tiles, z["tiles"]).max() <= 1e-9
    assert rel_err(img.sum(axis=1), z["rows"]).max() <= 1e-9
    assert rel_err(img.sum(axis=0), z["cols"]).max() <= 1e-9
    assert rel_err(img.sum(axis=(0, 1)), z["total"]).max() <= 1e-11


# ---- config 4 --------------------------------------------------------------------------------------
def test_8k_triangle_mesh_matches_oracle(drt, ctx):
    """a <= 8 192-triangle tessellation against the oracle's Triangle shape
    (Moller-Trumbore in double, O(N) per ray on the CPU)."""
    scene = drt.tessellated_room(16, 36, width=48, height=32)
    n = scene.mesh.n_triangles
    assert 8000 <= n <= 8192
    ctx.upload(scene)
    for spp, mb, ab in ((4, 4, 1.0), (6, 1, 0.4)):
        img, grad, st = ctx.render(drt.make_opts(spp, mb, ab, seed=3), stats=True)
        r_img, r_grad, r_st = restate_render(scene, drt.make_opts(spp, mb, ab, seed=3), threads=THREADS, want_stats=True)
        assert st.segments == r_st.segments and st.lit_paths == r_st.lit_paths
        assert rel_err(img, r_img).max() <= 1e-9
        assert np.abs(grad - r_grad).max() <= 1e-9 * np.abs(r_grad).max()
        assert (np.abs(grad[scene.mesh.param_base:]).sum(1) > 0).mean() > 0.02    # per-triangle albedo gradients
        assert st.tri_tests < 0.05 * st.segments * n


@pytest.fixture(scope="module")
def million(drt, ctx):
    scene = drt.tessellated_room(204, 362, width=128, height=128)
    assert 1_000_000 <= scene.mesh.n_triangles <= 1_100_000
    return scene


def test_bvh_equals_linear_scan_on_65536_rays_at_1m_triangles(drt, ctx, million):
    """2^16 explicit rays, two bounces each: the BVH traversal and the linear scan over all 1 M triangles
    return the same radiance bit for bit (closest t > 0, lower index on ties)."""
    ctx.upload(million)
    rng = np.random.default_rng(12)
    m = 1 << 16
    orig = rng.uniform(-1.5, 1.5, size=(m, 3)) + np.array([0, 0, 2.0])
    dirs = rng.normal(size=(m, 3)); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    keys = rng.integers(0, 2**62, size=m, dtype=np.uint64)
    a, _ = ctx.trace_rays(drt.make_opts(1, 2, 1.0), orig, dirs, keys, jac=False)
    b, _ = ctx.trace_rays(drt.make_opts(1, 2, 1.0, flags=drt.FLAG_IMAGE | drt.FLAG_NO_BVH), orig, dirs, keys, jac=False)
    assert np.array_equal(a, b)
    assert (a.sum(1) > 0).mean() > 0.02


def test_million_triangle_independent_seeds_agree_within_three_sigma(drt, ctx, million):
    """Independent streams on the 1 M-triangle scene: 8 x 8 pixel block means (2 048 samples each, close to
    Gaussian; single pixels at 32 spp are dominated by rare bright paths) agree within 3 sigma, sigma estimated
    from eight further independent renders."""
    ctx.upload(million)
    spp, mb = 32, 4
    runs = [ctx.render(drt.make_opts(spp, mb, 1.0, seed=s)) for s in range(1, 11)]
    imgs = np.stack([r[0] for r in runs])
    blocks = imgs.reshape(len(runs), 16, 8, 16, 8, 3).mean(axis=(2, 4))
    a, b = blocks[0], blocks[1]
    var = blocks[2:].var(axis=0, ddof=1)
    ok = np.abs(a - b) <= 3.0 * np.sqrt(2.0 * var) + 1e-12
    assert ok.mean() >= 0.95, ok.mean()
    m = imgs.reshape(len(runs), -1, 3).mean(1)
    assert (np.abs(m[0] - m[1:].mean(0)) <= 4.0 * m[1:].std(0, ddof=1) + 1e-5).all()
    # total per-triangle gradient mass (sum over triangles) agrees across seeds
    g = np.stack([r[1].sum(0) for r in runs])
    assert (np.abs(g[0] - g[1:].mean(0)) <= 4.0 * g[1:].std(0, ddof=1) + 1e-9 * np.abs(g[0])).all()
