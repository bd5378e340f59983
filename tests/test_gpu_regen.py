"""The path-regenerating kernel (render_regen_kernel): Russian-roulette renders of all-diffuse
analytic scenes.  Every lane steps one segment per iteration and free lanes take the next samples
of the warp's chunk of pixels, so the pixel sums are taken in another order than in the pass-based
kernel -- results agree to rounding, path statistics exactly."""
import os

import numpy as np
import pytest

from oracle_lib import rel_err, restate_render

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx_passes(drt):
    """A second context with regeneration switched off (DRTB_NO_REGEN is read by drtb_create)."""
    os.environ["DRTB_NO_REGEN"] = "1"
    try:
        c = drt.Context(0)
    finally:
        del os.environ["DRTB_NO_REGEN"]
    yield c
    c.close()


def _stats(st):
    return (st.paths, st.segments, st.lit_paths, st.truncated_paths)


@pytest.mark.parametrize("spp,mb,absorb,max_depth", [
    (33, 40, 0.5, 0),      # records deeper than the 16-entry ring: swept by their own lane
    (32, 0, 0.0, 0),       # nothing is ever absorbed: every unlit path is cut at 64 vertices
    (40, 0, 0.3, 8),       # a caller-imposed depth limit, min_bounces = 0 (roulette before the first segment)
    (256, 1, 0.5, 0),      # the reference's defaults, 8 refills per pixel
    (37, 2, 0.9, 0),       # almost everything absorbed at depth 2
])
def test_regeneration_equals_the_pass_based_kernel(drt, ctx, ctx_passes, spp, mb, absorb, max_depth):
    scene = drt.cornell_box(24, 16)
    ctx.upload(scene); ctx_passes.upload(scene)
    kw = dict(max_depth=max_depth) if max_depth else {}
    a = ctx.render(drt.make_opts(spp, mb, absorb, **kw), stats=True)
    b = ctx_passes.render(drt.make_opts(spp, mb, absorb, **kw), stats=True)
    assert _stats(a[2]) == _stats(b[2])
    assert rel_err(a[0], b[0]).max() <= 1e-12 and rel_err(a[1], b[1]).max() <= 1e-12
    assert np.array_equal(a[0] == 0.0, b[0] == 0.0)
    again = ctx.render(drt.make_opts(spp, mb, absorb, **kw))
    assert np.array_equal(a[0], again[0]) and np.array_equal(a[1], again[1])      # deterministic refill order


def test_regeneration_with_more_than_eight_parameters(drt, ctx):
    """Shared atomic gradient columns (9 .. 64 parameters) under the regenerating kernel."""
    P = lambda v, n: drt.Param(np.asarray(v, dtype=np.float64), n)
    sc = drt.SceneDesc()
    rng = np.random.default_rng(5)
    for i in range(10):
        sc.push_back(drt.Sphere((-2.5 + 0.55 * i, -1.0 + 0.3 * (i % 3), 3.0 + 0.2 * i), 0.45,
                                drt.DiffuseBxDF(P(rng.uniform(0.2, 0.9, 3), f"c{i}"))))
    sc.push_back(drt.Plane((0.0, 1.0, 0.0), -2.0, drt.DiffuseBxDF(P((0.5, 0.5, 0.5), "floor"))))
    sc.push_back(drt.Plane((0.0, 0.0, -1.0), -7.0, drt.DiffuseBxDF(P((0.4, 0.6, 0.5), "back"))))
    sc.push_back(drt.Sphere((0.0, 4.0, 3.0), 1.5, None, drt.AreaEmitter(P((5, 5, 5), "lamp"))))
    sc.camera = drt.Camera(40, 28).look_at((0, 0, 0), (0, 0, 1))
    ctx.upload(sc)
    assert sc.n_params == 13
    img, grad, st = ctx.render(drt.make_opts(36, 1, 0.4), stats=True)
    ref_img, ref_grad, ref_st = restate_render(sc, drt.make_opts(36, 1, 0.4), want_stats=True)
    assert (st.paths, st.segments, st.lit_paths) == (ref_st.paths, ref_st.segments, ref_st.lit_paths)
    assert rel_err(img, ref_img).max() <= 1e-9 and rel_err(grad, ref_grad).max() <= 1e-9


def test_regeneration_with_a_seed_image_and_shards(drt, ctx):
    """Per-pixel adjoint seeds and row-band shards go through the regenerating kernel unchanged."""
    W, H, spp = 32, 24, 48
    scene = drt.cornell_box(W, H)
    ctx.upload(scene)
    rng = np.random.default_rng(11)
    seed = rng.normal(size=(H, W, 3))
    o = drt.make_opts(spp, 1, 0.5, seed_scale=1.0 / spp)
    img, grad = ctx.render(o, seed_img=seed)
    ref_img, ref_grad = restate_render(scene, drt.make_opts(spp, 1, 0.5, seed_scale=1.0 / spp), seed_img=seed)
    assert rel_err(img, ref_img).max() <= 1e-9 and rel_err(grad, ref_grad).max() <= 1e-9
    from differentiable_renderer_b200 import sharding
    parts, gsum = [], 0.0
    for r in range(3):
        ys = sharding.shard_row_indices(H, r, 3, 4)
        si, sg = ctx.render(drt.make_opts(spp, 1, 0.5, seed_scale=1.0 / spp, shard_index=r, shard_count=3, band_rows=4),
                            seed_img=seed[ys])
        parts.append(si); gsum = gsum + sg
    assert np.array_equal(sharding.assemble_image(parts, H, 4), img)          # a pixel does not care who renders it
    assert rel_err(gsum, grad).max() <= 1e-12
