"""Edge cases of the hot path on the GPU, each against the CPU oracle: scenes
that overflow the straight-line scan windows, degenerate image sizes, empty
and open scenes, extreme Russian-roulette settings."""
import numpy as np
import pytest

import oracle_lib
from oracle_lib import rel_err, restate_render

pytestmark = pytest.mark.gpu


def _check(drt, ctx, scene, opts_kw, tol=1e-9):
    ctx.upload(scene)
    img, grad, st = ctx.render(drt.make_opts(**opts_kw), stats=True)
    ref_img, ref_grad, ref_st = restate_render(scene, drt.make_opts(**opts_kw), want_stats=True)
    assert st.paths == ref_st.paths and st.segments == ref_st.segments and st.lit_paths == ref_st.lit_paths
    assert img.shape == ref_img.shape
    if img.size:
        assert rel_err(img, ref_img).max() <= tol
    assert rel_err(grad, ref_grad).max() <= tol
    return img, grad, st


def many_plane_scene(drt, W, H, n_planes=13, n_spheres=0):
    """A polygonal room of n_planes walls (more than the 8 straight-line plane
    slots of csrc/path.cuh), floor, ceiling, and optional spheres."""
    P = lambda v, n: drt.Param(np.asarray(v, dtype=np.float64), n)
    cols = [P((0.7, 0.3, 0.2), "c0"), P((0.2, 0.6, 0.3), "c1"), P((0.5, 0.5, 0.7), "c2")]
    lamp = P((6.0, 5.0, 4.0), "lamp")
    sc = drt.SceneDesc()
    for i in range(n_planes - 2):
        a = 2.0 * np.pi * (i + 0.37) / (n_planes - 2)
        sc.push_back(drt.Plane((-np.cos(a), 0.05 * ((i % 3) - 1), -np.sin(a)), -4.0 - 0.1 * i, drt.DiffuseBxDF(cols[i % 3])))
    sc.push_back(drt.Plane((0.0, 1.0, 0.0), -2.0, drt.DiffuseBxDF(cols[2])))
    sc.push_back(drt.Plane((0.0, -1.5, 0.0), -3.5, drt.DiffuseBxDF(cols[0])))        # non-unit
    for j in range(n_spheres):
        sc.push_back(drt.Sphere((-2.0 + 0.5 * j, -1.2 + 0.25 * (j % 4), 1.5 + 0.35 * j), 0.3, drt.DiffuseBxDF(cols[j % 3])))
    sc.push_back(drt.Sphere((0.3, 1.8, 1.0), 0.6, None, drt.AreaEmitter(lamp)))
    sc.camera = drt.Camera(W, H).look_at((0.0, 0.0, -1.0), (0.0, 0.0, 1.0))
    return sc


@pytest.mark.parametrize("n_planes,n_spheres", [(13, 0), (9, 9), (20, 11), (3, 0)])
@pytest.mark.parametrize("spp,mb,absorb", [(7, 5, 1.0), (33, 1, 0.3)])
def test_scan_window_overflow_matches_oracle(drt, ctx, n_planes, n_spheres, spp, mb, absorb):
    """More planes / spheres than compile-time slots: the rolled continuation loops."""
    _check(drt, ctx, many_plane_scene(drt, 36, 20, n_planes, n_spheres), dict(spp=spp, min_bounces=mb, absorb=absorb))


@pytest.mark.parametrize("W,H,spp", [(1, 1, 1), (1, 1, 300), (1, 37, 3), (53, 1, 2), (3, 2, 31), (5, 3, 97)])
def test_degenerate_image_sizes(drt, ctx, W, H, spp):
    _check(drt, ctx, drt.cornell_box(W, H), dict(spp=spp, min_bounces=3, absorb=0.4))
    _check(drt, ctx, drt.cornell_box(W, H), dict(spp=spp, min_bounces=8, absorb=1.0))


def test_empty_scene_renders_black_with_zero_gradients(drt, ctx):
    """Scene<T> with no shapes (pathtracer.hpp:72-89: raycast finds nothing, trace returns 0)."""
    sc = drt.SceneDesc()
    sc.camera = drt.Camera(17, 9).look_at((0, 0, 0), (0, 0, 1))
    ctx.upload(sc)
    img, grad, st = ctx.render(drt.make_opts(5, 2, 0.5), stats=True)
    assert img.shape == (9, 17, 3) and not img.any() and grad.shape == (0, 3)
    assert st.paths == 17 * 9 * 5 and st.segments == st.paths and st.lit_paths == 0


def test_open_scene_paths_escape(drt, ctx):
    """One plane, one light, nothing else: most paths miss (pathtracer.hpp:134-135)."""
    P = lambda v, n: drt.Param(np.asarray(v, dtype=np.float64), n)
    sc = drt.SceneDesc()
    sc.push_back(drt.Plane((0.0, 1.0, 0.0), -1.0, drt.DiffuseBxDF(P((0.6, 0.5, 0.4), "floor"))))
    sc.push_back(drt.Sphere((0.0, 1.0, 3.0), 0.5, None, drt.AreaEmitter(P((3, 3, 3), "lamp"))))
    sc.camera = drt.Camera(31, 23).look_at((0, 0, 0), (0, -0.2, 1))
    img, grad, st = _check(drt, ctx, sc, dict(spp=40, min_bounces=6, absorb=1.0))
    assert st.segments < 2.5 * st.paths and img.any()


@pytest.mark.parametrize("mb,absorb", [(0, 0.97), (20, 1.0), (40, 0.5)])
def test_extreme_roulette_settings(drt, ctx, mb, absorb):
    """absorb near 1 kills almost everything; min_bounces deeper than the 16-deep ring of the
    compacting kernels; min_bounces = 40 with absorb = 0.5 reaches past 40 vertices."""
    img, grad, st = _check(drt, ctx, drt.cornell_box(24, 16), dict(spp=6, min_bounces=mb, absorb=absorb))
    assert st.truncated_paths == 0


def test_unbounded_reference_recursion_against_the_record_capacity(drt, ctx):
    """absorb = 0 never kills: upstream a path recurses until it meets the light (~40 segments
    on average in the closed box, pathtracer.hpp:121-136 has no depth limit).  The device keeps
    at most 64 vertices per path (include/drtb.h max_depth) and counts what it cut; what lies
    beyond carries a throughput of 0.5^64, so image and gradients still agree with the
    unbounded oracle far inside the tolerance, while the segment counts must differ."""
    scene = drt.cornell_box(24, 16)
    ctx.upload(scene)
    opts = dict(spp=6, min_bounces=0, absorb=0.0)
    img, grad, st = ctx.render(drt.make_opts(**opts), stats=True)
    ref_img, ref_grad, ref_st = restate_render(scene, drt.make_opts(**opts), want_stats=True)
    assert 0 < st.truncated_paths < st.paths and st.segments < ref_st.segments
    assert rel_err(img, ref_img).max() <= 1e-9 and rel_err(grad, ref_grad).max() <= 1e-9
    # a tighter capacity is a different (truncated) estimator, and says so
    img8, _, st8 = ctx.render(drt.make_opts(max_depth=8, **opts), stats=True)
    assert st8.truncated_paths > st.truncated_paths and st8.segments <= 8 * st8.paths
    assert (img8 <= img + 1e-12).all() and img8.sum() < img.sum()


def test_large_seed_and_offset_keys_do_not_collide(drt, ctx):
    """Different seeds give different streams; the same seed the same image, bit for bit."""
    ctx.upload(drt.cornell_box(32, 24))
    a = ctx.render(drt.make_opts(8, 4, 1.0, seed=2**40 + 12345))[0]
    b = ctx.render(drt.make_opts(8, 4, 1.0, seed=2**40 + 12345))[0]
    c = ctx.render(drt.make_opts(8, 4, 1.0, seed=2**40 + 12346))[0]
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    ref = restate_render(drt.cornell_box(32, 24), drt.make_opts(8, 4, 1.0, seed=2**40 + 12345))[0]
    assert rel_err(a, ref).max() <= 1e-9


def _coincident_scene(drt, first, W=24, H=16):
    """Three surfaces that every camera ray meets at bit-identical t: the unit
    plane z = 4 (axis window of the scan), the same plane written with the
    normal (0, 0, -2) (general-plane window; 8 / (2 d_z) == 4 / d_z exactly) and
    a copy of one of them.  Each carries its own emitter, so the image tells which
    one the closest-hit scan kept: pathtracer.hpp:80 keeps the first in scene order."""
    P = lambda v, n: drt.Param(np.asarray(v, dtype=np.float64), n)
    shapes = {
        "axis": drt.Plane((0.0, 0.0, -1.0), -4.0, None, drt.AreaEmitter(P((1.0, 0.0, 0.0), "axis"))),
        "general": drt.Plane((0.0, 0.0, -2.0), -8.0, None, drt.AreaEmitter(P((0.0, 1.0, 0.0), "general"))),
        "axis_copy": drt.Plane((0.0, 0.0, -1.0), -4.0, None, drt.AreaEmitter(P((0.0, 0.0, 1.0), "axis_copy"))),
        "general_copy": drt.Plane((0.0, 0.0, -2.0), -8.0, None, drt.AreaEmitter(P((1.0, 1.0, 0.0), "general_copy"))),
    }
    sc = drt.SceneDesc()
    for name in first:
        sc.push_back(shapes[name])
    sc.camera = drt.Camera(W, H).look_at((0.0, 0.0, 0.0), (0.0, 0.0, 1.0))
    return sc, np.asarray(shapes[first[0]].emitter.emission.value)


@pytest.mark.parametrize("order", [("axis", "general", "axis_copy"), ("general", "axis", "general_copy"),
                                   ("axis_copy", "axis", "general"), ("general_copy", "general", "axis")])
@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_exact_ties_keep_the_lower_scene_index(drt, ctx, order, precision):
    scene, colour = _coincident_scene(drt, order)
    kw = dict(spp=4, min_bounces=2, absorb=1.0)
    ctx.upload(scene)
    img, grad = ctx.render(drt.make_opts(precision=drt.F64 if precision == "f64" else drt.F32, **kw))
    ref_img, ref_grad = restate_render(scene, drt.make_opts(**kw))
    assert np.array_equal(ref_img, np.broadcast_to(colour, ref_img.shape))       # the oracle keeps scene-order shape 0
    assert np.array_equal(img, ref_img)
    assert np.array_equal(grad, ref_grad)


@pytest.mark.parametrize("mesh", [False, True])
def test_shard_that_owns_no_rows_reports_zeros(drt, ctx, mesh):
    """More shards than bands (H = 32, bands of 8 rows, 8 shards): shards 4..7 own nothing.  Their gradients
    must be OVERWRITTEN with zeros (drtb.h), not left stale, and a mesh scene must not divide by zero
    (drtb.cu reduce_partials / launch_wavefront)."""
    scene = drt.tessellated_room(2, 4, width=16, height=32) if mesh else drt.cornell_box(16, 32)
    ctx.upload(scene)
    whole_img, whole_grad = ctx.render(drt.make_opts(4, 2, 0.5))           # leaves non-zero gradients behind
    assert np.abs(whole_grad).max() > 0
    total = np.zeros_like(whole_grad)
    for s in range(8):
        o = drt.make_opts(4, 2, 0.5, shard_index=s, shard_count=8, band_rows=8)
        img, grad, st = ctx.render(o, stats=True)
        if s >= 4:
            assert img.shape[0] == 0 and st.segments == 0 and st.lit_paths == 0 and st.paths == 0
            assert np.array_equal(grad, np.zeros_like(grad))
        else:
            assert np.array_equal(img, whole_img[8 * s:8 * s + 8])
        total += grad
    assert np.abs(total - whole_grad).max() <= 1e-12 * np.abs(whole_grad).max()


def test_reserve_makes_the_first_render_as_fast_as_the_second(drt, ctx):
    """drtb_reserve / the preparation pass inside drtb_render: no allocation, attribute call or first-use
    kernel load lands between the events behind drtb_stats.kernel_ms."""
    with drt.Context(0) as fresh:
        fresh.upload(drt.cornell_box(96, 64))
        times = {}
        for spp, mb, ab in ((8, 8, 1.0), (8, 1, 0.5), (32, 8, 1.0), (32, 1, 0.5), (40, 16, 1.0)):
            first = fresh.render(drt.make_opts(spp, mb, ab), stats=True)[2].kernel_ms
            second = fresh.render(drt.make_opts(spp, mb, ab), stats=True)[2].kernel_ms
            times[(spp, mb, ab)] = (first, second)
            assert first <= 3.0 * second + 0.25, times


@pytest.mark.parametrize("opts_kw", [dict(spp=32, min_bounces=4, absorb=1.0), dict(spp=8, min_bounces=4, absorb=1.0),
                                     dict(spp=16, min_bounces=1, absorb=0.5),
                                     dict(spp=32, min_bounces=4, absorb=1.0, shard_index=1, shard_count=3)])
@pytest.mark.parametrize("precision", ["F64", "F32", "MIXED"])
def test_pinned_host_image_is_written_by_the_kernel_and_equals_the_copied_one(drt, ctx, opts_kw, precision):
    """drtb_render with a PINNED host image lets the analytic-scene kernels store the pixels straight into it (no
    D2H copy; DRTB_MIXED and mesh scenes keep the copy): the bits must be those of the pageable-buffer path."""
    import torch
    scene = drt.cornell_box(96, 64)
    ctx.upload(scene)
    o = drt.make_opts(precision=getattr(drt, precision), **opts_kw)
    img, grad = ctx.render(o)
    rows = drt.shard_rows(64, o.shard_index, max(1, o.shard_count), max(1, o.band_rows))
    h_img = torch.full((rows, 96, 3), float("nan"), dtype=torch.float64).pin_memory()
    h_grad = torch.full((scene.n_params, 3), float("nan"), dtype=torch.float64).pin_memory()
    ctx.render_host_ptrs(o, 0, h_img.data_ptr(), h_grad.data_ptr())
    if precision == "MIXED":          # the re-trace list is filled in a different order every run: same sums to rounding
        assert rel_err(h_img.numpy(), img).max() <= 1e-12 and rel_err(h_grad.numpy(), grad).max() <= 1e-12
    else:
        assert np.array_equal(h_img.numpy(), img) and np.array_equal(h_grad.numpy(), grad)


def test_host_alloc_gives_a_pinned_image_the_kernel_writes(drt, ctx):
    """drtb_host_alloc / drtb_host_free: pinned memory for C callers without the CUDA headers."""
    import ctypes as C
    from drt_b200 import abi
    lib = abi.load_library()
    scene = drt.cornell_box(64, 48)
    ctx.upload(scene)
    o = drt.make_opts(32, 4, 1.0)
    img, grad = ctx.render(o)
    p = C.c_void_p()
    assert lib.drtb_host_alloc(img.nbytes, C.byref(p)) == 0 and p.value
    try:
        view = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(img.size,))
        view[:] = np.nan
        g2 = np.empty_like(grad)
        ctx.render_host_ptrs(o, 0, p.value, g2.ctypes.data)
        assert np.array_equal(view.reshape(img.shape), img) and np.array_equal(g2, grad)
    finally:
        assert lib.drtb_host_free(p) == 0
    assert lib.drtb_host_free(None) == 0 and lib.drtb_host_alloc(0, C.byref(p)) == abi.ERR_INVALID
