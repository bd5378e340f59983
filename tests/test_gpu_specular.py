"""GPU parity for the two §8(f) rows built on the general kernel variant:
SpecularBxDF (reference bxdf.hpp:85-124, reflect() vector.hpp:602-606) and the
per-pixel gradient image of one parameter (README.md:138-145), through the C
ABI against the reference's golden vectors and the CPU oracle."""
from pathlib import Path

import numpy as np
import pytest

import oracle_lib
from oracle_lib import rel_err, restate_render

pytestmark = pytest.mark.gpu

GOLDEN = Path(__file__).resolve().parent / "golden"
IMG_TOL, GRAD_TOL = 1e-4, 1e-3          # north-star tolerances; double lands ~1e-12


@pytest.mark.parametrize("name", ["specbox_40x28_6spp_b4_p1", "specbox_40x28_6spp_b1_p05", "specbox_24x16_40spp_b3_p03"])
def test_specular_matches_reference_golden_vectors(drt, ctx, name):
    """Vectors produced by the reference's own SpecularBxDF; 6 spp runs the
    direct sweeps, 40 spp the lit-path ring."""
    z = np.load(GOLDEN / f"{name}.npz")
    W, H, spp, mb, ab, seed, _ = z["meta"]
    scene = drt.specular_box(int(W), int(H))
    ctx.upload(scene)
    opts = drt.make_opts(int(spp), int(mb), float(ab), seed=int(seed))
    img, grad, gimg = ctx.render_grad_image(opts, scene.params[4])
    assert rel_err(img, z["img"]).max() <= 1e-9 <= IMG_TOL
    assert rel_err(grad, z["grad"]).max() <= 1e-9 <= GRAD_TOL
    assert np.abs(gimg - z["gimg_gloss"]).max() <= 1e-9 * np.abs(z["gimg_gloss"]).max()
    # plain drtb_render of the same scene gives the same image and gradients
    img2, grad2 = ctx.render(drt.make_opts(int(spp), int(mb), float(ab), seed=int(seed)))
    assert np.array_equal(img2, img) and rel_err(grad2, grad).max() <= 1e-12


@pytest.mark.parametrize("mb,absorb,spp", [(5, 1.0, 8), (2, 0.4, 33), (0, 0.2, 5)])
def test_specular_matches_oracle_with_counts(drt, ctx, mb, absorb, spp):
    scene = drt.specular_box(44, 30, exponent_ball=50.0, exponent_wall=3.0)
    ctx.upload(scene)
    img, grad, st = ctx.render(drt.make_opts(spp, mb, absorb, seed=9), stats=True)
    ref_img, ref_grad, ref_st = restate_render(scene, drt.make_opts(spp, mb, absorb, seed=9), want_stats=True)
    assert st.segments == ref_st.segments and st.lit_paths == ref_st.lit_paths
    assert rel_err(img, ref_img).max() <= 1e-9
    assert rel_err(grad, ref_grad).max() <= 1e-9
    assert grad[4].min() > 0                       # the gloss tint does receive gradient


def test_specular_non_integer_exponent_nan_pixels_match_the_oracle(drt, ctx):
    """pow(negative, non-integer) = NaN upstream (bxdf.hpp:102, a lobe seen from behind):
    the same pixels are NaN here and every other pixel agrees."""
    scene = drt.specular_box(36, 20, exponent_ball=50.0, exponent_wall=1.5)
    ctx.upload(scene)
    img, _ = ctx.render(drt.make_opts(5, 5, 1.0, seed=4))
    ref_img, _ = restate_render(scene, drt.make_opts(5, 5, 1.0, seed=4))
    assert np.isnan(ref_img).any() and np.array_equal(np.isnan(img), np.isnan(ref_img))
    ok = ~np.isnan(ref_img)
    assert rel_err(img[ok], ref_img[ok]).max() <= 1e-9


def test_specular_float_instantiation_is_statistically_the_same(drt, ctx):
    scene = drt.specular_box(64, 48)
    ctx.upload(scene)
    a_img, a_grad = ctx.render(drt.make_opts(64, 4, 1.0))
    b_img, b_grad = ctx.render(drt.make_opts(64, 4, 1.0, precision=drt.F32))
    assert rel_err(b_grad, a_grad).max() <= GRAD_TOL
    bad = (rel_err(b_img, a_img) > IMG_TOL).any(axis=2).mean()
    assert bad <= 0.02


def test_specular_explicit_rays(drt, ctx):
    """drtb_trace_rays on a specular scene against the oracle's trace."""
    import ctypes as C
    scene = drt.specular_box(8, 8)
    ctx.upload(scene)
    rng = np.random.default_rng(5)
    n = 1500
    orig = rng.uniform(-1.5, 1.5, size=(n, 3)) + np.array([0, 0, 2.0])
    dirs = rng.normal(size=(n, 3)); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    dirs[:, 2] = np.abs(dirs[:, 2])                # towards the specular sphere / back wall
    keys = rng.integers(0, 2**62, size=n, dtype=np.uint64)
    opts = drt.make_opts(1, 6, 1.0)
    rad, jac = ctx.trace_rays(opts, orig, dirs, keys)
    lib = oracle_lib.load_restate()
    sc = scene.flatten()
    r_rad = np.zeros((n, 3)); r_jac = np.zeros((n, scene.n_params, 3))
    dp = C.POINTER(C.c_double)
    assert lib.drt_oracle_trace_rays(C.byref(sc), C.byref(opts), n, orig.ctypes.data_as(dp), dirs.ctypes.data_as(dp),
                                     keys.ctypes.data_as(C.POINTER(C.c_uint64)), r_rad.ctypes.data_as(dp),
                                     r_jac.ctypes.data_as(dp)) == 0
    assert rel_err(rad, r_rad).max() <= 1e-9 and rel_err(jac, r_jac).max() <= 1e-9
    assert np.abs(r_jac[:, 4]).max() > 0


# ---- gradient image -------------------------------------------------------------------
def test_gradient_image_matches_reference_golden(drt, ctx):
    z = np.load(GOLDEN / "cbox_48x32_8spp_b8_p1_gimg_red.npz")
    scene = drt.cornell_box(48, 32)
    ctx.upload(scene)
    img, grad, gimg = ctx.render_grad_image(drt.make_opts(8, 8, 1.0, seed_scale=1.0 / 8), 0)
    assert rel_err(img, z["img"]).max() <= 1e-9
    assert rel_err(grad, z["grad"]).max() <= 1e-9
    assert np.abs(gimg - z["gimg"]).max() <= 1e-12 * max(1.0, np.abs(z["gimg"]).max())
    assert np.array_equal(gimg == 0.0, z["gimg"] == 0.0)


@pytest.mark.parametrize("spp,k", [(8, 2), (40, 3), (5, 1), (64, 0)])
def test_gradient_image_sums_to_the_gradient(drt, ctx, spp, k):
    """sum over pixels of grad_img == grad[param]: both layouts (pixels per warp /
    passes per pixel), with a per-pixel seed image, and for a sharded render."""
    scene = drt.cornell_box(40, 24)
    ctx.upload(scene)
    rng = np.random.default_rng(spp)
    seed_img = rng.uniform(0.5, 1.5, size=(24, 40, 3))
    img, grad, gimg = ctx.render_grad_image(drt.make_opts(spp, 3, 0.3), k, seed_img=seed_img)
    assert rel_err(gimg.sum((0, 1)), grad[k]).max() <= 1e-12
    ref = restate_render(scene, drt.make_opts(spp, 3, 0.3), seed_img=seed_img, grad_image_of=k)
    assert rel_err(img, ref[0]).max() <= 1e-9 and rel_err(grad, ref[1]).max() <= 1e-9
    assert np.abs(gimg - ref[2]).max() <= 1e-9 * np.abs(ref[2]).max()
    parts = [ctx.render_grad_image(drt.make_opts(spp, 3, 0.3, shard_index=r, shard_count=3, band_rows=4), k)
             for r in range(3)]
    whole = ctx.render_grad_image(drt.make_opts(spp, 3, 0.3), k)
    from differentiable_renderer_b200 import sharding
    assert np.array_equal(sharding.assemble_image([p[2] for p in parts], 24, 4), whole[2])


def test_gradient_image_with_decorrelated_adjoint_and_errors(drt, ctx):
    scene = drt.cornell_box(32, 24)
    ctx.upload(scene)
    img, grad, gimg = ctx.render_grad_image(drt.make_opts(8, 4, 1.0, adjoint_seed=77), 2)
    _, g2 = ctx.render(drt.make_opts(8, 4, 1.0, seed=77, flags=drt.FLAG_GRAD))
    assert rel_err(grad, g2).max() <= 1e-12 and rel_err(gimg.sum((0, 1)), grad[2]).max() <= 1e-12
    with pytest.raises(drt.DrtbError):
        ctx.render_grad_image(drt.make_opts(8, 4, 1.0), 17)                         # no such parameter
    with pytest.raises(drt.DrtbError):
        ctx.render_grad_image(drt.make_opts(8, 4, 1.0, flags=drt.FLAG_IMAGE), 0)    # needs DRTB_FLAG_GRAD


def test_many_parameter_scene_uses_the_atomic_sink(drt, ctx):
    """> 8 parameters: the general kernel with red.global.add gradients."""
    P = lambda v, n: drt.Param(np.asarray(v, dtype=np.float64), n)
    sc = drt.SceneDesc()
    rng = np.random.default_rng(3)
    for i in range(10):                             # ten spheres, ten albedos, two of them specular
        col = P(rng.uniform(0.2, 0.9, 3), f"c{i}")
        bx = drt.SpecularBxDF(col, 3.0 + i) if i in (2, 7) else drt.DiffuseBxDF(col)
        sc.push_back(drt.Sphere((-2.5 + 0.55 * i, -1.0 + 0.3 * (i % 3), 3.0 + 0.2 * i), 0.45, bx))
    sc.push_back(drt.Plane((0.0, 1.0, 0.0), -2.0, drt.DiffuseBxDF(P((0.5, 0.5, 0.5), "floor"))))
    sc.push_back(drt.Plane((0.0, 0.0, -1.0), -7.0, drt.DiffuseBxDF(P((0.4, 0.6, 0.5), "back"))))
    sc.push_back(drt.Sphere((0.0, 4.0, 3.0), 1.5, None, drt.AreaEmitter(P((5, 5, 5), "lamp"))))
    sc.camera = drt.Camera(40, 28).look_at((0, 0, 0), (0, 0, 1))
    ctx.upload(sc)
    assert sc.n_params == 13
    for spp in (6, 36):
        img, grad, gimg = ctx.render_grad_image(drt.make_opts(spp, 4, 1.0), 7)
        ref = restate_render(sc, drt.make_opts(spp, 4, 1.0), grad_image_of=7)
        assert rel_err(img, ref[0]).max() <= 1e-9 and rel_err(grad, ref[1]).max() <= 1e-9
        assert np.abs(gimg - ref[2]).max() <= 1e-9 * np.abs(ref[2]).max()


# ---- mesh scenes (wavefront) -----------------------------------------------------------
def test_mesh_scene_with_a_specular_analytic_primitive_and_gradient_image(drt, ctx):
    sc = drt.tessellated_room(2, 4, width=32, height=24)
    tint = drt.Param(np.array([0.9, 0.8, 0.5]), "tint")
    sc.push_back(drt.Sphere((-1.2, -2.0, 2.5), 1.0, drt.SpecularBxDF(tint, 12.0)))
    sc.push_back(drt.Plane((0.0, 1.0, 0.0), -2.5, drt.DiffuseBxDF(drt.Param(np.array([0.3, 0.4, 0.5]), "floor"))))
    ctx.upload(sc)
    k = tint.index
    for spp in (4, 34):
        img, grad, gimg = ctx.render_grad_image(drt.make_opts(spp, 4, 1.0), k)
        ref = restate_render(sc, drt.make_opts(spp, 4, 1.0), grad_image_of=k)
        assert rel_err(img, ref[0]).max() <= 1e-9 and rel_err(grad, ref[1]).max() <= 1e-9
        assert np.abs(gimg - ref[2]).max() <= 1e-9 * np.abs(ref[2]).max()
        assert np.abs(ref[2]).max() > 0
