"""CPU tests of the CHECKER: the plain-C restatement (oracle/restate.c) against
golden vectors produced by the unmodified reference, against the survey's
known-answer vectors, and -- where /root/reference (or its prebuilt
oracle/_ref/libdrt_ref.so) exists -- against the reference itself."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

import oracle_lib
from oracle_lib import drt, ref_render, rel_err, restate_render

GOLDEN = Path(__file__).resolve().parent / "golden"
needs_ref = pytest.mark.skipif(not oracle_lib.have_ref(), reason="oracle/_ref not built and /root/reference absent")

COUNTER_CASES = ["cbox_48x32_8spp_b8_p1", "cbox_48x32_8spp_b1_p05", "cbox_48x32_8spp_b3_p03",
                 "cbox_40x24_5spp_b0_p025_s7", "cbox_32x32_40spp_b4_p1"]


def load_case(name):
    z = np.load(GOLDEN / f"{name}.npz")
    W, H, spp, mb, ab, seed, mode = z["meta"]
    return z, int(W), int(H), int(spp), int(mb), float(ab), int(seed), int(mode)


@pytest.mark.parametrize("name", COUNTER_CASES)
def test_restatement_matches_golden_bit_for_bit(name):
    z, W, H, spp, mb, ab, seed, _ = load_case(name)
    img, grad = restate_render(drt.cornell_box(W, H), drt.make_opts(spp, mb, ab, seed=seed))
    assert np.array_equal(img, z["img"])
    assert np.array_equal(grad, z["grad"])


def test_restatement_seed_image_golden():
    z, W, H, spp, mb, ab, seed, _ = load_case("cbox_32x24_6spp_b2_p04_seedimg")
    img, grad = restate_render(drt.cornell_box(W, H), drt.make_opts(spp, mb, ab, seed_scale=1.0 / spp),
                               seed_img=z["seed_img"])
    assert np.array_equal(img, z["img"])
    assert rel_err(grad, z["grad"]).max() < 1e-13


def test_restatement_threads_do_not_change_the_image():
    scene = drt.cornell_box(40, 24)
    a_img, a_grad = restate_render(scene, drt.make_opts(6, 2, 0.4), threads=1)
    b_img, b_grad = restate_render(scene, drt.make_opts(6, 2, 0.4), threads=4)
    assert np.array_equal(a_img, b_img)
    assert rel_err(b_grad, a_grad).max() < 1e-13


def test_restatement_explicit_rays_golden():
    z = np.load(GOLDEN / "rays_64_b3_p03.npz")
    lib = oracle_lib.load_restate()
    scene = drt.cornell_box(8, 8)
    sc = scene.flatten()
    opts = drt.make_opts(1, 3, 0.3)
    n = z["orig"].shape[0]
    rad = np.zeros((n, 3)); jac = np.zeros((n, len(scene.params), 3))
    dp = C.POINTER(C.c_double)
    orig, dirs, keys = (np.ascontiguousarray(z[k]) for k in ("orig", "dirs", "keys"))
    rc = lib.drt_oracle_trace_rays(C.byref(sc), C.byref(opts), n, orig.ctypes.data_as(dp), dirs.ctypes.data_as(dp),
                                   keys.ctypes.data_as(C.POINTER(C.c_uint64)), rad.ctypes.data_as(dp),
                                   jac.ctypes.data_as(dp))
    assert rc == 0
    assert np.array_equal(rad, z["radiance"])
    assert rel_err(jac, z["jac"]).max() < 1e-13


# ---- survey known-answer vectors, 9 printed decimals ----------
KATS = {
    (8, 1.0): [[898.343750000, 888.701315078, 768.609375000], [704.396093750, 691.322352896, 583.843125000],
               [1704.328125000, 1575.104482379, 1048.875000000], [1850.101562500, 1752.005372524, 1434.281250000]],
    (1, 0.5): [[958, 971.6412, 818], [755.48, 736.857418, 595.9], [1722, 1655.637618, 1044], [1874, 1762.643903, 1431]],
    (3, 0.3): [[915.084459426, 898.078004125, 776.296564505], [699.148856262, 699.590397862, 578.493447586],
               [1756.999610821, 1602.751586762, 1059.969666813], [1856.925590574, 1753.577212404, 1435.552504148]],
}


@pytest.mark.parametrize("setting", list(KATS))
def test_survey_kat_gradients(setting):
    mb, ab = setting
    _, grad = restate_render(drt.cornell_box(96, 64), drt.make_opts(8, mb, ab))
    assert np.abs(grad - np.array(KATS[setting])).max() < 1e-6


def test_survey_kat4_mean_radiance():
    img, _ = restate_render(drt.cornell_box(64, 64), drt.make_opts(16, 8, 1.0))
    assert np.abs(img.reshape(-1, 3).mean(0) - np.array([0.050204, 0.047712, 0.044227])).max() < 1e-6


# ---- against the unmodified reference ---------------------------------------------
@needs_ref
@pytest.mark.parametrize("name", COUNTER_CASES + ["cbox_64x64_16spp_b1_p05_libc"])
def test_reference_reproduces_golden(name):
    z, W, H, spp, mb, ab, seed, mode = load_case(name)
    img, grad = ref_render(drt.cornell_box(W, H), drt.make_opts(spp, mb, ab, seed=seed), rand_mode=mode)
    assert np.array_equal(img, z["img"])
    assert np.array_equal(grad, z["grad"])


@needs_ref
def test_survey_kat5_as_shipped_libc_stream():
    """Sequential unseeded glibc rand(), loop order of src/render.cpp:72-76."""
    z, *_ = load_case("cbox_64x64_16spp_b1_p05_libc")
    want = np.array([[852, 783.020600, 686], [505, 471.591018, 395.92], [2908, 2715.473220, 2074],
                     [3273, 3104.332303, 2889]])
    assert np.abs(z["grad"] - want).max() < 1e-5
    assert np.abs(z["img"].reshape(-1, 3).mean(0) - np.array([0.049942, 0.047368, 0.044083])).max() < 1e-6


@needs_ref
@pytest.mark.parametrize("mb,ab,spp,seed", [(2, 0.7, 3, 1), (6, 0.1, 2, 2), (0, 0.9, 9, 3), (12, 1.0, 2, 4)])
def test_restatement_equals_reference(mb, ab, spp, seed):
    scene = drt.cornell_box(36, 20, red=(0.7, 0.1, 0.0), green=(0.2, 0.6, 0.3), white=(0.4, 0.5, 0.6),
                            emission=(2.0, 1.5, 0.5))
    a_img, a_grad = ref_render(scene, drt.make_opts(spp, mb, ab, seed=seed), threads=2)
    b_img, b_grad = restate_render(scene, drt.make_opts(spp, mb, ab, seed=seed))
    assert np.array_equal(a_img, b_img)
    assert rel_err(b_grad, a_grad).max() < 1e-13


@needs_ref
def test_reference_shards_tile_the_image():
    scene = drt.cornell_box(24, 22)
    full, g = ref_render(scene, drt.make_opts(2, 2, 0.5))
    parts = [ref_render(scene, drt.make_opts(2, 2, 0.5, shard_index=r, shard_count=3, band_rows=4)) for r in range(3)]
    from differentiable_renderer_b200 import sharding
    assert np.array_equal(sharding.assemble_image([p[0] for p in parts], 22, 4), full)
    assert rel_err(sum(p[1] for p in parts), g).max() < 1e-13


# ---- independent gradient check ----------------------------------------------------
def test_gradients_are_exact_derivatives_of_the_fixed_stream_estimator():
    """RR never looks at throughput (pathtracer.hpp:128-130), so with the stream
    held fixed the estimator is a polynomial in every parameter and central
    differences are exact to O(h^2)."""
    W, H, spp, mb, ab, h = 32, 24, 4, 8, 1.0, 1e-4
    base = dict(red=[0.5, 0, 0], green=[0, 0.5, 0], white=[0.5, 0.5, 0.5], emission=[1, 1, 1])
    _, grad = restate_render(drt.cornell_box(W, H, **base), drt.make_opts(spp, mb, ab))
    names = ["red", "green", "white", "emission"]
    for k, c in [(0, 0), (0, 1), (1, 2), (2, 2), (3, 1)]:
        tot = []
        for sgn in (+1, -1):
            p = {n: list(v) for n, v in base.items()}
            p[names[k]][c] += sgn * h
            img, _ = restate_render(drt.cornell_box(W, H, **p), drt.make_opts(spp, mb, ab))
            tot.append(img[..., c].sum() * spp)
        fd = (tot[0] - tot[1]) / (2 * h)
        assert abs(fd - grad[k, c]) <= 1e-6 * max(1.0, abs(grad[k, c]))


def test_radiance_is_linear_in_emission():
    a, ga = restate_render(drt.cornell_box(24, 16), drt.make_opts(4, 3, 0.3))
    b, gb = restate_render(drt.cornell_box(24, 16, emission=(2, 2, 2)), drt.make_opts(4, 3, 0.3))
    assert np.allclose(b, 2 * a, rtol=1e-14, atol=0)
    # Euler: sum_c E_c dL/dE_c = L  ->  emission.grad . emission = spp * sum(img)
    assert np.allclose((ga[3] * 1.0), a.reshape(-1, 3).sum(0) * 4, rtol=1e-12)


# ---- the sample stream ---------------------------------------------------------------
def py_draw(key, slot):
    M = (1 << 64) - 1
    x = (key * 0x100000001B3 + slot) & M
    x = (x + 0x9E3779B97F4A7C15) & M
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M
    x ^= x >> 31
    return x % 2147483647


def test_stream_is_the_documented_hash_and_stays_below_rand_max():
    lib = oracle_lib.load_restate()
    prod = drt.load_library()                      # drtb_stream_draw needs no GPU
    rng = np.random.default_rng(7)
    keys = [0, 1, 2**63, 2**64 - 1] + [int(k) for k in rng.integers(0, 2**63, size=200)]
    for key in keys:
        for slot in (0, 1, 2, 17, 1000, 2**32 - 1):
            want = py_draw(key, slot)
            assert want <= 2147483646
            assert lib.drt_oracle_stream_draw(key, slot) == want
            assert prod.drtb_stream_draw(key, slot) == want


# ---- triangle meshes (new functionality; semantics in include/drtb.h) ---------------
def mixed_mesh_scene(W=32, H=24):
    """A tessellated room WITH analytic primitives in front of it: exercises the
    analytic-then-triangle scene order."""
    sc = drt.tessellated_room(2, 4, width=W, height=H)
    ball = drt.DiffuseBxDF(drt.Param(np.array([0.7, 0.6, 0.2]), "ball"))
    sc.push_back(drt.Sphere((-1.2, -2.0, 2.5), 1.0, ball))
    sc.push_back(drt.Plane((0.0, 1.0, 0.0), -2.5, drt.DiffuseBxDF(drt.Param(np.array([0.3, 0.4, 0.5]), "floor"))))
    return sc


@needs_ref
@pytest.mark.parametrize("mb,ab", [(4, 1.0), (1, 0.4)])
def test_mesh_restatement_equals_reference_with_triangle_shape(mb, ab):
    """The reference's own raycast/scatter/tape over a test-only Triangle<T> shape
    (oracle/ref_oracle.cpp) against the flat mesh extension of oracle/restate.c."""
    for scene in (drt.tessellated_room(2, 4, width=28, height=20), mixed_mesh_scene()):
        a_img, a_grad = ref_render(scene, drt.make_opts(3, mb, ab, seed=3))
        b_img, b_grad = restate_render(scene, drt.make_opts(3, mb, ab, seed=3))
        assert np.array_equal(a_img, b_img)
        assert rel_err(b_grad, a_grad).max() < 1e-13
        assert a_img.max() > 0


def test_mesh_restatement_matches_golden():
    z = np.load(GOLDEN / "mesh_room_98tri_32x24_4spp_b4.npz")
    img, grad = restate_render(drt.tessellated_room(2, 4, width=32, height=24), drt.make_opts(4, 4, 1.0))
    assert np.array_equal(img, z["img"]) and rel_err(grad, z["grad"]).max() < 1e-13


# ---- SpecularBxDF and the gradient image -------------------------
SPEC_CASES = ["specbox_40x28_6spp_b4_p1", "specbox_40x28_6spp_b1_p05", "specbox_24x16_40spp_b3_p03"]


@pytest.mark.parametrize("name", SPEC_CASES)
def test_specular_restatement_matches_reference_golden(name):
    """Golden vectors come from the reference's own SpecularBxDF (bxdf.hpp:85-124)."""
    z, W, H, spp, mb, ab, seed, _ = load_case(name)
    img, grad, gimg = restate_render(drt.specular_box(W, H), drt.make_opts(spp, mb, ab, seed=seed), grad_image_of=4)
    assert np.array_equal(img, z["img"])
    assert rel_err(grad, z["grad"]).max() < 1e-13
    assert np.abs(gimg - z["gimg_gloss"]).max() <= 1e-13 * np.abs(z["gimg_gloss"]).max()


@needs_ref
@pytest.mark.parametrize("mb,ab", [(5, 1.0), (1, 0.45)])
def test_specular_restatement_equals_reference(mb, ab):
    scene = drt.specular_box(36, 20, exponent_ball=50.0, exponent_wall=3.0, gloss=(0.9, 0.5, 0.3))
    a_img, a_grad, a_g = ref_render(scene, drt.make_opts(5, mb, ab, seed=4), grad_image_of=4)
    b_img, b_grad, b_g = restate_render(scene, drt.make_opts(5, mb, ab, seed=4), grad_image_of=4)
    assert np.array_equal(a_img, b_img) and np.isfinite(a_img).all()
    assert rel_err(b_grad, a_grad).max() < 1e-13
    assert np.abs(a_g - b_g).max() <= 1e-13 * np.abs(a_g).max()


@needs_ref
def test_specular_non_integer_exponent_nans_are_the_references():
    """Upstream clamps nothing: a reflected direction may leave below the surface, the
    next hit then sees the lobe from behind, and pow(negative, non-integer) is NaN
    (bxdf.hpp:102).  The restatement reproduces exactly those NaN pixels."""
    scene = drt.specular_box(36, 20, exponent_ball=50.0, exponent_wall=1.5)
    a_img, _ = ref_render(scene, drt.make_opts(5, 5, 1.0, seed=4))
    b_img, _ = restate_render(scene, drt.make_opts(5, 5, 1.0, seed=4))
    assert np.isnan(a_img).any() and np.array_equal(a_img, b_img, equal_nan=True)


def test_gradient_image_golden_and_pixel_sum():
    z, W, H, spp, mb, ab, seed, _ = load_case("cbox_48x32_8spp_b8_p1_gimg_red")
    img, grad, gimg = restate_render(drt.cornell_box(W, H), drt.make_opts(spp, mb, ab, seed_scale=1.0 / spp),
                                     grad_image_of=0)
    assert np.array_equal(img, z["img"]) and np.array_equal(gimg, z["gimg"])
    assert rel_err(gimg.sum((0, 1)), grad[0]).max() < 1e-13
    # requesting the gradient image does not change the image; gradients only re-associate
    img2, grad2 = restate_render(drt.cornell_box(W, H), drt.make_opts(spp, mb, ab, seed_scale=1.0 / spp))
    assert np.array_equal(img, img2) and rel_err(grad, grad2).max() < 1e-13


def test_specular_gradient_is_an_exact_derivative():
    """The lobe does not depend on the tint, so the fixed-stream estimator stays
    polynomial in it: central differences are exact to O(h^2)."""
    W, H, spp, mb, ab, h = 28, 20, 4, 4, 1.0, 1e-4
    _, grad = restate_render(drt.specular_box(W, H), drt.make_opts(spp, mb, ab))
    for c in range(3):
        tot = []
        for sgn in (+1, -1):
            g = [0.8, 0.7, 0.6]
            g[c] += sgn * h
            img, _ = restate_render(drt.specular_box(W, H, gloss=g), drt.make_opts(spp, mb, ab))
            tot.append(img[..., c].sum() * spp)
        fd = (tot[0] - tot[1]) / (2 * h)
        assert abs(fd - grad[4, c]) <= 1e-6 * max(1.0, abs(grad[4, c]))
