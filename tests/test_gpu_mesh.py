"""Triangle meshes + GPU-built BVH (config 4): parity with the oracle on small
tessellations, BVH-vs-linear-scan equality on the GPU at sizes the oracle cannot
reach, per-triangle albedo gradients."""
import numpy as np
import pytest

import oracle_lib
from oracle_lib import rel_err, restate_render
from test_oracle import mixed_mesh_scene

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mb,ab,spp", [(4, 1.0, 4), (1, 0.4, 6), (6, 1.0, 40)])
def test_small_mesh_matches_oracle(drt, ctx, mb, ab, spp):
    for scene in (drt.tessellated_room(2, 4, width=32, height=24), mixed_mesh_scene()):
        ctx.upload(scene)
        img, grad, st = ctx.render(drt.make_opts(spp, mb, ab, seed=2), stats=True)
        ref_img, ref_grad, ref_st = restate_render(scene, drt.make_opts(spp, mb, ab, seed=2), want_stats=True)
        assert st.segments == ref_st.segments and st.lit_paths == ref_st.lit_paths
        assert rel_err(img, ref_img).max() <= 1e-9
        assert np.abs(grad - ref_grad).max() <= 1e-9 * np.abs(ref_grad).max()
        assert st.bvh_nodes > 0 and st.tri_tests > 0
        assert st.tri_tests < 0.6 * st.segments * scene.mesh.n_triangles      # the BVH actually culls


def test_mesh_golden_vector_from_reference_triangle_shape(drt, ctx):
    z = np.load(oracle_lib.ROOT / "tests" / "golden" / "mesh_room_98tri_32x24_4spp_b4.npz")
    ctx.upload(drt.tessellated_room(2, 4, width=32, height=24))
    img, grad = ctx.render(drt.make_opts(4, 4, 1.0))
    assert rel_err(img, z["img"]).max() <= 1e-9
    assert np.abs(grad - z["grad"]).max() <= 1e-9 * np.abs(z["grad"]).max()


def test_bvh_equals_linear_scan_on_a_medium_mesh(drt, ctx):
    scene = drt.tessellated_room(24, 48, width=96, height=64)          # 16 k triangles
    assert scene.mesh.n_triangles > 15000
    ctx.upload(scene)
    a_img, a_grad, a = ctx.render(drt.make_opts(4, 4, 1.0), stats=True)
    b_img, b_grad, b = ctx.render(drt.make_opts(4, 4, 1.0, flags=drt.FLAG_IMAGE | drt.FLAG_GRAD | drt.FLAG_NO_BVH), stats=True)
    assert np.array_equal(a_img, b_img)
    assert a.segments == b.segments and a.lit_paths == b.lit_paths
    assert np.abs(a_grad - b_grad).max() <= 1e-12 * np.abs(b_grad).max()      # atomics: order only
    assert a.tri_tests < b.tri_tests / 100
    nz = np.abs(a_grad[scene.mesh.param_base:]).sum(1) > 0
    assert nz.mean() > 0.05                                           # gradients are spread over the triangles


def test_deterministic_gradient_mode_is_bit_reproducible_and_order_independent(drt, ctx):
    """DRTB_FLAG_DETERMINISTIC ("offer a deterministic two-stage reduction for tests"): per-triangle
    gradients summed in 64-bit fixed point.  The BVH and the linear scan finish their rays in different orders and
    the traversal refills its lanes dynamically, yet the gradients must come out bit-identical -- and within the
    fixed-point resolution of the floating-point atomics' result."""
    scene = drt.tessellated_room(24, 48, width=96, height=64)          # 16 k triangles
    ctx.upload(scene)
    det = drt.FLAG_IMAGE | drt.FLAG_GRAD | drt.FLAG_DETERMINISTIC
    a_img, a_grad = ctx.render(drt.make_opts(4, 4, 1.0, flags=det))
    b_img, b_grad = ctx.render(drt.make_opts(4, 4, 1.0, flags=det | drt.FLAG_NO_BVH))
    c_img, c_grad = ctx.render(drt.make_opts(4, 4, 1.0, flags=det))
    assert np.array_equal(a_img, b_img) and np.array_equal(a_grad, b_grad) and np.array_equal(a_grad, c_grad)
    f_img, f_grad = ctx.render(drt.make_opts(4, 4, 1.0))
    assert np.array_equal(a_img, f_img)
    assert np.abs(a_grad - f_grad).max() <= 1e-7                       # <= 2^-33 per contribution, a few hundred of them
    assert np.abs(a_grad - f_grad).max() <= 1e-8 * np.abs(f_grad).max()
    # the analytic-only scene (12 gradient scalars, per-chunk rows) is deterministic anyway: the flag changes nothing
    ctx.upload(drt.cornell_box(48, 32))
    g0 = ctx.render(drt.make_opts(8, 4, 1.0))[1]
    g1 = ctx.render(drt.make_opts(8, 4, 1.0, flags=det))[1]
    assert np.array_equal(g0, g1)


def test_mesh_explicit_rays_and_jacobian(drt, ctx):
    import ctypes as C
    scene = mixed_mesh_scene()
    ctx.upload(scene)
    rng = np.random.default_rng(5)
    n = 48
    orig = rng.uniform(-1, 1, size=(n, 3)) + np.array([0, 0, 1.5])
    dirs = rng.normal(size=(n, 3)); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    keys = rng.integers(0, 2**62, size=n, dtype=np.uint64)
    rad, jac = ctx.trace_rays(drt.make_opts(1, 3, 0.3), orig, dirs, keys)
    lib = oracle_lib.load_restate()
    sc, mesh = scene.flatten(), scene.flatten_mesh()
    o = drt.make_opts(1, 3, 0.3)
    ref_rad = np.zeros((n, 3)); ref_jac = np.zeros((n, scene.n_params, 3))
    dp = C.POINTER(C.c_double)
    rc = lib.drt_oracle_trace_rays_mesh(C.byref(sc), C.byref(mesh), C.byref(o), n, orig.ctypes.data_as(dp),
                                        dirs.ctypes.data_as(dp), keys.ctypes.data_as(C.POINTER(C.c_uint64)),
                                        ref_rad.ctypes.data_as(dp), ref_jac.ctypes.data_as(dp))
    assert rc == 0 and ref_rad.max() > 0
    assert rel_err(rad, ref_rad).max() <= 1e-9 and rel_err(jac, ref_jac).max() <= 1e-9


def test_mesh_f32_within_outlier_budget(drt, ctx):
    scene = drt.tessellated_room(8, 16, width=96, height=64)
    ctx.upload(scene)
    img64, grad64 = ctx.render(drt.make_opts(32, 4, 1.0))
    img32, grad32 = ctx.render(drt.make_opts(32, 4, 1.0, precision=drt.F32))
    bad = (rel_err(img32, img64) > 1e-4).any(axis=-1).mean()
    assert bad <= 2e-2
    assert rel_err(grad32.sum(0), grad64.sum(0)).max() <= 1e-3


def test_mesh_validation_and_detach(drt, ctx):
    from drt_b200 import abi
    scene = drt.tessellated_room(2, 4, width=16, height=16)
    scene.mesh.indices[0, 0] = 10**6
    with pytest.raises(drt.DrtbError) as e:
        ctx.upload(scene)
    assert e.value.code == abi.ERR_INVALID
    ctx.upload(drt.cornell_box(16, 16))                 # a new scene detaches the mesh
    img, grad = ctx.render(drt.make_opts(4, 2, 0.5))
    ref_img, _ = restate_render(drt.cornell_box(16, 16), drt.make_opts(4, 2, 0.5))
    assert rel_err(img, ref_img).max() <= 1e-9


def test_million_triangle_build_and_render(drt, ctx):
    """Config 4 scale: 1 M triangles, GPU LBVH build, BVH == linear scan on a ray subset."""
    scene = drt.tessellated_room(204, 362, width=128, height=128)
    n = scene.mesh.n_triangles
    assert 1_000_000 <= n <= 1_100_000
    ctx.upload(scene)
    img, grad, st = ctx.render(drt.make_opts(8, 4, 1.0), stats=True)
    assert np.isfinite(img).all() and np.isfinite(grad).all() and img.mean() > 0
    assert st.truncated_paths == 0 and 3.5 < st.segments / st.paths <= 4.0
    assert st.tri_tests / st.segments < 64                       # ~log N work per ray, not N
    rng = np.random.default_rng(11)
    m = 2048
    orig = rng.uniform(-1, 1, size=(m, 3)) + np.array([0, 0, 1.5])
    dirs = rng.normal(size=(m, 3)); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    keys = rng.integers(0, 2**62, size=m, dtype=np.uint64)
    a, _ = ctx.trace_rays(drt.make_opts(1, 2, 1.0), orig, dirs, keys, jac=False)
    b, _ = ctx.trace_rays(drt.make_opts(1, 2, 1.0, flags=drt.FLAG_IMAGE | drt.FLAG_NO_BVH), orig, dirs, keys, jac=False)
    assert np.array_equal(a, b) and a.max() > 0


def test_bvh_equals_linear_scan_for_far_ray_origins_and_an_offset_mesh(drt, ctx):
    """The float slab test must stay conservative when the float rounding of the ray is NOT small against the mesh:
    a mesh far from the coordinate origin (|coordinates| ~ 100 extents) and rays that start ~500 extents away, aimed
    at triangle VERTICES -- points that lie exactly on the faces of their leaf boxes.  The leaf padding scales with
    the coordinate magnitude and the slab compare is widened relative to t (bvh.cuh, node8_step)."""
    scene = drt.tessellated_room(24, 48, width=16, height=16)          # 16 k triangles, extent 6
    scene.mesh.vertices += np.array([500.0, -300.0, 200.0])
    ctx.upload(scene)
    rng = np.random.default_rng(5)
    m = 1 << 16
    v = scene.mesh.vertices[rng.integers(0, scene.mesh.vertices.shape[0], size=m)]
    out = rng.normal(size=(m, 3)); out /= np.linalg.norm(out, axis=1, keepdims=True)
    orig = v + out * rng.uniform(1000.0, 3000.0, size=(m, 1))
    dirs = -out + rng.normal(size=(m, 3)) * 1e-9
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    keys = rng.integers(0, 2**62, size=m, dtype=np.uint64)
    a, _ = ctx.trace_rays(drt.make_opts(1, 2, 1.0), orig, dirs, keys, jac=False)
    b, _ = ctx.trace_rays(drt.make_opts(1, 2, 1.0, flags=drt.FLAG_IMAGE | drt.FLAG_NO_BVH), orig, dirs, keys, jac=False)
    assert np.array_equal(a, b) and np.isfinite(a).all()
