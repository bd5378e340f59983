"""The drop-in C++ headers on the GPU beyond the reference application: drt::RenderOptions::devices
(drtb_multi_render: every GPU of the box behind one call) and Triangle<T> shapes (gpu::flatten -> drtb_mesh -> GPU
BVH).  tests/cpp/test_gpu_dropin.cpp does the C++ side; here the triangle scene is rebuilt through the C ABI's
Python mirror and must give the same image bit for bit."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
GXX = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"


@pytest.fixture(scope="module")
def dropin_exe(tmp_path_factory):
    exe = tmp_path_factory.mktemp("cpp") / "test_gpu_dropin"
    lib = ROOT / "differentiable-renderer_b200" / "lib"
    cmd = [GXX, "-std=c++17", "-O1", "-Wall", "-Wextra", "-I", str(ROOT / "include"),
           str(ROOT / "tests" / "cpp" / "test_gpu_dropin.cpp"), "-o", str(exe), "-L", str(lib), "-ldrtb", f"-Wl,-rpath,{lib}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0 and r.stderr.strip() == "", r.stderr
    return exe


def triangle_room(drt):
    """The scene of test_gpu_dropin.cpp part 2 through the Python mirror of the C ABI."""
    P = lambda v, n: drt.Param(np.asarray(v, dtype=np.float64), n)
    white, emission = P((0.5, 0.5, 0.5), "white"), P((1, 1, 1), "emission")
    floor_col, wall_col = P((0.7, 0.6, 0.5), "floor"), P((0.3, 0.5, 0.8), "wall")
    sc = drt.SceneDesc()
    sc.push_back(drt.Sphere((0., 0., 3.), 1., drt.DiffuseBxDF(white)))
    sc.push_back(drt.Sphere((0., 3., 3.), 1., None, drt.AreaEmitter(emission)))
    quads = [((-3, -3, 0), (3, -3, 0), (3, -3, 6), (-3, -3, 6), floor_col), ((-3, -3, 6), (3, -3, 6), (3, 3, 6), (-3, 3, 6), wall_col),
             ((-3, -3, 0), (-3, -3, 6), (-3, 3, 6), (-3, 3, 0), wall_col), ((3, -3, 0), (3, 3, 0), (3, 3, 6), (3, -3, 6), floor_col)]
    verts, idx, cols = [], [], []
    for a, b, c, d, col in quads:
        for tri in ((a, c, b), (a, d, c)):
            base = len(verts)
            verts += [tri[0], tri[1], tri[2]]
            idx.append((base, base + 1, base + 2))
            cols.append(col)
    sc.camera = drt.Camera(64, 48).look_at((0, 0, 0), (0, 0, 1))
    return sc, np.array(verts, dtype=np.float64), np.array(idx, dtype=np.int32), cols, (floor_col, wall_col)


def test_cpp_multi_device_and_triangles(drt, dropin_exe, tmp_path):
    import ctypes as C
    from drt_b200 import abi
    lib = drt.load_library()
    n = max(1, min(8, lib.drtb_device_count()))
    out = tmp_path / "tri.bin"
    r = subprocess.run([str(dropin_exe), str(n), str(out)], capture_output=True, text=True)
    assert r.returncode == 0 and "all GPU drop-in checks passed" in r.stdout, r.stdout + r.stderr
    raw = np.fromfile(out, dtype=np.float64)
    cpp_img, cpp_grad = raw[:64 * 48 * 3].reshape(48, 64, 3), raw[64 * 48 * 3:].reshape(2, 3)
    # the same scene through the C ABI directly (raw drtb_mesh with per-triangle parameter indices)
    sc, verts, idx, cols, (floor_col, wall_col) = triangle_room(drt)
    flat = sc.flatten()
    floor_i, wall_i = sc.n_params, sc.n_params + 1                      # two more parameters behind the objects'
    params = np.concatenate([sc.param_values(), [floor_col.value, wall_col.value]]).astype(np.float64)
    flat.params, flat.n_params = params.ctypes.data_as(C.POINTER(C.c_double)), params.shape[0]
    color = np.array([floor_i if c is floor_col else wall_i for c in cols], dtype=np.int32)
    mesh = abi.Mesh()
    mesh.vertices, mesh.n_vertices = verts.ctypes.data_as(C.POINTER(C.c_double)), verts.shape[0]
    mesh.indices, mesh.n_triangles = idx.ctypes.data_as(C.POINTER(C.c_int32)), idx.shape[0]
    mesh.color, mesh.emission = color.ctypes.data_as(C.POINTER(C.c_int32)), None
    h = C.c_void_p()
    assert lib.drtb_create(0, C.byref(h)) == 0
    try:
        assert lib.drtb_scene_upload(h, C.byref(flat)) == 0 and lib.drtb_mesh_upload(h, C.byref(mesh)) == 0
        img = np.empty((48, 64, 3)); grad = np.empty((params.shape[0], 3))
        o = drt.make_opts(16, 4, 1.0)
        dp = C.POINTER(C.c_double)
        assert lib.drtb_render(h, C.byref(o), None, img.ctypes.data_as(dp), grad.ctypes.data_as(dp), None) == 0
    finally:
        lib.drtb_destroy(h)
    assert np.array_equal(img, cpp_img)
    assert np.abs(grad[[floor_i, wall_i]] - cpp_grad).max() <= 1e-12 * np.abs(cpp_grad).max()


def test_multi_handle_through_the_c_abi_matches_one_device(drt, ctx):
    """drtb_multi_* with every device of the box, raw C ABI: Cornell box with a per-pixel seed image, and a mesh
    scene (compact bands copied to their rows)."""
    import ctypes as C
    lib = drt.load_library()
    n = max(1, min(8, lib.drtb_device_count()))
    devs = (C.c_int * n)(*range(n))
    dp = C.POINTER(C.c_double)
    m = C.c_void_p()
    assert lib.drtb_multi_create(devs, n, C.byref(m)) == 0, lib.drtb_multi_last_error(None)
    try:
        assert lib.drtb_multi_device_count(m) == n
        rng = np.random.default_rng(3)
        for scene in (drt.cornell_box(80, 52), drt.tessellated_room(3, 6, width=40, height=36)):
            H, W = scene.camera.height, scene.camera.width
            seed_img = rng.uniform(-1, 1, size=(H, W, 3))
            ctx.upload(scene)
            o = drt.make_opts(12, 3, 0.4, seed_scale=1.0 / 12)
            ref_img, ref_grad, ref_st = ctx.render(o, seed_img=seed_img, stats=True)
            flat = scene.flatten()
            assert lib.drtb_multi_scene_upload(m, C.byref(flat)) == 0, lib.drtb_multi_last_error(m)
            mesh = scene.flatten_mesh()
            if mesh is not None:
                assert lib.drtb_multi_mesh_upload(m, C.byref(mesh)) == 0, lib.drtb_multi_last_error(m)
            img = np.full((H, W, 3), np.nan); grad = np.empty_like(ref_grad)
            st = drt.abi.Stats()
            o2 = drt.make_opts(12, 3, 0.4, seed_scale=1.0 / 12)
            rc = lib.drtb_multi_render(m, C.byref(o2), seed_img.ctypes.data_as(dp), img.ctypes.data_as(dp),
                                       grad.ctypes.data_as(dp), C.byref(st))
            assert rc == 0, lib.drtb_multi_last_error(m)
            # Russian roulette: the regenerating kernel adds a pixel's lit samples in the order they finish, which depends
            # on which pixels share a chunk -- and so on the sharding: equal up to the last bit, not bit for bit
            bad_rows = np.nonzero((np.abs(img - ref_img) > 4e-16 * np.abs(ref_img)).any(axis=(1, 2)))[0]
            assert bad_rows.size == 0, (f"{n} devices, mesh={mesh is not None}: rows {bad_rows.tolist()} differ, "
                                        f"max |diff| {np.nanmax(np.abs(img - ref_img))}, NaNs {int(np.isnan(img).sum())}")
            assert np.abs(grad - ref_grad).max() <= 1e-11 * np.abs(ref_grad).max()
            assert st.paths == ref_st.paths and st.segments == ref_st.segments and st.lit_paths == ref_st.lit_paths
        bad = drt.make_opts(0, 3, 0.4)
        assert lib.drtb_multi_render(m, C.byref(bad), None, img.ctypes.data_as(dp), grad.ctypes.data_as(dp), None) != 0
        assert b"spp" in lib.drtb_multi_last_error(m)
    finally:
        lib.drtb_multi_destroy(m)
