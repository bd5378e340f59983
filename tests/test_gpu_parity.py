"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle
on the same scene and the same pre-generated sample stream.

Tolerances are the north-star's: image 1e-4 relative per pixel, gradients 1e-3
relative per parameter (the double instantiation lands ~1e-12 / 1e-13)."""
import numpy as np
import pytest

import oracle_lib
from oracle_lib import rel_err, restate_render

pytestmark = pytest.mark.gpu

IMG_TOL = 1e-4
GRAD_TOL = 1e-3

SETTINGS = [(8, 1.0), (1, 0.5), (3, 0.3), (0, 0.25), (16, 1.0)]


@pytest.mark.parametrize("mb,absorb", SETTINGS)
@pytest.mark.parametrize("spp", [8, 5, 32, 40])
def test_f64_matches_oracle(drt, ctx, mb, absorb, spp):
    scene = drt.cornell_box(48, 32)
    ctx.upload(scene)
    img, grad, st = ctx.render(drt.make_opts(spp, mb, absorb), stats=True)
    ref_img, ref_grad, ref_st = restate_render(scene, drt.make_opts(spp, mb, absorb), want_stats=True)
    assert st.truncated_paths == 0
    assert st.paths == ref_st.paths and st.segments == ref_st.segments and st.lit_paths == ref_st.lit_paths
    assert rel_err(img, ref_img).max() <= IMG_TOL
    assert rel_err(grad, ref_grad).max() <= GRAD_TOL
    # the double instantiation is far inside the tolerance
    assert rel_err(img, ref_img).max() <= 1e-9
    assert rel_err(grad, ref_grad).max() <= 1e-9
    # exact zeros stay exact zeros
    assert np.array_equal(img == 0.0, ref_img == 0.0)
