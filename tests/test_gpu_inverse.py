"""Config 3: the inverse-rendering loop (image pass -> MSE seed -> decorrelated
adjoint pass -> drtb_set_params) recovers the wall albedos."""
import importlib.util
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_gradient_descent_recovers_the_wall_albedos():
    spec = importlib.util.spec_from_file_location("inverse_render", ROOT / "examples" / "inverse_render.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    theta, err, hist, ips = mod.fit(256, 256, 64, 4, iters=100)          # the survey's size for config 3 (256^2 .. 512^2)
    # the loss floor is the Monte Carlo variance of two independent 64-spp-class images
    assert hist[-1] < 0.7 * hist[0], (hist[0], hist[-1])
    assert err < 0.03, (theta, err)
    # red wall reflects red only, green wall green only: the zero channels must be found too
    assert theta[0][1] < 0.02 and theta[0][2] < 0.02 and theta[1][0] < 0.02 and theta[1][2] < 0.02
