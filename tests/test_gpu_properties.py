"""Size-independent properties of the estimator, checked on the GPU at
BASELINE.json's full config-2 size (1024x1024, 256 spp, 8 bounces) where the
CPU oracle would need tens of core-minutes, plus the 3-sigma agreement of
independent seeds."""
import numpy as np
import pytest

from oracle_lib import rel_err, restate_render

pytestmark = pytest.mark.gpu

W = H = 1024
SPP, MB, AB = 256, 8, 1.0


@pytest.fixture(scope="module")
def full(drt, ctx):
    ctx.upload(drt.cornell_box(W, H))
    img, grad, st = ctx.render(drt.make_opts(SPP, MB, AB), stats=True)
    return img, grad, st


def test_full_size_counts_and_sanity(full):
    img, grad, st = full
    assert st.paths == W * H * SPP and st.truncated_paths == 0
    assert 7.2 < st.segments / st.paths < 7.45               # 7.31 segments/path at B=8
    assert 0.15 < st.lit_paths / st.paths < 0.17             # 15.9 % of paths carry radiance
    assert np.isfinite(img).all() and np.isfinite(grad).all() and (img >= 0).all()
    assert np.abs(img.reshape(-1, 3).mean(0) - np.array([0.0500, 0.0475, 0.0440])).max() < 5e-4


def test_full_size_euler_identity(full):
    """Radiance is linear in the emission: emission.grad . emission = sum of all
    path radiances = spp * sum(img), channel by channel."""
    img, grad, _ = full
    assert rel_err(grad[3] * 1.0, img.reshape(-1, 3).sum(0) * SPP).max() <= 1e-10


def test_full_size_linearity_in_emission(drt, ctx, full):
    img, grad, _ = full
    ctx.upload(drt.cornell_box(W, H, emission=(2.0, 0.5, 4.0)))
    img2, grad2 = ctx.render(drt.make_opts(SPP, MB, AB))
    assert rel_err(img2, img * np.array([2.0, 0.5, 4.0])).max() <= 1e-12
    assert rel_err(grad2[3], grad[3]).max() <= 1e-12         # d/dE does not depend on E


def test_full_size_gradient_is_the_exact_derivative(drt, ctx, full):
    """Fixed stream => polynomial in the albedo => central differences exact to O(h^2)."""
    _, grad, _ = full
    h = 1e-4
    for k, c, name in [(0, 1, "red"), (2, 2, "white")]:     # red[1] is an albedo that is exactly 0
        tot = []
        for sgn in (+1, -1):
            kw = dict(red=[0.5, 0, 0], white=[0.5, 0.5, 0.5])
            kw[name][c] += sgn * h
            ctx.upload(drt.cornell_box(W, H, **kw))
            img, _ = ctx.render(drt.make_opts(SPP, MB, AB, flags=drt.FLAG_IMAGE))
            tot.append(img[..., c].sum() * SPP)
        fd = (tot[0] - tot[1]) / (2 * h)
        assert abs(fd - grad[k, c]) <= 1e-6 * abs(grad[k, c])


def test_full_size_f32_agrees_with_f64(drt, ctx, full):
    img, grad, _ = full
    ctx.upload(drt.cornell_box(W, H))
    img32, grad32 = ctx.render(drt.make_opts(SPP, MB, AB, precision=drt.F32))
    assert rel_err(grad32, grad).max() <= 1e-3
    bad = (rel_err(img32, img) > 1e-4).any(axis=-1).mean()
    assert bad <= 2e-2, f"{bad:.2e} of the pixels are outside 1e-4"
    assert rel_err(img32.reshape(-1, 3).mean(0), img.reshape(-1, 3).mean(0)).max() <= 1e-5


def test_independent_seeds_agree_within_three_sigma(drt, ctx):
    """Different `seed` => independent streams.  sigma per pixel / per gradient
    scalar is estimated from 8 further independent renders."""
    w = h = 128
    spp = 64
    scene = drt.cornell_box(w, h)
    ctx.upload(scene)
    runs = [ctx.render(drt.make_opts(spp, MB, AB, seed=s)) for s in range(1, 11)]
    imgs = np.stack([r[0] for r in runs]); grads = np.stack([r[1] for r in runs])
    a, b = imgs[0], imgs[1]
    var = imgs[2:].var(axis=0, ddof=1)
    ok = np.abs(a - b) <= 3.0 * np.sqrt(2.0 * var) + 1e-12
    assert ok.mean() >= 0.97
    gs = grads[2:].std(axis=0, ddof=1)
    assert (np.abs(grads[0] - grads[1]) <= 3.0 * np.sqrt(2.0) * gs * 1.5).all()
    # and the GPU on one seed vs the CPU oracle on another
    o_img, o_grad = restate_render(scene, drt.make_opts(spp, MB, AB, seed=77), threads=8)
    assert (np.abs(o_grad - grads[0]) <= 3.0 * np.sqrt(2.0) * gs * 1.5).all()
    m = imgs[2:].reshape(8, -1, 3).mean(1)
    assert (np.abs(o_img.reshape(-1, 3).mean(0) - m.mean(0)) <= 4.0 * m.std(0, ddof=1) + 1e-4).all()


def test_config5_segment_rate_is_flat_across_bounce_counts(drt, ctx):
    """BASELINE.json configs[4] (max-bounce sweep): a path of B bounces costs about B segments -- ray segments per
    second stay within 25 % of their best for B = 1 .. 8 (deeper records lose a little more: the sweep table under
    profiles/ goes to B = 16 at 2048^2, 128 spp on 1 / 2 / 8 GPUs).  Best of three timed renders each."""
    ctx.upload(drt.cornell_box(1024, 1024))
    rate = {}
    for B in (1, 2, 4, 8):
        o = drt.make_opts(64, B, 1.0)
        best = None
        for _ in range(3):
            _, _, st = ctx.render(o, stats=True)
            best = st.kernel_ms if best is None else min(best, st.kernel_ms)
        rate[B] = st.segments / best / 1e3                      # Msegments/s
    assert min(rate.values()) >= 0.75 * max(rate.values()), rate
    assert rate[8] > 35000, rate                                 # ~45 Gsegments/s on a B200
