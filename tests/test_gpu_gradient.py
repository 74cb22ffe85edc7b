"""DenseCRF parameter learning: rss_crf_gradient against DenseCRF::gradient of the UNMODIFIED reference
(densecrf.cpp:238-297 + objective.cpp's LogLikelihood, compiled in place into oracle/_ref): objective value and the
gradient w.r.t. the Potts weight of every pairwise term, for every NormalizationType; plus a finite-difference check."""
import numpy as np
import pytest

from conftest import CONFIG, FOREST

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import rovinasemanticsegmentation_b200 as rss
    c = rss.Context(CONFIG, FOREST, 0)
    yield c
    c.close()


def _problem(N=6000, M=5, seed=0):
    rng = np.random.default_rng(seed)
    f3 = np.stack([rng.uniform(0, 30, N), rng.uniform(0, 30, N), rng.uniform(0, 8, N)], 1).astype(np.float32)
    f2 = np.ascontiguousarray(f3[:, :2] * np.float32(0.5))
    U = rng.random((N, M), dtype=np.float32) * 2
    gt = rng.integers(-1, M, N).astype(np.int32)  # -1 = unlabelled point (skipped by the objective)
    return U, f3, f2, gt


@pytest.mark.parametrize("norm_type", [3, 2, 1, 0])  # SYMMETRIC, AFTER, BEFORE, NO_NORMALIZATION (pairwise.h)
def test_gradient_matches_reference(ctx, orc, norm_type):
    U, f3, f2, gt = _problem()
    r0, g0 = orc.ref_crf_gradient(U, [(f3, 3.0), (f2, 1.5)], 3, gt, 0.0, norm_type)
    crf = ctx.crf(U.shape[0], [U.shape[1]])
    crf.set_unary(U, 0)
    crf.add_pairwise(f3, 3.0, norm_type)
    crf.add_pairwise(f2, 1.5, norm_type)
    r, g = crf.gradient(3, gt)
    # inference still works after a gradient call and equals the plain call
    Q = crf.inference(3)
    crf.close()
    assert abs(r - r0) <= 2e-6 * max(1.0, abs(r0)), (r, r0)
    assert np.allclose(g, g0, rtol=2e-3, atol=2e-6), (g, g0)
    if norm_type != 0:  # (without normalisation the marginals saturate and last-bit differences are amplified)
        assert np.abs(Q - orc.crf_inference(U, [(f3, 3.0), (f2, 1.5)], 3, norm_type)).max() <= 1e-4


def test_gradient_is_the_derivative_of_the_objective(ctx):
    """Finite differences of the library's own objective in the Potts weight of the first term, robust > 0, a two-layer CRF."""
    U, f3, f2, gt = _problem(N=4000, seed=3)
    rng = np.random.default_rng(9)
    U2 = rng.random((U.shape[0], 4), dtype=np.float32)

    def objective(w0):
        crf = ctx.crf(U.shape[0], [U.shape[1], 4])
        crf.set_unary(U, 0)
        crf.set_unary(U2, 1)
        crf.add_pairwise(f3, w0, 3)
        crf.add_pairwise(f2, 1.5, 3)
        out = crf.gradient(4, gt, layer=0, robust=0.05)
        crf.close()
        return out

    r, g = objective(3.0)
    eps = 0.05
    fd = (objective(3.0 + eps)[0] - objective(3.0 - eps)[0]) / (2 * eps)
    assert abs(fd - g[0]) <= 0.02 * abs(g[0]) + 1e-6, (fd, g)
