"""CPU tests: pin the oracle (oracle/oracle.c) against
  * cv2 4.13.0 outputs frozen in tests/golden/cv_golden.npz,
  * outputs of the UNMODIFIED reference (libforest, permutohedral.cpp) frozen in ref_golden.npz,
  * and, where oracle/_ref/libref_oracle.so is present, the reference itself run live.
"""
import numpy as np
import pytest

from conftest import FOREST


def test_lab_matches_cv2(orc, cv_golden):
    assert np.array_equal(orc.bgr2lab(cv_golden["lab_src"]), cv_golden["lab_dst"])


def test_border_reflect_matches_cv2(orc, cv_golden):
    for b in (5, 20):
        assert np.array_equal(orc.border_reflect(cv_golden["border_src"], b), cv_golden["border_dst_%d" % b])


def test_resize_u8_matches_cv2(orc, cv_golden):
    big = cv_golden["resize_src"]
    for S in cv_golden["resize_sizes"]:
        win = np.ascontiguousarray(big[2:2 + S, 3:3 + S])
        assert np.array_equal(orc.resize_u8c3(win, 11), cv_golden["resize_dst_%d" % S]), S
        assert np.array_equal(orc.resize_u8c3(win, 5), cv_golden["resize5_dst_%d" % S]), S


def test_resize_f32_matches_cv2(orc, cv_golden):
    for C in (8, 9):
        src = cv_golden["up_src_%d" % C]
        assert np.array_equal(orc.resize_f32(src, 32, 24), cv_golden["up_dst_%d" % C])
        assert np.array_equal(orc.resize_f32(src, 48, 36), cv_golden["up3_dst_%d" % C])


def _rows(ref_golden):
    return np.concatenate([ref_golden["rf_rows_color"].astype(np.float32), ref_golden["rf_rows_tail"]], axis=1)


def test_forest_matches_reference_golden(orc, ref_golden):
    f = orc.Forest(FOREST)
    assert (f.T, f.L, f.classes) == (4, 2, [8, 9])
    leaf, post = f.predict(_rows(ref_golden))
    assert np.array_equal(leaf, ref_golden["rf_leaf"])
    assert np.array_equal(post, ref_golden["rf_post"])  # bit-exact: same add order t=0..3


@pytest.mark.parametrize("name", ["d6", "d5", "d3", "d2"])
def test_lattice_matches_reference_golden(orc, ref_golden, name):
    feats = ref_golden["lat_%s_feats" % name]
    lat = orc.Lattice(feats)
    assert lat.V == int(ref_golden["lat_%s_V" % name])
    off, bary = lat.get()
    assert np.array_equal(off, ref_golden["lat_%s_off" % name])  # same first-seen vertex numbering
    assert np.array_equal(bary, ref_golden["lat_%s_bary" % name])
    assert np.array_equal(lat.compute(ref_golden["lat_%s_in9" % name]), ref_golden["lat_%s_out9" % name])
    ones = np.ones((feats.shape[0], 1), np.float32)
    assert np.array_equal(lat.compute(ones), ref_golden["lat_%s_out1" % name])


def test_live_reference_forest_and_lattice(orc):
    """Where the compiled reference is available, compare on a fresh full-size frame too."""
    if not orc.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    from rovinasemanticsegmentation_b200 import synth
    Kinv, R, t = synth.calibration()
    rgb, depth = synth.frame(77)
    feats, xs, ys = orc.extract(orc.default_config(), 4, rgb, depth, Kinv, R, t, 0.5, 15.0)
    mine, ref = orc.Forest(FOREST), orc.RefForest(FOREST)
    l0, p0 = mine.predict(feats)
    l1, p1 = ref.predict(feats, mine.sumC)
    assert np.array_equal(l0, l1) and np.array_equal(p0, p1)
    xyz, col = synth.local_map(seed=5, n_points=20001)
    f6 = orc.features_xyzrgb(xyz, col, 0.5, 4.0)
    a, b = orc.Lattice(f6), orc.RefLattice(f6)
    assert a.V == b.V
    x = np.random.default_rng(0).random((20001, 8), dtype=np.float32)
    assert np.array_equal(a.compute(x), b.compute(x))


def test_three_layer_forest_matches_live_reference(orc):
    """The small 3-layer forest (4 + 7 + 5 classes, tests/golden/make_forest_3layer.py): oracle == compiled reference."""
    import os
    from conftest import GOLDEN
    path = os.path.join(GOLDEN, "forest_3layer.dat")
    mine = orc.Forest(path)
    assert (mine.T, mine.L, mine.classes) == (3, 3, [4, 7, 5])
    if not orc.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    from rovinasemanticsegmentation_b200 import synth
    Kinv, R, t = synth.calibration()
    rgb, depth = synth.frame(78)
    feats, xs, ys = orc.extract(orc.default_config(), 6, rgb, depth, Kinv, R, t, 0.5, 15.0)
    l0, p0 = mine.predict(feats)
    l1, p1 = orc.RefForest(path).predict(feats, mine.sumC)
    assert np.array_equal(l0, l1) and np.array_equal(p0, p1)


@pytest.mark.parametrize("d", [1, 2, 3, 4, 5, 6, 7])
def test_lattice_every_dimension_and_padding_vs_live_reference(orc, d):
    """Oracle lattice == compiled reference for every feature dimension and every N mod 4 (the reference pads its last
    SSE block with zero-feature points whose vertices exist for the blur, permutohedral.cpp:192-198)."""
    if not orc.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    rng = np.random.default_rng(d)
    for N in (1001, 1002, 1003, 1004):
        t = np.linspace(0, 5, N)
        f = (np.stack([np.sin((k + 1) * t + k) * (1.5 + k) for k in range(d)], axis=1) + rng.normal(0, 0.03, (N, d))).astype(np.float32)
        a, b = orc.Lattice(f), orc.RefLattice(f)
        assert a.V == b.V, (d, N)
        oa, ba = a.get()
        ob, bb = b.get()
        assert np.array_equal(oa, ob) and np.array_equal(ba, bb)
        x = rng.random((N, 5), dtype=np.float32)
        assert np.array_equal(a.compute(x), b.compute(x))
        ones = np.ones((N, 1), np.float32)
        assert np.array_equal(a.compute(ones), b.compute(ones))  # the scalar path used for the normalisation


# ---------------------------------------------------------------------------------------------------------------------
# DenseCRF glue (soft-max, normalisation, Potts, unary): the oracle's restatement vs the UNMODIFIED reference sources
# densecrf.cpp / pairwise.cpp / labelcompatibility.cpp / unary.cpp (oracle/Makefile compiles them against oracle/shim's
# Eigen stand-in).  Tolerance 1e-6: the stand-in's reductions and expf may differ from the oracle's in the last bit.
# ---------------------------------------------------------------------------------------------------------------------
GLUE_TOL = 1e-6


@pytest.fixture(scope="module")
def crf_golden():
    import os
    from conftest import GOLDEN
    return np.load(os.path.join(GOLDEN, "crf_golden.npz"))


def _golden_problem(g, name):
    kernels, k = [], 0
    while "%s_feat%d" % (name, k) in g:
        kernels.append((g["%s_feat%d" % (name, k)], float(g["%s_w%d" % (name, k)])))
        k += 1
    return g[name + "_unary"], kernels, int(g[name + "_iters"])


@pytest.mark.parametrize("name", ["node6d", "two_kernels"])
@pytest.mark.parametrize("norm_type", [0, 1, 2, 3])
def test_crf_glue_matches_reference_golden(orc, crf_golden, name, norm_type):
    """Frozen outputs of DenseCRF::inference + currentMap for every NormalizationType (pairwise.cpp:40-80)."""
    U, kernels, iters = _golden_problem(crf_golden, name)
    Q = orc.crf_inference(U, kernels, iters, norm_type)
    ref = crf_golden["%s_Q_norm%d" % (name, norm_type)]
    assert np.abs(Q - ref).max() <= GLUE_TOL
    assert (Q.argmax(1) == crf_golden["%s_map_norm%d" % (name, norm_type)]).mean() >= 0.999  # ties aside


def test_crf2d_glue_matches_reference_golden(orc, crf_golden):
    """DenseCRF2D::addPairwiseGaussian / addPairwiseBilateral feature builders + inference (densecrf.cpp:61-81)."""
    g = crf_golden
    W, H = 48, 32
    Q = orc.crf_inference(g["crf2d_unary"], [(orc.features_gaussian2d(W, H, 3, 3), 3.0),
                                             (orc.features_bilateral2d(W, H, 80, 80, 13, 13, 13, g["crf2d_rgb"]), 10.0)], 5)
    assert np.abs(Q - g["crf2d_Q"]).max() <= GLUE_TOL


def test_crf_glue_matches_live_reference(orc):
    """The same live, on the bench's own kernels at a size the reference finishes in seconds: a 160x120 keyframe's
    Gaussian 3-D + bilateral 5-D kernels, 10 iterations, both layers; and startInference / stepInference."""
    if not orc.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    from rovinasemanticsegmentation_b200 import synth
    W, H = 160, 120
    rgb, depth = synth.frame(12, W, H)
    Kinv, R, t = synth.calibration(W, H)
    post = orc.segment_frame(orc.default_config(), orc.Forest(FOREST), 2, rgb, depth, Kinv, R, t, 0.5, 15.0, 0.0)
    xyz = orc.cloud(depth, Kinv, R, t, 0.5, 15.0).reshape(-1, 3)
    xyz[np.isnan(xyz[:, 0])] = t
    f3 = (xyz * np.float32(1.0 / 0.05)).astype(np.float32)
    f5 = orc.features_bilateral2d(W, H, 80, 80, 13, 13, 13, rgb)
    N, off = W * H, 0
    for M in (8, 9):
        U = -post[off:off + N * M].reshape(N, M)
        Q0 = orc.crf_inference(U, [(f3, 3.0), (f5, 10.0)], 10)
        Q1, mp = orc.ref_crf_inference(U, [(f3, 3.0), (f5, 10.0)], 10, want_map=True)
        assert np.abs(Q0 - Q1).max() <= GLUE_TOL
        assert (Q0.argmax(1) == mp).mean() >= 0.9999
        off += N * M
    U = -post[:N * 8].reshape(N, 8)
    Qs = orc.ref_crf_step_inference(U, f5, 10.0, 3)  # startInference + 3 x stepInference == inference(3)
    assert np.abs(Qs - orc.crf_inference(U, [(f5, 10.0)], 3)).max() <= GLUE_TOL


def test_projector_known_answers(orc):
    """orc_project_zbuffer (this project's definition of the un-vendored fps_mapper projector, src/segmenter.cpp:581):
    hand-checkable cases - principal ray, nearest point wins, equal depth -> lower index, range and image clipping."""
    W, H = 64, 48
    K = np.array([[50, 0, 32], [0, 50, 24], [0, 0, 1]], np.float32)
    R = np.eye(3, dtype=np.float32)
    t = np.array([1.0, 2.0, 3.0], np.float32)
    pts = np.array([
        [1.0, 2.0, 5.0],    # 0: on the optical axis, z = 2 -> pixel (32, 24)
        [1.0, 2.0, 4.0],    # 1: same ray, z = 1: nearer, wins the pixel
        [1.0, 2.0, 4.0],    # 2: duplicate of 1: equal z, the lower index (1) keeps the pixel
        [1.4, 2.0, 5.0],    # 3: x = 0.4 at z = 2 -> u = 50 * 0.4 / 2 + 32 = 42
        [1.0, 2.0, 2.0],    # 4: behind the camera
        [1.0, 2.0, 3.2],    # 5: z = 0.2 < zmin
        [9.0, 2.0, 4.0],    # 6: far outside the image
        [1.0, 2.2, 13.0],   # 7: z = 10 = zmax (inclusive): v = 50 * 0.2 / 10 + 24 = 25
    ], np.float32)
    idx = orc.project_zbuffer(pts, K, R, t, W, H, 0.5, 10.0)
    assert idx[24, 32] == 1 and idx[24, 42] == 3 and idx[25, 32] == 7
    assert (idx >= 0).sum() == 3
    # a rotated camera: 90 degrees about y (camera z -> map x): the point ahead in map x lands on the principal point
    Ry = np.array([[0, 0, 1], [0, 1, 0], [-1, 0, 0]], np.float32)
    idx = orc.project_zbuffer(np.array([[4.0, 2.0, 3.0]], np.float32), K, Ry, t, W, H, 0.5, 10.0)
    assert idx[24, 32] == 0 and (idx >= 0).sum() == 1
