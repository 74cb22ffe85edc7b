"""GPU parity tests of the frame path (features -> forest -> upsample), through the C ABI, against the oracle.
Bar (BASELINE.json north_star): features, leaf indices and summed log-posteriors bit-exact."""
import numpy as np
import pytest

from conftest import CONFIG, FOREST

pytestmark = pytest.mark.gpu

DMIN, DMAX = 0.5, 15.0


@pytest.fixture(scope="module")
def ctx():
    import rovinasemanticsegmentation_b200 as rss
    c = rss.Context(CONFIG, FOREST, 0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def frame():
    from rovinasemanticsegmentation_b200 import synth
    rgb, depth = synth.frame(1000)
    return (rgb, depth) + synth.calibration()


def test_info(ctx):
    i = ctx.info
    assert (i.feature_length, i.num_trees, i.layer_count, i.total_classes) == (366, 4, 2, 17)
    assert list(i.class_counts[:2]) == [8, 9] and list(i.unknown_label[:2]) == [7, 8]
    assert i.rf_prediction_stride == 2 and i.sm_count > 0


@pytest.mark.parametrize("stride", [2, 5])
def test_extract_features_bit_exact(ctx, orc, frame, stride):
    rgb, depth, Kinv, R, t = frame
    f0, x0, y0 = orc.extract(orc.default_config(), stride, rgb, depth, Kinv, R, t, DMIN, DMAX)
    f1, x1, y1 = ctx.extract_features(rgb, depth, Kinv, R, t, stride, DMIN, DMAX)
    assert np.array_equal(x0, x1) and np.array_equal(y0, y1)  # same samples, raster order
    assert np.array_equal(f0[:, :363], f1[:, :363])           # Lab patch (integer arithmetic)
    assert np.array_equal(f0[:, 363], f1[:, 363])             # depth
    assert np.array_equal(f0[:, 364], f1[:, 364])             # height
    assert np.array_equal(f0[:, 365], f1[:, 365])             # normal angle
    assert f0.tobytes() == f1.tobytes()


def test_intermediates_bit_exact(ctx, orc, frame):
    rgb, depth, Kinv, R, t = frame
    H, W = depth.shape
    ctx.extract_features(rgb, depth, Kinv, R, t, 2, DMIN, DMAX, want_feats=False)
    lab, xyz, nrm = ctx.frame_intermediates(W, H)
    assert np.array_equal(lab, orc.border_reflect(orc.bgr2lab(rgb), 77))
    xyz0 = orc.cloud(depth, Kinv, R, t, DMIN, DMAX)
    assert np.array_equal(np.isnan(xyz0), np.isnan(xyz))  # invalid-depth pixels are NaN on both sides
    assert np.array_equal(xyz0[~np.isnan(xyz0)], xyz[~np.isnan(xyz)])
    nrm0 = orc.normals(xyz0)
    assert np.array_equal(np.isnan(nrm0), np.isnan(nrm))
    m = ~np.isnan(nrm0)
    assert np.array_equal(nrm0[m], nrm[m])


def test_labelled_extraction(ctx, orc, frame):
    rgb, depth, Kinv, R, t = frame
    H, W = depth.shape
    rng = np.random.default_rng(3)
    labels = rng.integers(-2, 8, size=(2, H, W)).astype(np.int8)
    for et in (0, 1):
        f0, x0, y0, l0 = orc.extract(orc.default_config(), 4, rgb, depth, Kinv, R, t, DMIN, DMAX, et, labels)
        f1, x1, y1, l1 = ctx.extract_features(rgb, depth, Kinv, R, t, 4, DMIN, DMAX, et, labels)
        assert np.array_equal(x0, x1) and np.array_equal(y0, y1) and np.array_equal(l0, l1)
        assert f0.tobytes() == f1.tobytes()


def test_forest_predict_bit_exact(ctx, orc, frame, ref_golden):
    rgb, depth, Kinv, R, t = frame
    f, xs, ys = ctx.extract_features(rgb, depth, Kinv, R, t, 2, DMIN, DMAX)
    leaf1, post1 = ctx.forest_predict(n=f.shape[0])  # device-resident features
    leaf0, post0 = orc.Forest(FOREST).predict(f)
    assert np.array_equal(leaf0, leaf1)
    assert post0.tobytes() == post1.tobytes()
    # host-supplied features + the frozen outputs of the unmodified reference
    rows = np.concatenate([ref_golden["rf_rows_color"].astype(np.float32), ref_golden["rf_rows_tail"]], axis=1)
    leaf2, post2 = ctx.forest_predict(rows)
    assert np.array_equal(leaf2, ref_golden["rf_leaf"]) and np.array_equal(post2, ref_golden["rf_post"])


@pytest.mark.parametrize("fill", [0.0, -1000.0])
def test_segment_frame_bit_exact(ctx, orc, frame, fill):
    rgb, depth, Kinv, R, t = frame
    p1 = ctx.segment_frame(rgb, depth, Kinv, R, t, fill)
    p0 = orc.segment_frame(orc.default_config(), orc.Forest(FOREST), 2, rgb, depth, Kinv, R, t, DMIN, DMAX, fill)
    assert p0.tobytes() == p1.tobytes()
    # argmax labels agree trivially then; check the layout [layer][y][x][class]
    H, W = depth.shape
    assert p1.size == 17 * H * W


def test_edge_cases(ctx, orc):
    from rovinasemanticsegmentation_b200 import synth
    Kinv, R, t = synth.calibration(64, 48)
    rgb, depth = synth.frame(5, 64, 48)
    # all depth invalid -> zero samples
    f, xs, ys = ctx.extract_features(rgb, np.zeros_like(depth), Kinv, R, t, 2, DMIN, DMAX)
    assert f.shape[0] == 0
    # small ragged image, odd stride
    f0, x0, y0 = orc.extract(orc.default_config(), 3, rgb, depth, Kinv, R, t, DMIN, DMAX)
    f1, x1, y1 = ctx.extract_features(rgb, depth, Kinv, R, t, 3, DMIN, DMAX)
    assert np.array_equal(x0, x1) and f0.tobytes() == f1.tobytes()
    # depth at the limits of the valid range (patch half-size 77 and 2)
    d2 = depth.copy()
    d2[::2, ::2] = 500
    d2[1::2, ::2] = 15000
    f0, x0, y0 = orc.extract(orc.default_config(), 1, rgb, d2, Kinv, R, t, DMIN, DMAX)
    f1, x1, y1 = ctx.extract_features(rgb, d2, Kinv, R, t, 1, DMIN, DMAX)
    assert f0.tobytes() == f1.tobytes()
