"""GPU parity tests of the permutohedral lattice and the mean-field loop, through the C ABI.
Bar: CRF marginals within 1e-4 abs of the reference path, labels >= 99.9 % agreement."""
import numpy as np
import pytest

from conftest import CONFIG, FOREST

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def ctx():
    import rovinasemanticsegmentation_b200 as rss
    c = rss.Context(CONFIG, FOREST, 0)
    yield c
    c.close()


@pytest.mark.parametrize("name", ["d6", "d5", "d3", "d2"])
def test_filter_matches_reference_golden(ctx, ref_golden, name):
    """splat/blur/slice vs outputs of the unmodified permutohedral.cpp (frozen)."""
    feats = ref_golden["lat_%s_feats" % name]
    x = ref_golden["lat_%s_in9" % name]
    crf = ctx.crf(feats.shape[0], 9)
    crf.add_pairwise(feats, 1.0)
    assert crf.lattice_size(0) == int(ref_golden["lat_%s_V" % name])  # including the reference's zero-feature padding point
    out = crf.filter(x)
    ref = ref_golden["lat_%s_out9" % name]
    assert np.abs(out - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())
    crf.close()


def test_lattice_partition_identical(ctx, orc):
    """Same simplex and barycentric weights per point: filtering an indicator reproduces the oracle exactly up to
    summation order."""
    from rovinasemanticsegmentation_b200 import synth
    xyz, col = synth.local_map(seed=3, n_points=50_000)
    f6 = orc.features_xyzrgb(xyz, col, 2.0, 4.0)
    lat = orc.Lattice(f6)
    crf = ctx.crf(f6.shape[0], 1)
    crf.add_pairwise(f6, 1.0)
    assert crf.lattice_size(0) == lat.V  # N is a multiple of 4: no padding points in the reference either
    x = np.random.default_rng(0).random((f6.shape[0], 1), dtype=np.float32)
    ref = lat.compute(x)
    assert np.abs(crf.filter(x) - ref).max() <= 2e-6 * np.abs(ref).max()  # summation order only
    crf.close()


@pytest.mark.parametrize("M", [8, 9, 3])
def test_inference_node_kernel(ctx, orc, M):
    """One 6-D Potts kernel with the node's parameters (xyz*0.5, rgb*4, w=10, 10 iterations)."""
    from rovinasemanticsegmentation_b200 import synth
    N = 60_000
    xyz, col = synth.local_map(seed=M, n_points=N)
    f6 = orc.features_xyzrgb(xyz, col, 0.5, 4.0)
    lab = (np.floor(xyz[:, 0] * 1.5).astype(int) + np.floor(xyz[:, 1]).astype(int)) % M
    U = synth.unary_from_labels(lab, M, seed=1)
    Q0 = orc.crf_inference(U, [(f6, 10.0)], 10)
    crf = ctx.crf(N, M)
    crf.set_unary(U)
    crf.add_pairwise(f6, 10.0)
    Q1, l1 = crf.inference(10, unknown=M - 1, want_labels=True)
    assert np.abs(Q0 - Q1).max() <= TOL
    l0 = orc.gated_argmax(Q0, M - 1)
    assert (l0 == l1).mean() >= 0.999
    crf.close()


def test_inference_two_kernels_two_layers(ctx, orc):
    """Gaussian 3-D + bilateral 5-D kernels, two label layers through shared lattices == per-layer reference CRFs."""
    from rovinasemanticsegmentation_b200 import synth
    W, H = 160, 120
    N = W * H
    rgb, depth = synth.frame(11, W, H)
    Kinv, R, t = synth.calibration(W, H)
    xyz = orc.cloud(depth, Kinv, R, t, 0.0, 100.0).reshape(-1, 3)
    f3 = (xyz / np.float32(0.05)).astype(np.float32)
    f5 = orc.features_bilateral2d(W, H, 80, 80, 13, 13, 13, rgb)
    rng = np.random.default_rng(5)
    labs = [rng.integers(0, 8, N), rng.integers(0, 9, N)]
    Us = [synth.unary_from_labels(labs[0], 8, 2), synth.unary_from_labels(labs[1], 9, 3)]
    crf = ctx.crf(N, [8, 9])
    crf.set_unary(Us[0], 0)
    crf.set_unary(Us[1], 1)
    crf.add_pairwise(f3, 3.0)
    crf.add_pairwise(f5, 10.0)
    Q, lab = crf.inference(10, unknown=[7, 8], want_labels=True)
    for l in range(2):
        Q0 = orc.crf_inference(Us[l], [(f3, 3.0), (f5, 10.0)], 10)
        assert np.abs(Q0 - Q[l]).max() <= TOL
        assert (orc.gated_argmax(Q0, [7, 8][l]) == lab[l]).mean() >= 0.999
    crf.close()


def test_densecrf2d_builders_and_map(ctx, orc):
    """DenseCRF2D::addPairwiseGaussian/Bilateral feature builders + plain MAP (dense_inference.cpp settings)."""
    from rovinasemanticsegmentation_b200 import synth
    W, H, M = 96, 64, 5
    rgb, _ = synth.frame(21, W, H)
    lab = (np.arange(W * H) // 700) % M
    U = synth.unary_from_labels(lab, M, 4)
    crf = ctx.crf(W * H, M)
    crf.set_unary(U)
    crf.add_pairwise_gaussian(W, H, 3, 3, 3.0)
    crf.add_pairwise_bilateral(W, H, 80, 80, 13, 13, 13, rgb, 10.0)
    Q1, l1 = crf.inference(5, want_labels=True)
    Q0 = orc.crf_inference(U, [(orc.features_gaussian2d(W, H, 3, 3), 3.0),
                               (orc.features_bilateral2d(W, H, 80, 80, 13, 13, 13, rgb), 10.0)], 5)
    assert np.abs(Q0 - Q1).max() <= TOL
    assert (Q0.argmax(1) == l1).mean() >= 0.999
    crf.close()


def test_zero_iterations_and_no_pairwise(ctx, orc):
    from rovinasemanticsegmentation_b200 import synth
    N, M = 1000, 4
    U = synth.unary_from_labels(np.arange(N) % M, M, 9)
    crf = ctx.crf(N, M)
    crf.set_unary(U)
    Q = crf.inference(0)
    assert np.abs(Q - orc.crf_inference(U, [], 0)).max() <= 5e-6
    crf.close()


def test_unary_accumulate_and_map_worker(ctx, orc):
    """segmenter.cpp:597-657: scatter-add posteriors through an index image, xyz/rgb kernel, gated argmax."""
    from rovinasemanticsegmentation_b200 import synth
    N = 30_000
    xyz, col = synth.local_map(seed=8, n_points=N)
    rng = np.random.default_rng(2)
    npix = 64 * 48
    Ms = [8, 9]
    crf = ctx.crf(N, Ms)
    un = [np.zeros((N, m), np.float32) for m in Ms]
    for k in range(3):
        idx = rng.integers(-1, N, npix).astype(np.int32)
        idx[rng.random(npix) < 0.3] = -1
        idx[np.unique(idx[idx >= 0], return_index=True)[1]]  # (indices may repeat: the adds must accumulate)
        post = [np.log(rng.dirichlet(np.ones(m), npix)).astype(np.float32) * 2 for m in Ms]
        crf.unary_accumulate(idx, np.concatenate([p.reshape(-1) for p in post]))
        for l in range(2):
            orc.unary_accumulate(idx, post[l], un[l])
    crf.add_pairwise_xyzrgb(xyz, col, 0.5, 4.0, 10.0)
    Q, lab = crf.inference(10, unknown=[7, 8], want_labels=True)
    f6 = orc.features_xyzrgb(xyz, col, 0.5, 4.0)
    for l in range(2):
        Q0 = orc.crf_inference(-un[l], [(f6, 10.0)], 10)
        assert np.abs(Q0 - Q[l]).max() <= TOL
        assert (orc.gated_argmax(Q0, [7, 8][l]) == lab[l]).mean() >= 0.999
    crf.close()


def test_keyframe_fused(ctx, orc):
    """rss_segment_keyframe == oracle frame path + per-layer oracle CRF (configs[1]/[2])."""
    import rovinasemanticsegmentation_b200 as rss
    from rovinasemanticsegmentation_b200 import synth
    W, H = 320, 240
    rgb, depth = synth.frame(31, W, H)
    Kinv, R, t = synth.calibration(W, H)
    prm = rss.KeyframeParams(0.05, 3.0, 80.0, 13.0, 10.0, 10, 0.0)
    labels, Q = ctx.segment_keyframe(rgb, depth, Kinv, R, t, prm, want_Q=True)
    post = orc.segment_frame(orc.default_config(), orc.Forest(FOREST), 2, rgb, depth, Kinv, R, t, 0.5, 15.0, 0.0)
    xyz = orc.cloud(depth, Kinv, R, t, 0.5, 15.0).reshape(-1, 3)
    bad = np.isnan(xyz[:, 0])
    xyz[bad] = t
    f3 = (xyz * np.float32(1.0 / 0.05)).astype(np.float32)
    f5 = orc.features_bilateral2d(W, H, 80, 80, 13, 13, 13, rgb)
    off = 0
    N = W * H
    for l, (M, unk) in enumerate(((8, 7), (9, 8))):
        U = -post[off:off + N * M].reshape(N, M)
        Q0 = orc.crf_inference(U, [(f3, 3.0), (f5, 10.0)], 10)
        Q1 = Q[off:off + N * M].reshape(N, M)
        assert np.abs(Q0 - Q1).max() <= TOL
        assert (orc.gated_argmax(Q0, unk) == labels[l]).mean() >= 0.999
        off += N * M


@pytest.mark.parametrize("k", [0, 5])
def test_keyframe_fused_640x480(ctx, orc, k):
    """The EXACT workload bench.py times (BASELINE configs[1]+[2]: 640x480, bench.KF parameters, the bench's own frame seeds;
    1080 tiles, two CTA waves, V ~ 40 k + 11 k) against the oracle: frame path + per-layer CRF, 10 iterations.
    Reference: src/segmenter.cpp:349-434, :639-657."""
    import bench
    import rovinasemanticsegmentation_b200 as rss
    from rovinasemanticsegmentation_b200 import synth
    W, H = bench.W, bench.H
    KF = bench.KF
    rgb, depth = synth.frame(bench.frame_seeds(0)[k], W, H)
    Kinv, R, t = synth.calibration(W, H)
    prm = rss.KeyframeParams(KF["sigma_xyz"], KF["w_gauss"], KF["sigma_px"], KF["sigma_rgb"], KF["w_bilateral"], KF["iters"],
                             KF["fill"])
    labels, Q = ctx.segment_keyframe(rgb, depth, Kinv, R, t, prm, want_Q=True)
    l0, Q0, _ = orc.keyframe(orc.Forest(FOREST), rgb, depth, Kinv, R, t, **KF)
    off, N = 0, W * H
    for l, M in enumerate((8, 9)):
        Q1 = Q[off:off + N * M].reshape(N, M)
        assert np.abs(Q0[l] - Q1).max() <= TOL
        assert (l0[l] == labels[l]).mean() >= 0.999
        off += N * M
    # the resident-frame entry (rgb = depth = NULL), which is what the bench's `value` pass calls, gives the same labels
    again = ctx.segment_keyframe(None, None, Kinv, R, t, prm, W=W, H=H)
    assert (again == labels).mean() >= 0.9999  # float atomics in the splat: not bit-reproducible run to run


def test_unary_accumulate_from_resident_posteriors(ctx, orc):
    """Frame worker -> map worker without a host round trip: rss_segment_frame leaves the posteriors on the device and
    rss_crf_unary_accumulate(posteriors = NULL) scatters them through the index image (segmenter.cpp:597-616)."""
    from rovinasemanticsegmentation_b200 import synth
    W, H = 160, 120
    rgb, depth = synth.frame(91, W, H)
    Kinv, R, t = synth.calibration(W, H)
    N = 5000
    rng = np.random.default_rng(4)
    idx = rng.integers(-1, N, W * H).astype(np.int32)
    post = ctx.segment_frame(rgb, depth, Kinv, R, t, 0.0)               # host copy for the expectation ...
    crf_a = ctx.crf(N, [8, 9])
    crf_a.unary_accumulate(idx, post)                                    # ... through the host path
    ctx.segment_frame(rgb, depth, Kinv, R, t, 0.0, want_host=False)      # resident only
    crf_b = ctx.crf(N, [8, 9])
    crf_b.unary_accumulate(idx)                                          # device-resident posteriors
    xyz, col = synth.local_map(seed=9, n_points=N)
    for c in (crf_a, crf_b):
        c.add_pairwise_xyzrgb(xyz, col, 0.5, 4.0, 10.0)
    Qa, Qb = crf_a.inference(3), crf_b.inference(3)
    for l in range(2):
        assert np.abs(Qa[l] - Qb[l]).max() <= 1e-5   # float atomics: the accumulation order differs
    # and against the oracle
    un = [np.zeros((N, m), np.float32) for m in (8, 9)]
    off = 0
    for l, m in enumerate((8, 9)):
        orc.unary_accumulate(idx, post[off:off + W * H * m].reshape(W * H, m), un[l])
        off += W * H * m
    f6 = orc.features_xyzrgb(xyz, col, 0.5, 4.0)
    for l in range(2):
        assert np.abs(orc.crf_inference(-un[l], [(f6, 10.0)], 3) - Qb[l]).max() <= TOL
    crf_a.close(); crf_b.close()
