"""GPU tests of the map worker with everything on the device (src/segmenter.cpp:559-657): the local map's
cloud resident in the CRF, key-frame posteriors kept on the device, the library's own z-buffer projector instead of a
host index image, unary accumulation, the 6-D kernel from the resident cloud, mean field, gated argmax."""
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


def test_projector_index_image_bit_exact(ctx, orc):
    """The z-buffer kernel against the oracle's serial projector: identical index images (nearest point per pixel, equal
    depth -> lower index), including points behind the camera, outside the image and outside [zmin, zmax]."""
    from rovinasemanticsegmentation_b200 import synth
    N, W, H = 400_000, 320, 240
    xyz, col = synth.local_map(seed=11, n_points=N)
    xyz[1000:1100] = xyz[0:100]  # exact duplicates: equal z, the lower index must win
    K = synth.intrinsics(W, H)
    crf = ctx.crf(N, [8, 9])
    crf.set_cloud(xyz, col)
    rgb, depth = synth.frame(5, W, H)
    Kinv, Rc, tc = synth.calibration(W, H)
    ctx.segment_frame(rgb, depth, Kinv, Rc, tc, 0.0, want_host=False)
    for k in range(4):
        R, t = synth.map_keyframe_pose(k, 4)
        idx = crf.project_accumulate(W, H, K, R, t, 0.3, 8.0, want_index=True)
        ref = orc.project_zbuffer(xyz, K, R, t, W, H, 0.3, 8.0)
        assert np.array_equal(idx, ref)
        assert (idx >= 0).mean() > 0.05  # the camera does see the room
    crf.close()


def test_map_worker_on_device_vs_oracle(ctx, orc):
    """Three key frames segmented first (posteriors kept in device slots, like the frame worker running ahead of the map
    worker), then fused into one local map: projector -> accumulate -> xyz/rgb kernel -> mean field -> gated argmax.  The
    oracle does the same with its own projector, unary accumulation and CRF."""
    from rovinasemanticsegmentation_b200 import synth
    N, W, H, KF = 300_000, 320, 240, 3
    xyz, col = synth.local_map(seed=13, n_points=N)
    K = synth.intrinsics(W, H)
    Kinv, Rc, tc = synth.calibration(W, H)
    posts = []
    for k in range(KF):
        rgb, depth = synth.frame(40 + k, W, H)
        posts.append(ctx.segment_frame(rgb, depth, Kinv, Rc, tc, 0.0))  # host copy only for the oracle's side
        ctx.posteriors_keep(k)
    crf = ctx.crf(N, [8, 9])
    crf.set_cloud(xyz, col)
    un = [np.zeros((N, m), np.float32) for m in (8, 9)]
    for k in range(KF):
        R, t = synth.map_keyframe_pose(k, KF)
        crf.project_accumulate(W, H, K, R, t, 0.3, 8.0, slot=k)
        idx = orc.project_zbuffer(xyz, K, R, t, W, H, 0.3, 8.0)
        off = 0
        for l, m in enumerate((8, 9)):
            orc.unary_accumulate(idx, posts[k][off:off + W * H * m].reshape(W * H, m), un[l])
            off += W * H * m
    crf.add_pairwise_cloud(0.5, 4.0, 10.0)
    Q, lab = crf.inference(5, unknown=[7, 8], want_labels=True)
    f6 = orc.features_xyzrgb(xyz, col, 0.5, 4.0)
    for l in range(2):
        Q0 = orc.crf_inference(-un[l], [(f6, 10.0)], 5)
        assert np.abs(Q0 - Q[l]).max() <= 1e-4
        assert (orc.gated_argmax(Q0, [7, 8][l]) == lab[l]).mean() >= 0.999
    crf.close()
    # state errors: no cloud, no kept posteriors in the slot
    import rovinasemanticsegmentation_b200 as rss
    crf = ctx.crf(1000, [8, 9])
    with pytest.raises(rss.RssError):
        crf.project_accumulate(W, H, K, R, t, 0.3, 8.0)
    crf.set_cloud(xyz[:1000], col[:1000])
    with pytest.raises(rss.RssError):
        crf.project_accumulate(W, H, K, R, t, 0.3, 8.0, slot=77)
    crf.close()


def test_sorted_fused_path_for_incoherent_point_sets(ctx, orc, monkeypatch):
    """A local map's points come in no particular order: the library sorts them by their first two lattice vertices and
    runs the fused point kernel over the sorted order (crf.cu, TileMap::perm).  Same marginals as the oracle and as the
    generic kernels; a set that stays incoherent after sorting (fine scales: about as many vertices as points) keeps the
    generic path."""
    from rovinasemanticsegmentation_b200 import synth
    N = 150_000
    xyz, col = synth.local_map(seed=3, n_points=N)
    rng = np.random.default_rng(1)
    shuffle = rng.permutation(N)  # make sure the input order carries no coherence at all
    xyz, col = np.ascontiguousarray(xyz[shuffle]), np.ascontiguousarray(col[shuffle])
    U = [rng.random((N, m), dtype=np.float32) * 3 for m in (8, 9)]

    def run(wxyz, wrgb):
        crf = ctx.crf(N, [8, 9])
        for l in range(2):
            crf.set_unary(U[l], l)
        crf.add_pairwise_xyzrgb(xyz, col, wxyz, wrgb, 10.0)
        path = crf.path()
        Q, labels = crf.inference(5, unknown=[7, 8], want_Q=True, want_labels=True)
        V = crf.lattice_size(0)
        crf.close()
        return path, Q, labels, V

    (fused, sorted_), Q, labels, V = run(0.5, 4.0)
    assert fused and sorted_ and V < N // 20
    f6 = orc.features_xyzrgb(xyz, col, 0.5, 4.0)
    for l, (M, unk) in enumerate(((8, 7), (9, 8))):
        Q0 = orc.crf_inference(U[l], [(f6, 10.0)], 5)
        assert np.abs(Q0 - Q[l]).max() <= 1e-4
        assert (orc.gated_argmax(Q0, unk) == labels[l]).mean() >= 0.999
    monkeypatch.setenv("RSS_NO_POINT_SORT", "1")
    (fused_g, sorted_g), Qg, labels_g, _ = run(0.5, 4.0)
    monkeypatch.delenv("RSS_NO_POINT_SORT")
    assert not fused_g and not sorted_g
    assert max(np.abs(Qg[l] - Q[l]).max() for l in range(2)) <= 1e-4 and (labels_g == labels).mean() >= 0.999
    (fused_f, sorted_f), _, _, Vf = run(20.0, 40.0)
    assert not fused_f and not sorted_f and Vf > N // 2
