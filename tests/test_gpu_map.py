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
