"""Service-level drop-in: the payloads of srv/SingleFrameSegmentation.srv, built the way the node
builds them (src/segmenter.cpp:463-497), answered through rss_service_single_frame, equal the frame worker's output."""
import os

import numpy as np
import pytest

import rovinasemanticsegmentation_b200 as rss
from rovinasemanticsegmentation_b200 import service, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FOREST = os.path.join(ROOT, "tests", "golden", "forest_shared.dat")
pytestmark = pytest.mark.gpu


def _request(rgb, depth, Kinv, R, t, pad=0):
    H, W = depth.shape
    cloud = service.rectified_cloud(depth, Kinv, R, t)
    def msg(a, enc):
        row = a.reshape(H, -1).view(np.uint8)
        if pad:
            row = np.concatenate([row, np.zeros((H, pad), np.uint8)], 1)
        return service.ImageMsg(H, W, enc, row.shape[1], row.tobytes())
    return type("Req", (), {"rgb": msg(np.ascontiguousarray(rgb), "rgb8"), "depth": msg(cloud, "32FC3")})()


@pytest.mark.parametrize("wh,pad", [((160, 120), 0), ((96, 64), 8)])
def test_single_frame_service_equals_frame_worker(wh, pad):
    W, H = wh
    rgb, depth = synth.frame(11, W, H)
    Kinv, R, t = synth.calibration(W, H)
    srv = service.SegmentationServer(rss.DEFAULT_CONFIG, FOREST, Kinv, R, t, 0)
    try:
        resp = srv.segment_frame(_request(rgb, depth, Kinv, R, t, pad))
        # the node's rectification marks depth outside [0.5, 15] m invalid, exactly the frame worker's depth_min / depth_max
        want = srv.ctx.segment_frame(rgb, depth, Kinv, R, t, 0.0)
        assert resp.label_distribution.dtype == np.float32
        assert resp.label_distribution.shape == (srv.ctx.sumC * W * H,)
        assert resp.label_distribution.tobytes() == want.tobytes()
        # layout [layer][y][x][class] like the reference server's concatenation (single_frame_segmentation_server.py:47)
        info = srv.segmentation_information()
        assert sum(info["class_counts"]) == srv.ctx.sumC
        l0 = resp.label_distribution[:H * W * info["class_counts"][0]].reshape(H, W, info["class_counts"][0])
        assert np.isfinite(l0).all()
    finally:
        srv.close()


def test_map_services_and_errors():
    W, H = 64, 48
    Kinv, R, t = synth.calibration(W, H)
    srv = service.SegmentationServer(rss.DEFAULT_CONFIG, FOREST, Kinv, R, t, 0)
    try:
        info = srv.segmentation_information()
        assert info["layer_names"] == ["material", "object"] or len(info["layer_names"]) == srv.ctx.info.layer_count
        assert len(info["class_colors"]) == 3 * len(info["class_names"])
        labels = np.arange(2 * 10, dtype=np.uint8).reshape(2, 10)
        srv.store_map_result(7, labels)
        assert srv.stored_semantics_ids() == [7]
        mid, pts = srv.local_map_segmentation(7, [info["layer_names"][1], info["layer_names"][0]])
        assert mid == 7 and pts.tolist() == labels[1].tolist() + labels[0].tolist()
        assert srv.local_map_segmentation(7, ["no such layer"]) is None
        assert srv.local_map_segmentation(8, [info["layer_names"][0]]) is None
        bad = type("Req", (), {"rgb": service.ImageMsg(H, W, "mono8", W, bytes(W * H)),
                               "depth": service.ImageMsg(H, W, "32FC3", 12 * W, bytes(12 * W * H))})()
        with pytest.raises(ValueError):
            srv.segment_frame(bad)
    finally:
        srv.close()
