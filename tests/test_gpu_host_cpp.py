"""GPU tests of the C++ host side: the reference-named adapters (host/rss_adapters.hpp) and the multi-GPU keyframe
worker (host/keyframe_worker.cpp), compiled with g++ against librss.so and checked against the oracle."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import CONFIG, FOREST, ROOT

pytestmark = pytest.mark.gpu
PKG = os.path.join(ROOT, "rovinasemanticsegmentation_b200")


def _write_inputs(tmp, frames, Kinv, R, t):
    with open(os.path.join(tmp, "frames.raw"), "wb") as f:
        for rgb, depth in frames:
            f.write(np.ascontiguousarray(rgb, np.uint8).tobytes())
            f.write(np.ascontiguousarray(depth, np.uint16).tobytes())
    np.concatenate([Kinv.reshape(-1), R.reshape(-1), t.reshape(-1)]).astype(np.float32).tofile(os.path.join(tmp, "calib.raw"))


def test_reference_named_adapters(orc, tmp_path):
    from rovinasemanticsegmentation_b200 import synth
    tmp = str(tmp_path)
    W, H = 160, 120
    rgb, depth = synth.frame(41, W, H)
    Kinv, R, t = synth.calibration(W, H)
    _write_inputs(tmp, [(rgb, depth)], Kinv, R, t)
    exe = os.path.join(ROOT, "tests", "cpp", "adapters_check")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-o", exe, exe + ".cpp", "-L" + PKG, "-lrss",
                           "-Wl,-rpath," + PKG])
    subprocess.check_call([exe, CONFIG, FOREST, os.path.join(tmp, "frames.raw"), os.path.join(tmp, "calib.raw"), str(W), str(H), tmp])
    ld = lambda n, dt: np.fromfile(os.path.join(tmp, n), dt)
    # FeatureExtractor + RandomForest: bit-exact
    f0, x0, y0 = orc.extract(orc.default_config(), 2, rgb, depth, Kinv, R, t, 0.5, 15.0)
    assert np.array_equal(ld("xs.bin", np.int32), x0) and np.array_equal(ld("ys.bin", np.int32), y0)
    assert ld("feats.bin", np.float32).tobytes() == f0.tobytes()
    leaf0, post0 = orc.Forest(FOREST).predict(f0)
    assert np.array_equal(ld("leaves.bin", np.int32).reshape(leaf0.shape), leaf0)
    assert ld("post.bin", np.float32).tobytes() == post0.tobytes()
    n = f0.shape[0]
    assert ld("post_single.bin", np.float32).tobytes() == post0[::n // 7 + 1].tobytes()
    # SingleFrameSegmentation service adapter == the frame worker (fill 0), bit for bit
    post_frame = orc.segment_frame(orc.default_config(), orc.Forest(FOREST), 2, rgb, depth, Kinv, R, t, 0.5, 15.0, 0.0)
    assert ld("service.bin", np.float32).tobytes() == post_frame.tobytes()
    # DenseCRF2D / DenseCRF
    M, N = 5, W * H
    U = ld("unary.bin", np.float32).reshape(N, M)
    Q0 = orc.crf_inference(U, [(orc.features_gaussian2d(W, H, 3, 3), 3.0),
                               (orc.features_bilateral2d(W, H, 80, 80, 13, 13, 13, rgb), 10.0)], 5)
    Q1 = ld("Q.bin", np.float32).reshape(N, M)
    assert np.abs(Q0 - Q1).max() <= 1e-4
    assert (ld("map.bin", np.int16) == Q0.argmax(1)).mean() >= 0.999
    assert np.abs(ld("Qstep.bin", np.float32).reshape(N, M) - Q0).max() <= 1e-4  # start + 5 steps == inference(5)
    assert (ld("mapstep.bin", np.int16) == Q0.argmax(1)).mean() >= 0.999
    f3 = ld("feats3.bin", np.float32).reshape(N, 3)
    Q2 = orc.crf_inference(U, [(f3, 10.0)], 10)
    assert (ld("gated.bin", np.uint8) == orc.gated_argmax(Q2, M - 1)).mean() >= 0.999


def test_keyframe_worker_matches_binding(tmp_path):
    """The C++ worker pool (2 contexts in flight) returns the label maps the Python binding returns."""
    import rovinasemanticsegmentation_b200 as rss
    from rovinasemanticsegmentation_b200 import build, synth
    tmp = str(tmp_path)
    W, H = 320, 240
    frames = [synth.frame(50 + k, W, H) for k in range(3)]
    Kinv, R, t = synth.calibration(W, H)
    _write_inputs(tmp, frames, Kinv, R, t)
    exe = build.build_host()
    out = subprocess.check_output([exe, "--config", CONFIG, "--forest", FOREST, "--frames-file", os.path.join(tmp, "frames.raw"),
                                   "--calib", os.path.join(tmp, "calib.raw"), "--width", str(W), "--height", str(H),
                                   "--frames", "6", "--inflight", "2", "--out", os.path.join(tmp, "labels.bin")], text=True)
    line = json.loads(out.strip().splitlines()[-1])
    assert line["value"] > 0 and line["keyframes"] == 6
    got = np.fromfile(os.path.join(tmp, "labels.bin"), np.uint8).reshape(3, 2, H * W)
    prm = rss.KeyframeParams(0.05, 3.0, 80.0, 13.0, 10.0, 10, 0.0)
    with rss.Context(CONFIG, FOREST, 0) as ctx:
        for k, (rgb, depth) in enumerate(frames):
            ref = ctx.segment_keyframe(rgb, depth, Kinv, R, t, prm)
            assert (ref == got[k]).mean() >= 0.999  # splat uses float atomics: summation order varies run to run
