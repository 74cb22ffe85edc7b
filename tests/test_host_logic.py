"""Host-side logic that needs no GPU: service message marshalling (srv/SingleFrameSegmentation.srv), training defaults,
the libforest model parser used by the training tests."""
import ctypes as C
import os

import numpy as np
import pytest

import rovinasemanticsegmentation_b200 as rss
from rovinasemanticsegmentation_b200 import service, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FOREST = os.path.join(ROOT, "tests", "golden", "forest_shared.dat")


def test_train_params_default_follow_the_reference_learner():
    lib = rss.load_library()
    p = rss.TrainParams()
    lib.rss_train_params_default(C.byref(p))
    # DecisionTreeLearner() defaults (learning.h:109-118) + src/train.cpp:225-249 + resources/config.json
    assert (p.num_trees, p.max_depth, p.min_split_examples, p.min_child_split_examples) == (4, 30, 50, 1)
    assert p.num_features == 0 and p.use_bootstrap == 1 and p.num_bootstrap_examples == 0 and p.smoothing == 1.0


def test_rectified_cloud_is_the_nodes_request_payload():
    """src/segmenter.cpp:463-488: (R * Kinv) * (d x, d y, d) + t, NaN outside [0.5, 15] m; the camera-frame z of a valid point
    is the raw depth (what rss_service_single_frame relies on)."""
    W, H = 64, 48
    rgb, depth = synth.frame(5, W, H)
    Kinv, R, t = synth.calibration(W, H)
    cloud = service.rectified_cloud(depth, Kinv, R, t)
    assert cloud.shape == (H, W, 3) and cloud.dtype == np.float32
    d = depth.astype(np.float32) / np.float32(1000.0)
    bad = (d < 0.5) | (d > 15.0)
    assert np.isnan(cloud[bad]).all() and np.isfinite(cloud[~bad]).all() and bad.any() and (~bad).any()
    Rm = np.asarray(R, np.float64).reshape(3, 3)
    z = (cloud[~bad].astype(np.float64) - np.asarray(t, np.float64)) @ Rm[:, 2]
    assert np.array_equal(np.rint(z * 1000.0).astype(np.int64), depth[~bad].astype(np.int64))


def test_image_message_marshalling():
    H, W = 5, 7
    a = np.arange(H * W * 3, dtype=np.uint8).reshape(H, W, 3)
    packed = service.ImageMsg(H, W, "rgb8", 3 * W, a.tobytes())
    assert np.array_equal(service.image_to_array(packed), a)
    padded_rows = np.concatenate([a.reshape(H, -1), np.zeros((H, 5), np.uint8)], 1)
    padded = service.ImageMsg(H, W, "rgb8", 3 * W + 5, padded_rows.tobytes())
    assert np.array_equal(service.image_to_array(padded), a)
    f = np.linspace(0, 1, H * W * 3, dtype=np.float32).reshape(H, W, 3)
    assert np.array_equal(service.image_to_array(service.ImageMsg(H, W, "32FC3", 12 * W, f.tobytes())), f)
    with pytest.raises(ValueError):
        service.image_to_array(service.ImageMsg(H, W, "bgr8", 3 * W, a.tobytes()))
    with pytest.raises(ValueError):
        service.image_to_array(service.ImageMsg(H, W, "rgb8", 3 * W, a.tobytes()[:-1]))


def test_forest_dat_parser_matches_the_oracle_loader(orc):
    trees = orc.read_forest_dat(FOREST)
    F = orc.Forest(FOREST)
    assert len(trees) == F.T and [len(t["feat"]) for t in trees] == F.nodes
    for t in trees:
        leaves = t["left"] == 0
        assert leaves.sum() == (~leaves).sum() + 1  # a binary tree
        for i in np.flatnonzero(leaves)[:20]:
            assert [len(h) for h in t["multi"][i]] == F.classes


def test_reference_gradient_harness_is_consistent(orc):
    """oracle/_ref's DenseCRF::gradient (the checker of rss_crf_gradient): its Potts-weight gradient is the derivative of its
    own objective (central finite differences), for the symmetric normalisation of the reference's path."""
    if not orc.ref_available():
        pytest.skip("oracle/_ref is not built here")
    rng = np.random.default_rng(0)
    N, M = 3000, 5
    f = np.stack([rng.uniform(0, 30, N), rng.uniform(0, 30, N), rng.uniform(0, 8, N)], 1).astype(np.float32)
    U = rng.random((N, M), dtype=np.float32) * 2
    gt = rng.integers(-1, M, N).astype(np.int16)
    r, g = orc.ref_crf_gradient(U, [(f, 3.0), (f[:, :2] * np.float32(0.5), 1.5)], 3, gt)
    eps = 1e-2
    r1, _ = orc.ref_crf_gradient(U, [(f, 3.0 + eps), (f[:, :2] * np.float32(0.5), 1.5)], 3, gt)
    r0, _ = orc.ref_crf_gradient(U, [(f, 3.0 - eps), (f[:, :2] * np.float32(0.5), 1.5)], 3, gt)
    assert r < 0 and abs((r1 - r0) / (2 * eps) - g[0]) <= 0.01 * abs(g[0])


def test_meanfield_plans_and_layout_helpers(tmp_path):
    """Host-side logic of the fused mean-field path (csrc/meanfield.cu / meanfield.cuh), checked without a GPU by a small nvcc
    program linked against librss.so (tests/cpp/plans_check.cu): the cooperative blur's phase plan covers every lattice axis
    exactly once, in order, in consecutive phases (permutohedral.cpp:555-569 blurs axes 0..d in turn - the ping / pong parity
    of the result depends on the phase count), the Q-tile row rotation is a permutation inside a warp, the point-major
    (corner, point) index is a bijection, and the tile maps of image / 1-D point sets have the expected tile counts."""
    import subprocess
    root = os.path.dirname(os.path.dirname(rss.__file__))
    lib_dir = os.path.dirname(rss.LIB_PATH)
    assert os.path.exists(rss.LIB_PATH), "librss.so is not built"
    exe = str(tmp_path / "plans_check")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    subprocess.check_call([nvcc, "-std=c++17", "-O1", "-Wno-deprecated-gpu-targets",
                           "-I" + os.path.join(root, "rovinasemanticsegmentation_b200", "csrc"), "-I" + os.path.join(root, "include"),
                           "-o", exe, os.path.join(root, "tests", "cpp", "plans_check.cu"), "-L" + lib_dir, "-lrss",
                           "-Xlinker", "-rpath", "-Xlinker", lib_dir])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "plans_check OK" in r.stdout, r.stdout + r.stderr
