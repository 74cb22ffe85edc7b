"""Generates the committed fixtures in tests/golden/.  Run HERE (the build container), where
/root/reference (-> oracle/_ref/libref_oracle.so) and cv2 4.13.0 exist; neither travels to the GPU box.

    python tests/golden/make_golden.py [--retrain]

Outputs
  forest_shared.dat   multi-label forest (material 8 + object 9 classes) trained by the UNMODIFIED
                      reference learner (third-party/libforest/src/learning.cpp) with the settings of
                      src/train.cpp:225-249 / resources/config.json:38-40 on synthetic frames.
                      Training is unseeded in the reference (std::random_device), so the file is FROZEN:
                      it is only regenerated with --retrain.
  cv_golden.npz       cv2 4.13.0 outputs: cvtColor(BGR2Lab) 8U, copyMakeBorder(REFLECT),
                      resize(INTER_LINEAR) 8UC3 windows -> 11x11, resize(INTER_LINEAR) 32FC(8|9) 2x.
  ref_golden.npz      outputs of the unmodified reference: RandomForest::multiClassLogPosterior +
                      DecisionTree::findLeafNode on sample rows; Permutohedral::init/compute on small clouds.
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from rovinasemanticsegmentation_b200 import synth  # noqa: E402

FOREST = os.path.join(HERE, "forest_shared.dat")
DMIN, DMAX = 0.5, 15.0


def train_forest():
    oracle.set_threads(8)
    cfg = oracle.default_config()
    Kinv, R, t = synth.calibration()
    feats = []
    for seed in range(6):
        rgb, depth = synth.frame(seed)
        f, xs, ys = oracle.extract(cfg, 5, rgb, depth, Kinv, R, t, DMIN, DMAX)  # training_sample_stride 5
        feats.append(f.copy())
    feats = np.concatenate(feats)
    thr = synth.label_thresholds(feats)
    labels = synth.labels_from_features(feats, thr)
    for l, C in enumerate((8, 9)):
        cnt = np.bincount(labels[:, l], minlength=C)
        assert (cnt > 0).all() and len(cnt) == C, cnt
    print("training on", feats.shape, "samples")
    oracle.ref_forest_train(feats, labels, FOREST, num_trees=4, max_depth=30, min_split=50, threads=8)
    print("wrote", FOREST, os.path.getsize(FOREST), "bytes")


def cv_golden():
    import cv2
    cv2.setNumThreads(1)
    rng = np.random.default_rng(1234)
    out = {}
    cols = rng.integers(0, 256, size=(64, 64, 3), dtype=np.uint8)
    cols[0, :16] = np.array([[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 255, 0], [0, 0, 255], [1, 1, 1], [254, 254, 254],
                             [128, 128, 128], [12, 200, 77], [49, 49, 49], [77, 13, 240], [3, 2, 1], [10, 10, 10],
                             [11, 11, 11], [200, 100, 50], [50, 100, 200]], np.uint8)
    out["lab_src"] = cols
    out["lab_dst"] = cv2.cvtColor(cols, cv2.COLOR_BGR2Lab)
    img = rng.integers(0, 256, size=(9, 13, 3), dtype=np.uint8)
    out["border_src"] = img
    out["border_dst_5"] = cv2.copyMakeBorder(img, 5, 5, 5, 5, cv2.BORDER_REFLECT)
    out["border_dst_20"] = cv2.copyMakeBorder(img, 20, 20, 20, 20, cv2.BORDER_REFLECT)
    sizes = [1, 3, 5, 7, 9, 11, 13, 21, 33, 55, 77, 101, 155]
    big = rng.integers(0, 256, size=(160, 160, 3), dtype=np.uint8)
    out["resize_src"] = big
    out["resize_sizes"] = np.array(sizes, np.int32)
    for S in sizes:
        out["resize_dst_%d" % S] = cv2.resize(np.ascontiguousarray(big[2:2 + S, 3:3 + S]), (11, 11))
        out["resize5_dst_%d" % S] = cv2.resize(np.ascontiguousarray(big[2:2 + S, 3:3 + S]), (5, 5))
    for C in (8, 9):
        src = (rng.standard_normal((12, 16, C)) * 4).astype(np.float32)
        src[rng.random((12, 16)) < 0.25] = 0
        out["up_src_%d" % C] = src
        out["up_dst_%d" % C] = cv2.resize(src, (32, 24))
        out["up3_dst_%d" % C] = cv2.resize(src, (48, 36))
    np.savez_compressed(os.path.join(HERE, "cv_golden.npz"), **out)
    print("wrote cv_golden.npz")


def ref_golden():
    oracle.set_threads(8)
    out = {}
    cfg = oracle.default_config()
    Kinv, R, t = synth.calibration()
    rgb, depth = synth.frame(1000)
    f, xs, ys = oracle.extract(cfg, 2, rgb, depth, Kinv, R, t, DMIN, DMAX)
    rng = np.random.default_rng(99)
    rows = np.sort(rng.choice(f.shape[0], size=768, replace=False))
    # store the colour part as u8 (exact small ints) to keep the fixture small
    out["rf_rows_color"] = f[rows, :363].astype(np.uint8)
    out["rf_rows_tail"] = f[rows, 363:].copy()
    rf = oracle.RefForest(FOREST)
    leaf, post = rf.predict(f[rows], 17)
    out["rf_leaf"] = leaf
    out["rf_post"] = post
    # lattices
    for name, d, N in (("d6", 6, 3001), ("d5", 5, 2500), ("d3", 3, 2048), ("d2", 2, 1777)):
        xyz, col = synth.local_map(seed=d, n_points=N)
        if d == 6:
            feats = oracle.features_xyzrgb(xyz, col, 0.5 * 8, 4.0)
        elif d == 5:
            feats = np.concatenate([xyz[:, :2] * 3, col * 9], axis=1).astype(np.float32)
        elif d == 3:
            feats = (xyz * 2.5).astype(np.float32)
        else:
            feats = (xyz[:, :2] * 4).astype(np.float32)
        lat = oracle.RefLattice(feats)
        off, bary = lat.get()
        x9 = rng.random((N, 9), dtype=np.float32)
        x1 = np.ones((N, 1), np.float32)
        out["lat_%s_feats" % name] = feats
        out["lat_%s_V" % name] = np.int32(lat.V)
        out["lat_%s_off" % name] = off
        out["lat_%s_bary" % name] = bary
        out["lat_%s_in9" % name] = x9
        out["lat_%s_out9" % name] = lat.compute(x9)
        out["lat_%s_out1" % name] = lat.compute(x1)
    np.savez_compressed(os.path.join(HERE, "ref_golden.npz"), **out)
    print("wrote ref_golden.npz")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--retrain", action="store_true")
    a = ap.parse_args()
    oracle.build(ref=True)
    if a.retrain or not os.path.exists(FOREST):
        train_forest()
    cv_golden()
    ref_golden()
