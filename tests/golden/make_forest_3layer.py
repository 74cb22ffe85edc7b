"""Generates tests/golden/forest_3layer.dat + config_3layer.json: a SMALL multi-label forest with three label layers of
4, 7 and 5 classes (16 labels = 4 float4 channel groups, layer boundaries NOT multiples of 4), trained by the
UNMODIFIED reference learner (oracle/_ref, only buildable where /root/reference is mounted) on synthetic frames.
It exercises the code paths the shipped 8 + 9 forest does not: the generic (non-tile) mean-field path of the keyframe
call, the unaligned soft-max, three layers through the up-sample / unary layout.

    python tests/golden/make_forest_3layer.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from rovinasemanticsegmentation_b200 import synth  # noqa: E402

OUT = os.path.join(HERE, "forest_3layer.dat")
CFG = os.path.join(HERE, "config_3layer.json")


def main():
    oracle.build(ref=True)
    cfg = oracle.default_config()
    Kinv, R, t = synth.calibration()
    feats = []
    for seed in (301, 302, 303):
        rgb, depth = synth.frame(seed)
        f, xs, ys = oracle.extract(cfg, 6, rgb, depth, Kinv, R, t, 0.5, 15.0)
        feats.append(f)
    feats = np.concatenate(feats)
    c = 363 // 2 - (363 // 2) % 3
    qL = np.quantile(feats[:, c], [0.25, 0.5, 0.75])
    l0 = np.digitize(feats[:, c], qL)                                               # 4 classes: centre-pixel L
    qd = np.quantile(feats[:, 363], np.linspace(0, 1, 8)[1:-1])
    l1 = np.digitize(feats[:, 363], qd)                                             # 7 classes: depth
    qh = np.quantile(feats[:, 364], np.linspace(0, 1, 6)[1:-1])
    l2 = np.digitize(feats[:, 364], qh)                                             # 5 classes: height
    labels = np.stack([l0, l1, l2], axis=1).astype(np.int32)
    assert labels.max(0).tolist() == [3, 6, 4]
    print("training on", feats.shape)
    oracle.ref_forest_train(feats, labels, OUT, num_trees=3, max_depth=12, min_split=80, threads=8)
    base = json.load(open(os.path.join(ROOT, "resources", "keyframe_config.json")))
    names = [["a%d" % k for k in range(4)], ["b%d" % k for k in range(7)], ["c%d" % k for k in range(5)]]
    base["color_codings"] = [{"name": "layer%d" % l,
                              "coding": [{"name": ("Unknown" if k == len(n) - 1 else n[k]), "color": [k, k, k], "label": k}
                                         for k in range(len(n))]} for l, n in enumerate(names)]
    json.dump(base, open(CFG, "w"), indent=1)
    print(OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
