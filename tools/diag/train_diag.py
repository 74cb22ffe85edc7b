"""Diagnostic: GPU-trained vs reference-trained forest on synthetic frame features (node counts, depth, accuracies)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle
import rovinasemanticsegmentation_b200 as rss
from rovinasemanticsegmentation_b200 import synth
oracle.build(ref=True); oracle.set_threads(8)
cfg = oracle.default_config()
W, H = 320, 240
Kinv, R, t = synth.calibration(W, H)
fs = []
for seed in range(4):
    rgb, depth = synth.frame(seed, W, H)
    fs.append(oracle.extract(cfg, 3, rgb, depth, Kinv, R, t, 0.5, 15.0)[0].copy())
train = np.concatenate(fs[:3]); test = fs[3]
print("train", train.shape, "nan feats:", int(np.isnan(train).sum()))
thr = synth.label_thresholds(train)
ltrain, ltest = synth.labels_from_features(train, thr), synth.labels_from_features(test, thr)
cc = [int(ltrain[:, 0].max()) + 1, int(ltrain[:, 1].max()) + 1]
def stats(path, name):
    trees = oracle.read_forest_dat(path)
    for T in trees:
        depth = {0: 0}
        for i, l in enumerate(T["left"]):
            if l: depth[l] = depth[l + 1] = depth[i] + 1
        print(name, "nodes", len(T["feat"]), "maxdepth", max(depth.values()), "root", T["feat"][0], T["thr"][0])
    F = oracle.Forest(path)
    for nm, X, Y in (("train", train, ltrain), ("test", test, ltest)):
        _, post = F.predict(X); off = 0
        for l, C in enumerate(cc):
            print(name, nm, "layer", l, "acc %.4f" % (post[:, off:off + C].argmax(1) == Y[:, l]).mean()); off += C
with rss.Context(rss.DEFAULT_CONFIG, None, 0) as ctx:
    for seed in (3, 4):
        t0 = time.time()
        st = ctx.forest_train(train, ltrain, cc, "/tmp/gpu.dat", num_trees=4, max_depth=30, min_split_examples=50, seed=seed)
        print("gpu train wall %.3f s, device+host %.1f ms, nodes %d levels %d" % (time.time() - t0, st.train_ms, st.nodes, st.levels))
        stats("/tmp/gpu.dat", "GPU%d" % seed)
t0 = time.time()
oracle.ref_forest_train(train, ltrain, "/tmp/ref.dat", num_trees=4, max_depth=30, min_split=50, threads=8)
print("ref train wall %.3f s" % (time.time() - t0))
stats("/tmp/ref.dat", "REF")
