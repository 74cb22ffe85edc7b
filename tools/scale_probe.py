import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import rovinasemanticsegmentation_b200 as rss
from rovinasemanticsegmentation_b200 import synth
N = int(sys.argv[1]); wxyz = float(sys.argv[2]); wrgb = float(sys.argv[3])
xyz, col = synth.local_map(seed=43, n_points=N)
ctx = rss.Context(rss.DEFAULT_CONFIG, None, 0)
t = time.time(); crf = ctx.crf(N, 1); crf.add_pairwise_xyzrgb(xyz, col, wxyz, wrgb, 1.0); print("build", time.time() - t, "V", crf.lattice_size(0), flush=True)
x = np.random.default_rng(0).random((N, 1), dtype=np.float32)
t = time.time(); y = crf.filter(x); print("filter", time.time() - t, float(y.mean()), flush=True)
crf.close()
t = time.time(); crf = ctx.crf(N, [8, 9]); 
for l, M in enumerate((8, 9)): crf.set_unary(np.random.default_rng(l).random((N, M), dtype=np.float32), l)
crf.add_pairwise_xyzrgb(xyz, col, wxyz, wrgb, 10.0); print("build2", time.time() - t, flush=True)
t = time.time(); lab = crf.inference(10, unknown=[7, 8], want_Q=False, want_labels=True); print("inference", time.time() - t, ctx.timings()["meanfield_ms"], flush=True)
