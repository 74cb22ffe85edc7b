"""Build / filter / inference timings of an N-point local map (generic path), with the per-kernel break-down.
    python tools/scale_probe.py N wxyz wrgb"""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import rovinasemanticsegmentation_b200 as rss
from rovinasemanticsegmentation_b200 import synth
N = int(sys.argv[1]); wxyz = float(sys.argv[2]); wrgb = float(sys.argv[3])
xyz, col = synth.local_map(seed=43, n_points=N)
ctx = rss.Context(rss.DEFAULT_CONFIG, None, 0)
for rep in range(2):
    t = time.time(); crf = ctx.crf(N, [8, 9])
    for l, M in enumerate((8, 9)): crf.set_unary(np.random.default_rng(l).random((N, M), dtype=np.float32), l)
    t = time.time(); crf.add_pairwise_xyzrgb(xyz, col, wxyz, wrgb, 10.0); print("build", round(time.time() - t, 4), "V", crf.lattice_size(0), flush=True)
    if rep == 1: ctx.profile_enable(True)
    lab = crf.inference(10, unknown=[7, 8], want_Q=False, want_labels=True); print("meanfield ms/iter", ctx.timings()["meanfield_ms"] / 10, flush=True)
    if rep == 1:
        for k, (ms, n) in sorted(ctx.profile_report().items(), key=lambda kv: -kv[1][0])[:8]: print("   %-28s %8.3f ms  n=%d  %.1f us/launch" % (k, ms, n, 1000 * ms / n))
    crf.close()
