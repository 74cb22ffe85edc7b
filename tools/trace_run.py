import sys, ctypes as C
sys.path.insert(0, "/root/repo")
import numpy as np
import rovinasemanticsegmentation_b200 as rss
from rovinasemanticsegmentation_b200 import synth
ctx = rss.Context(rss.DEFAULT_CONFIG, "/root/repo/tests/golden/forest_shared.dat", 0)
rgb, depth = synth.frame(10000); Kinv, R, t = synth.calibration()
prm = rss.KeyframeParams(0.05, 3.0, 80.0, 13.0, 10.0, 10, 0.0)
for k in range(4): ctx.segment_keyframe(rgb, depth, Kinv, R, t, prm)
buf = (C.c_ulonglong * 128)()
ctx._lib.rss_debug_trace(buf)
a = np.array(buf[:], dtype=np.int64).reshape(4, 2, 16)
names = ["start", "setup", "phase1", "sync", "splatA", "splatB"]
for mode in (2, 3, 1):
    for b in (0, 1):
        tr = a[mode, b]; t0 = tr[0]
        print("mode", mode, "block", "0" if b == 0 else "900", " ".join("%s:%.1f" % (names[i], (tr[i] - t0) / 1000.0) for i in range(6) if tr[i] > 0))
