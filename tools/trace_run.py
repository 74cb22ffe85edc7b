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
a = np.array(buf[:], dtype=np.int64).reshape(4, 32)
for b in (0, 1):
    tr = a[b]; t0 = tr[0]
    print("blur block", b, " ".join("%d:%.1f" % (i, (tr[i] - t0) / 1000.0) for i in range(14) if tr[i] > 0))
