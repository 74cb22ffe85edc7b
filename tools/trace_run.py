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
for mode in (2, 3, 1):
    for b in (0, 1):
        tr = a[mode, b]; t0 = tr[0]
        print("mode", mode, "block", "0" if b == 0 else "300", " ".join("%d:%.1f" % (i, (tr[i] - t0) / 1000.0) for i in range(16) if tr[i] > 0))
tr = a[0, 0]
print("gather first batch (last instance): loads %.2f us, fma %.2f us, red %.2f us" % ((tr[1]-tr[0])/1000.0, (tr[2]-tr[1])/1000.0, (tr[3]-tr[2])/1000.0))
