import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import oracle, rovinasemanticsegmentation_b200 as rss
from test_gpu_edges import _problem
from conftest import CONFIG, FOREST
oracle.set_threads(8)
ctx = rss.Context(CONFIG, FOREST, 0)
for (N, M, d) in ((5003, 6, 3), (5003, 6, 2), (5120, 6, 3), (5003, 17, 3), (5003, 6, 5)):
    f, U = _problem(N, M, d, d)
    Q0 = oracle.crf_inference(U, [(f, 4.0)], 6)
    crf = ctx.crf(N, M); crf.set_unary(U); crf.add_pairwise(f, 4.0)
    Q1 = crf.inference(6)
    err = np.abs(Q0 - Q1).max(1)
    bad = np.nonzero(err > 1e-4)[0]
    print(N, M, d, "V", crf.lattice_size(0), "max err", err.max(), "bad points", len(bad), bad[:20], bad[-5:] if len(bad) else "")
    # filter only
    x = np.random.default_rng(0).random((N, M), dtype=np.float32)
    lat = oracle.Lattice(f); ref = lat.compute(x); out = crf.filter(x)
    print("   filter max err", np.abs(ref - out).max(), "V ref", lat.V)
    # step-wise (generic path) result
    crf.ctx._check(crf._lib.rss_crf_start_inference(crf.h)); crf.ctx._check(crf._lib.rss_crf_step_inference(crf.h, 6))
    import ctypes as C
    Q2 = np.empty((N, M), np.float32); crf.ctx._check(crf._lib.rss_crf_current(crf.h, 0, Q2.ctypes.data_as(C.POINTER(C.c_float)), None, None))
    print("   generic path max err", np.abs(Q0 - Q2).max())
    crf.close()
