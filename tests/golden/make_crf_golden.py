"""Generates tests/golden/crf_golden.npz: outputs of the UNMODIFIED reference DenseCRF glue (densecrf.cpp, pairwise.cpp,
labelcompatibility.cpp, unary.cpp compiled by oracle/Makefile against oracle/shim's Eigen stand-in) on small problems.
Run HERE (the build container, where /root/reference exists); the fixture travels to the GPU box.

    python tests/golden/make_crf_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from rovinasemanticsegmentation_b200 import synth  # noqa: E402


def problems():
    """name -> (unary (N, M), [(feats, w)], iterations)"""
    out = {}
    N, M = 3000, 6
    xyz, col = synth.local_map(seed=41, n_points=N)
    lab = (np.floor(xyz[:, 0] * 1.5).astype(int) + np.floor(xyz[:, 1]).astype(int)) % M
    U = synth.unary_from_labels(lab, M, seed=2)
    f6 = oracle.features_xyzrgb(xyz, col, 0.5, 4.0)
    out["node6d"] = (U, [(f6, 10.0)], 10)  # the map worker's kernel (segmenter.cpp:629-643)
    f3 = (xyz * np.float32(2.0)).astype(np.float32)
    f5 = np.concatenate([xyz[:, :2] * np.float32(1.5), col * np.float32(6.0)], axis=1).astype(np.float32)
    out["two_kernels"] = (U, [(f3, 3.0), (f5, 10.0)], 5)
    return out


def main():
    oracle.build(ref=True)
    assert oracle.ref_available(), "needs /root/reference"
    g = {}
    for name, (U, kernels, iters) in problems().items():
        g[name + "_unary"] = U
        for k, (f, w) in enumerate(kernels):
            g["%s_feat%d" % (name, k)] = f
            g["%s_w%d" % (name, k)] = np.float32(w)
        g[name + "_iters"] = np.int32(iters)
        for nt in (0, 1, 2, 3):  # NO_NORMALIZATION, BEFORE, AFTER, SYMMETRIC
            Q, mp = oracle.ref_crf_inference(U, kernels, iters, nt, want_map=True)
            g["%s_Q_norm%d" % (name, nt)] = Q
            g["%s_map_norm%d" % (name, nt)] = mp
    # DenseCRF2D (examples/dense_inference.cpp's model)
    W, H, M = 48, 32, 4
    rgb, _ = synth.frame(23, W, H)
    U = synth.unary_from_labels((np.arange(W * H) // 300) % M, M, 5)
    Q, mp = oracle.ref_crf2d_inference(W, H, U, (3, 3, 3.0), (80, 80, 13, 13, 13, 10.0), rgb, 5)
    g.update(crf2d_unary=U, crf2d_rgb=rgb, crf2d_Q=Q, crf2d_map=mp)
    np.savez_compressed(os.path.join(HERE, "crf_golden.npz"), **g)
    print("wrote crf_golden.npz with", len(g), "arrays")


if __name__ == "__main__":
    main()
