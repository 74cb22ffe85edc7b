"""TEST INFRASTRUCTURE ONLY: ctypes bindings for the CPU oracle.

* ``liboracle.so``  - plain-C restatement of the reference path (oracle/oracle.c)
* ``_ref/libref_oracle.so`` - the UNMODIFIED reference libforest + permutohedral lattice compiled from
  /root/reference by oracle/Makefile (only buildable where the reference is mounted; the built file travels).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
package.  The product package (rovinasemanticsegmentation_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_PATH = os.path.join(HERE, "_ref", "libref_oracle.so")

WITH_ANY_LABEL, WITH_POSITIVE_LABEL, NO_LABEL = 0, 1, 2


def build(ref=True):
    """Build liboracle.so (always) and _ref/libref_oracle.so (when /root/reference is present)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref and os.path.isdir("/root/reference/third-party"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


_f32p, _i32p, _u8p, _u16p, _i8p = (C.POINTER(t) for t in (C.c_float, C.c_int32, C.c_uint8, C.c_uint16, C.c_int8))


class FeConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("use_color_patch", "use_depth", "use_height", "use_normal", "patch_size",
                                       "patch_size_reduce")]


class Pairwise(C.Structure):
    _fields_ = [("feats", _f32p), ("d", C.c_int), ("potts_w", C.c_float)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build(ref=False)
        L = C.CDLL(LIB_PATH)
        L.orc_forest_read.restype = C.c_void_p
        L.orc_forest_read.argtypes = [C.c_char_p]
        L.orc_lattice_init.restype = C.c_void_p
        for n in ("orc_forest_free", "orc_lattice_free"):
            getattr(L, n).argtypes = [C.c_void_p]
        for n in ("orc_forest_trees", "orc_forest_layers", "orc_lattice_vertices"):
            getattr(L, n).argtypes = [C.c_void_p]
        L.orc_forest_nodes.argtypes = [C.c_void_p, C.c_int]
        L.orc_forest_classes.argtypes = [C.c_void_p, C.c_int]
        _lib = L
    return _lib


def set_threads(n):
    lib().orc_set_threads(int(n))


def default_config(**kw):
    c = FeConfig(1, 1, 1, 1, 77, 11)
    for k, v in kw.items():
        setattr(c, k, int(v))
    return c


# ------------------------------------------------------------------ OpenCV restatements
def bgr2lab(img):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty_like(img)
    lib().orc_bgr2lab_u8(_p(img, C.c_uint8), C.c_int64(img.size // 3), _p(out, C.c_uint8))
    return out


def border_reflect(img, b):
    img = np.ascontiguousarray(img, np.uint8)
    H, W, _ = img.shape
    out = np.empty((H + 2 * b, W + 2 * b, 3), np.uint8)
    lib().orc_border_reflect_u8c3(_p(img, C.c_uint8), W, H, b, _p(out, C.c_uint8))
    return out


def resize_u8c3(win, r):
    win = np.ascontiguousarray(win, np.uint8)
    S = win.shape[0]
    assert win.shape == (S, S, 3)
    out = np.empty((r, r, 3), np.uint8)
    lib().orc_resize_linear_u8c3(_p(win, C.c_uint8), S * 3, S, _p(out, C.c_uint8), r)
    return out


def resize_f32(img, dw, dh):
    img = np.ascontiguousarray(img, np.float32)
    sh, sw, ch = img.shape
    out = np.empty((dh, dw, ch), np.float32)
    lib().orc_resize_linear_f32(_p(img, C.c_float), sw, sh, ch, _p(out, C.c_float), dw, dh)
    return out


# ------------------------------------------------------------------ features
def _calib(Kinv, R, t):
    return (np.ascontiguousarray(Kinv, np.float32).reshape(9), np.ascontiguousarray(R, np.float32).reshape(9),
            np.ascontiguousarray(t, np.float32).reshape(3))


def extract(cfg, stride, rgb, depth, Kinv, R, t, dmin, dmax, extract_type=NO_LABEL, labels=None):
    rgb = np.ascontiguousarray(rgb, np.uint8)
    depth = np.ascontiguousarray(depth, np.uint16)
    H, W = depth.shape
    Kinv, R, t = _calib(Kinv, R, t)
    D = lib().orc_feature_length(C.byref(cfg))
    cap = ((W + stride - 1) // stride) * ((H + stride - 1) // stride)
    feats = np.empty((cap, D), np.float32)
    xs = np.empty(cap, np.int32)
    ys = np.empty(cap, np.int32)
    L = 0
    lab_p = None
    out_lab = None
    if labels is not None:
        labels = np.ascontiguousarray(labels, np.int8)
        L = labels.shape[0]
        lab_p = _p(labels, C.c_int8)
        out_lab = np.empty((cap, L), np.int32)
    n = lib().orc_extract(C.byref(cfg), stride, _p(rgb, C.c_uint8), _p(depth, C.c_uint16), W, H, _p(Kinv, C.c_float),
                          _p(R, C.c_float), _p(t, C.c_float), C.c_float(dmin), C.c_float(dmax), extract_type, lab_p, L,
                          _p(feats, C.c_float), _p(xs, C.c_int32), _p(ys, C.c_int32),
                          _p(out_lab, C.c_int32) if out_lab is not None else None)
    if out_lab is not None:
        return feats[:n], xs[:n], ys[:n], out_lab[:n]
    return feats[:n], xs[:n], ys[:n]


def cloud(depth, Kinv, R, t, dmin, dmax):
    depth = np.ascontiguousarray(depth, np.uint16)
    H, W = depth.shape
    Kinv, R, t = _calib(Kinv, R, t)
    xyz = np.empty((H, W, 3), np.float32)
    lib().orc_cloud(_p(depth, C.c_uint16), W, H, _p(Kinv, C.c_float), _p(R, C.c_float), _p(t, C.c_float),
                    C.c_float(dmin), C.c_float(dmax), _p(xyz, C.c_float))
    return xyz


def normals(xyz, want_dist=False):
    xyz = np.ascontiguousarray(xyz, np.float32)
    H, W, _ = xyz.shape
    nrm = np.empty((H, W, 3), np.float32)
    dist = np.empty((H, W), np.float32) if want_dist else None
    lib().orc_normals(_p(xyz, C.c_float), W, H, _p(nrm, C.c_float), _p(dist, C.c_float) if want_dist else None)
    return (nrm, dist) if want_dist else nrm


# ------------------------------------------------------------------ forest
class Forest:
    def __init__(self, path):
        self.h = lib().orc_forest_read(path.encode())
        if not self.h:
            raise IOError("cannot read forest " + path)
        self.T = lib().orc_forest_trees(self.h)
        self.L = lib().orc_forest_layers(self.h)
        self.classes = [lib().orc_forest_classes(self.h, l) for l in range(self.L)]
        self.sumC = sum(self.classes)
        self.nodes = [lib().orc_forest_nodes(self.h, t) for t in range(self.T)]

    def predict(self, feats):
        feats = np.ascontiguousarray(feats, np.float32)
        n, D = feats.shape
        leaf = np.empty((self.T, n), np.int32)
        post = np.empty((n, self.sumC), np.float32)
        lib().orc_forest_predict(C.c_void_p(self.h), _p(feats, C.c_float), n, D, _p(leaf, C.c_int32),
                                 _p(post, C.c_float))
        return leaf, post

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_forest_free(self.h)
            self.h = None


def segment_frame(cfg, forest, stride, rgb, depth, Kinv, R, t, dmin, dmax, fill=0.0):
    rgb = np.ascontiguousarray(rgb, np.uint8)
    depth = np.ascontiguousarray(depth, np.uint16)
    H, W = depth.shape
    Kinv, R, t = _calib(Kinv, R, t)
    out = np.empty(forest.sumC * H * W, np.float32)
    lib().orc_segment_frame(C.byref(cfg), C.c_void_p(forest.h), stride, _p(rgb, C.c_uint8), _p(depth, C.c_uint16), W, H,
                            _p(Kinv, C.c_float), _p(R, C.c_float), _p(t, C.c_float), C.c_float(dmin), C.c_float(dmax),
                            C.c_float(fill), _p(out, C.c_float))
    return out


# ------------------------------------------------------------------ lattice / CRF
class Lattice:
    """feats: (N, d) array = d x N column-major like the reference's MatrixXf(d, N)."""

    def __init__(self, feats):
        feats = np.ascontiguousarray(feats, np.float32)
        self.N, self.d = feats.shape
        self.h = lib().orc_lattice_init(_p(feats, C.c_float), self.d, self.N)
        self.V = lib().orc_lattice_vertices(C.c_void_p(self.h))

    def get(self):
        off = np.empty((self.N, self.d + 1), np.int32)
        bary = np.empty((self.N, self.d + 1), np.float32)
        lib().orc_lattice_get(C.c_void_p(self.h), _p(off, C.c_int32), _p(bary, C.c_float))
        return off, bary

    def compute(self, x):
        x = np.ascontiguousarray(x, np.float32)
        N, M = x.shape
        out = np.empty_like(x)
        lib().orc_lattice_compute(C.c_void_p(self.h), _p(x, C.c_float), M, _p(out, C.c_float))
        return out

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_lattice_free(self.h)
            self.h = None


def crf_inference(unary, kernels, iters, norm_type=3):
    """unary: (N, M) energies; kernels: list of (feats (N,d), potts_w); norm_type: pairwise.h NormalizationType
    (3 = NORMALIZE_SYMMETRIC, the reference's default).  Returns Q (N, M)."""
    unary = np.ascontiguousarray(unary, np.float32)
    N, M = unary.shape
    keep = [np.ascontiguousarray(f, np.float32) for f, _ in kernels]
    arr = (Pairwise * max(1, len(kernels)))()
    for k, (f, (_, w)) in enumerate(zip(keep, kernels)):
        arr[k].feats = _p(f, C.c_float)
        arr[k].d = f.shape[1]
        arr[k].potts_w = w
    Q = np.empty_like(unary)
    lib().orc_crf_inference_ex(N, M, _p(unary, C.c_float), arr, len(kernels), iters, int(norm_type), _p(Q, C.c_float))
    return Q


def project_zbuffer(xyz, K, R, t, W, H, zmin, zmax):
    """Pinhole z-buffer projection of a cloud (map frame) into a key frame: (H, W) int32 index image, -1 = empty."""
    xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
    K, R, t = _calib(K, R, t)
    out = np.empty((H, W), np.int32)
    lib().orc_project_zbuffer(_p(xyz, C.c_float), xyz.shape[0], _p(K, C.c_float), _p(R, C.c_float), _p(t, C.c_float), W, H,
                              C.c_float(zmin), C.c_float(zmax), _p(out, C.c_int32))
    return out


def unary_accumulate(index_image, posterior, unary):
    index_image = np.ascontiguousarray(index_image, np.int32)
    posterior = np.ascontiguousarray(posterior, np.float32)
    assert unary.dtype == np.float32 and unary.flags.c_contiguous
    lib().orc_unary_accumulate(_p(index_image, C.c_int32), index_image.size, _p(posterior, C.c_float), unary.shape[1],
                               _p(unary, C.c_float))
    return unary


def gated_argmax(Q, unknown):
    Q = np.ascontiguousarray(Q, np.float32)
    N, M = Q.shape
    out = np.empty(N, np.uint8)
    lib().orc_gated_argmax(_p(Q, C.c_float), M, N, unknown, _p(out, C.c_uint8))
    return out


def features_gaussian2d(W, H, sx, sy):
    f = np.empty((W * H, 2), np.float32)
    lib().orc_features_gaussian2d(W, H, C.c_float(sx), C.c_float(sy), _p(f, C.c_float))
    return f


def features_bilateral2d(W, H, sx, sy, sr, sg, sb, im):
    im = np.ascontiguousarray(im, np.uint8)
    f = np.empty((W * H, 5), np.float32)
    lib().orc_features_bilateral2d(W, H, C.c_float(sx), C.c_float(sy), C.c_float(sr), C.c_float(sg), C.c_float(sb),
                                   _p(im, C.c_uint8), _p(f, C.c_float))
    return f


def features_xyzrgb(xyz, rgb, wxyz, wrgb):
    xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
    rgb = np.ascontiguousarray(rgb, np.float32).reshape(-1, 3)
    f = np.empty((xyz.shape[0], 6), np.float32)
    lib().orc_features_xyzrgb(xyz.shape[0], _p(xyz, C.c_float), _p(rgb, C.c_float), C.c_float(wxyz), C.c_float(wrgb),
                              _p(f, C.c_float))
    return f


def keyframe(forest, rgb, depth, Kinv, R, t, sigma_xyz, w_gauss, sigma_px, sigma_rgb, w_bilateral, iters, fill=0.0,
             layers=((8, 7), (9, 8)), stride=2, dmin=0.5, dmax=15.0, cfg=None):
    """The whole keyframe path of rss_segment_keyframe on the CPU (src/segmenter.cpp:349-434 frame worker, then per layer
    a DenseCRF with a Gaussian kernel on the back-projected points and a bilateral kernel, :629-657).
    layers: (label count, "Unknown" label) per layer.  Returns (labels uint8 [L][N], [Q_l (N, M_l)], posteriors)."""
    depth = np.ascontiguousarray(depth, np.uint16)
    H, W = depth.shape
    N = W * H
    post = segment_frame(cfg or default_config(), forest, stride, rgb, depth, Kinv, R, t, dmin, dmax, fill)
    xyz = cloud(depth, Kinv, R, t, dmin, dmax).reshape(-1, 3)
    xyz[np.isnan(xyz[:, 0])] = np.asarray(t, np.float32)
    f3 = (xyz * np.float32(1.0 / sigma_xyz)).astype(np.float32)
    f5 = features_bilateral2d(W, H, sigma_px, sigma_px, sigma_rgb, sigma_rgb, sigma_rgb, rgb)
    labels = np.empty((len(layers), N), np.uint8)
    Qs, off = [], 0
    for l, (M, unk) in enumerate(layers):
        Q = crf_inference(-post[off:off + N * M].reshape(N, M), [(f3, w_gauss), (f5, w_bilateral)], iters)
        labels[l] = gated_argmax(Q, unk)
        Qs.append(Q)
        off += N * M
    return labels, Qs, post


# ------------------------------------------------------------------ the unmodified reference (oracle/_ref)
_ref = None


def ref_available():
    return os.path.exists(REF_PATH)


def ref():
    global _ref
    if _ref is None:
        R = C.CDLL(REF_PATH)
        R.ref_forest_load.restype = C.c_void_p
        R.ref_forest_load.argtypes = [C.c_char_p]
        R.ref_lattice_init.restype = C.c_void_p
        R.ref_forest_free.argtypes = [C.c_void_p]
        R.ref_lattice_free.argtypes = [C.c_void_p]
        R.ref_forest_num_trees.argtypes = [C.c_void_p]
        R.ref_lattice_vertices.argtypes = [C.c_void_p]
        _ref = R
    return _ref


def ref_forest_train(feats, labels, out_path, num_trees=4, max_depth=30, min_split=50, threads=8):
    feats = np.ascontiguousarray(feats, np.float32)
    labels = np.ascontiguousarray(labels, np.int32)
    n, D = feats.shape
    rc = ref().ref_forest_train(_p(feats, C.c_float), n, D, _p(labels, C.c_int32), labels.shape[1], num_trees, max_depth,
                                min_split, threads, out_path.encode())
    if rc:
        raise IOError("reference learner could not write " + out_path)


def ref_forest_train_opts(feats, labels, out_path, num_trees=1, max_depth=30, min_split=50, min_child_split=1,
                          num_features=0, use_bootstrap=True, smoothing=1.0, threads=1):
    """The unmodified reference learner with its options exposed (0 features = autoconf's ceil(sqrt(D)))."""
    feats = np.ascontiguousarray(feats, np.float32)
    labels = np.ascontiguousarray(labels, np.int32)
    n, D = feats.shape
    L = ref()
    L.ref_forest_train_opts.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_float, C.c_int, C.c_char_p]
    rc = L.ref_forest_train_opts(feats.ctypes.data, n, D, labels.ctypes.data, labels.shape[1], num_trees, max_depth, min_split,
                                 min_child_split, num_features, 1 if use_bootstrap else 0, smoothing, threads, out_path.encode())
    if rc:
        raise IOError("reference learner could not write " + out_path)


def ref_forest_update_histograms(in_path, feats, labels, out_path, smoothing=1.0):
    """DecisionTreeLearner::updateMultiHistograms of the unmodified reference on every tree of the forest file in_path."""
    feats = np.ascontiguousarray(feats, np.float32)
    labels = np.ascontiguousarray(labels, np.int32)
    n, D = feats.shape
    L = ref()
    L.ref_forest_update_histograms.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_char_p]
    rc = L.ref_forest_update_histograms(in_path.encode(), feats.ctypes.data, n, D, labels.ctypes.data, labels.shape[1], smoothing,
                                        out_path.encode())
    if rc:
        raise IOError("reference histogram update failed (%d)" % rc)


def read_forest_dat(path):
    """libforest binary model (RandomForest::write, classifier.cpp:210-220) -> list of trees
    {feat, thr, left: arrays; multi: per node list of per-layer float arrays}."""
    b = open(path, "rb").read()
    o = [0]
    def i32():
        v = int(np.frombuffer(b, np.int32, 1, o[0])[0]); o[0] += 4; return v
    def arr(dt, n):
        v = np.frombuffer(b, dt, n, o[0]).copy(); o[0] += 4 * n; return v
    trees = []
    for _ in range(i32()):
        n = i32(); feat = arr(np.int32, n)
        assert i32() == n; thr = arr(np.float32, n)
        assert i32() == n; left = arr(np.int32, n)
        assert i32() == n
        single = [arr(np.float32, i32()) for _ in range(n)]
        assert i32() == n
        multi = [[arr(np.float32, i32()) for _ in range(i32())] for _ in range(n)]
        trees.append({"feat": feat, "thr": thr, "left": left, "single": single, "multi": multi})
    assert o[0] == len(b)
    return trees


class RefForest:
    def __init__(self, path):
        self.h = ref().ref_forest_load(path.encode())
        if not self.h:
            raise IOError(path)
        self.T = ref().ref_forest_num_trees(self.h)

    def predict(self, feats, sumC):
        feats = np.ascontiguousarray(feats, np.float32)
        n, D = feats.shape
        leaf = np.empty((self.T, n), np.int32)
        post = np.empty((n, sumC), np.float32)
        got = ref().ref_forest_predict(C.c_void_p(self.h), _p(feats, C.c_float), n, D, _p(leaf, C.c_int32),
                                       _p(post, C.c_float))
        assert got == sumC, (got, sumC)
        return leaf, post

    def __del__(self):
        if getattr(self, "h", None):
            ref().ref_forest_free(self.h)
            self.h = None


def ref_crf_inference(unary, kernels, iters, norm_type=3, want_map=False):
    """DenseCRF::inference of the UNMODIFIED reference (densecrf.cpp / pairwise.cpp / labelcompatibility.cpp / unary.cpp
    compiled against oracle/shim's Eigen stand-in).  Same arguments as crf_inference."""
    unary = np.ascontiguousarray(unary, np.float32)
    N, M = unary.shape
    keep = [np.ascontiguousarray(f, np.float32) for f, _ in kernels]
    K = len(kernels)
    fp = (C.POINTER(C.c_float) * max(1, K))(*[_p(f, C.c_float) for f in keep])
    d = (C.c_int * max(1, K))(*[f.shape[1] for f in keep])
    w = (C.c_float * max(1, K))(*[float(k[1]) for k in kernels])
    Q = np.empty_like(unary)
    mp = np.empty(N, np.int16) if want_map else None
    ref().ref_crf_inference(N, M, _p(unary, C.c_float), fp, d, w, K, int(norm_type), int(iters), _p(Q, C.c_float),
                            _p(mp, C.c_int16) if want_map else None)
    return (Q, mp) if want_map else Q


def ref_crf_gradient(unary, kernels, iters, gt, robust=0.0, norm_type=3):
    """DenseCRF::gradient of the UNMODIFIED reference (densecrf.cpp:238-297) with the log-likelihood objective
    (objective.cpp:36-52): returns (objective, gradient w.r.t. the Potts weight of every pairwise term)."""
    unary = np.ascontiguousarray(unary, np.float32)
    N, M = unary.shape
    keep = [np.ascontiguousarray(f, np.float32) for f, _ in kernels]
    K = len(kernels)
    fp = (C.POINTER(C.c_float) * max(1, K))(*[_p(f, C.c_float) for f in keep])
    d = (C.c_int * max(1, K))(*[f.shape[1] for f in keep])
    w = (C.c_float * max(1, K))(*[float(k[1]) for k in kernels])
    gt = np.ascontiguousarray(gt, np.int16)
    g = np.zeros(max(1, K), np.float32)
    L = ref()
    L.ref_crf_gradient.restype = C.c_double
    r = L.ref_crf_gradient(N, M, _p(unary, C.c_float), fp, d, w, K, int(norm_type), int(iters), _p(gt, C.c_int16),
                           C.c_float(robust), _p(g, C.c_float))
    return float(r), g[:K]


def ref_crf2d_inference(W, H, unary, gauss, bilateral, im, iters):
    """DenseCRF2D with addPairwiseGaussian(sx, sy, w) + addPairwiseBilateral(sx, sy, sr, sg, sb, im, w) of the reference."""
    unary = np.ascontiguousarray(unary, np.float32)
    im = np.ascontiguousarray(im, np.uint8)
    M = unary.shape[1]
    Q = np.empty_like(unary)
    mp = np.empty(W * H, np.int16)
    f = C.c_float
    ref().ref_crf2d_inference(W, H, M, _p(unary, f), f(gauss[0]), f(gauss[1]), f(gauss[2]), f(bilateral[0]), f(bilateral[1]),
                              f(bilateral[2]), f(bilateral[3]), f(bilateral[4]), _p(im, C.c_uint8), f(bilateral[5]), int(iters),
                              _p(Q, f), _p(mp, C.c_int16))
    return Q, mp


def ref_crf_step_inference(unary, feats, w, steps):
    unary = np.ascontiguousarray(unary, np.float32)
    feats = np.ascontiguousarray(feats, np.float32)
    N, M = unary.shape
    Q = np.empty_like(unary)
    ref().ref_crf_step_inference(N, M, _p(unary, C.c_float), _p(feats, C.c_float), feats.shape[1], C.c_float(w), int(steps),
                                 _p(Q, C.c_float))
    return Q


class RefLattice:
    def __init__(self, feats):
        feats = np.ascontiguousarray(feats, np.float32)
        self.N, self.d = feats.shape
        self.h = ref().ref_lattice_init(_p(feats, C.c_float), self.d, self.N)
        self.V = ref().ref_lattice_vertices(self.h)

    def get(self):
        off = np.empty((self.N, self.d + 1), np.int32)
        bary = np.empty((self.N, self.d + 1), np.float32)
        ref().ref_lattice_get(C.c_void_p(self.h), self.d, self.N, _p(off, C.c_int32), _p(bary, C.c_float))
        return off, bary

    def compute(self, x):
        x = np.ascontiguousarray(x, np.float32)
        N, M = x.shape
        out = np.empty_like(x)
        ref().ref_lattice_compute(C.c_void_p(self.h), _p(x, C.c_float), M, N, _p(out, C.c_float))
        return out

    def __del__(self):
        if getattr(self, "h", None):
            ref().ref_lattice_free(self.h)
            self.h = None
