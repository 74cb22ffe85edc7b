"""B200-native per-keyframe inference path of RovinaSemanticSegmentation.

This package is a thin ctypes binding of ``librss.so`` (hand-written sm_100a CUDA behind the C ABI declared in
``include/rss.h``).  There is NO CPU fallback: importing works anywhere (so that the library can be built and its
exported symbols checked on a machine without a GPU), but creating a :class:`Context` without the compiled
library or without a CUDA device raises.  The package never imports ``oracle`` (the CPU checker).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librss.so")
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "rss.h")
DEFAULT_CONFIG = os.path.join(os.path.dirname(HERE), "resources", "keyframe_config.json")

WITH_ANY_LABEL, WITH_POSITIVE_LABEL, NO_LABEL = 0, 1, 2
NO_NORMALIZATION, NORMALIZE_BEFORE, NORMALIZE_AFTER, NORMALIZE_SYMMETRIC = 0, 1, 2, 3
MAX_LAYERS = 8


class RssError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("rss status %d: %s" % (status, message))
        self.status = status


class Info(C.Structure):
    _fields_ = ([(n, C.c_int) for n in ("feature_color_patch", "feature_depth", "feature_height", "feature_normal",
                                        "patch_size", "patch_size_reduce", "feature_length", "num_trees", "total_nodes",
                                        "total_leaves", "layer_count")] +
                [("class_counts", C.c_int * MAX_LAYERS), ("total_classes", C.c_int),
                 ("unknown_label", C.c_int * MAX_LAYERS), ("use_dense_crf", C.c_int), ("dcrf_iterations", C.c_int),
                 ("rf_prediction_stride", C.c_int), ("dcrf_xyz_kernel", C.c_float), ("dcrf_rgb_kernel", C.c_float),
                 ("dcrf_kernel_weight", C.c_float), ("depth_min", C.c_float), ("depth_max", C.c_float),
                 ("cuda_device", C.c_int), ("sm_count", C.c_int)])


class KeyframeParams(C.Structure):
    _fields_ = [("sigma_xyz", C.c_float), ("w_gauss", C.c_float), ("sigma_px", C.c_float), ("sigma_rgb", C.c_float),
                ("w_bilateral", C.c_float), ("iters", C.c_int), ("fill", C.c_float)]


class Timings(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("h2d_ms", "features_ms", "forest_ms", "upsample_ms", "lattice_ms",
                                         "meanfield_ms", "d2h_ms", "total_ms")]


_lib = None


class TrainParams(C.Structure):  # rss_train_params
    _fields_ = [("num_trees", C.c_int), ("max_depth", C.c_int), ("min_split_examples", C.c_int),
                ("min_child_split_examples", C.c_int), ("num_features", C.c_int), ("use_bootstrap", C.c_int),
                ("num_bootstrap_examples", C.c_int), ("smoothing", C.c_float), ("seed", C.c_uint64)]


class TrainStats(C.Structure):  # rss_train_stats
    _fields_ = [("trees", C.c_int), ("features_per_node", C.c_int), ("bootstrap_examples", C.c_int),
                ("nodes", C.c_longlong), ("levels", C.c_longlong), ("train_ms", C.c_double)]


def load_library():
    """Loads librss.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RssError(-1, "librss.so is not built (run `python -m rovinasemanticsegmentation_b200.build`); "
                               "there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.rss_last_error.restype = C.c_char_p
        L.rss_last_error.argtypes = [C.c_void_p]
        L.rss_status_string.restype = C.c_char_p
        L.rss_kernel_launches.restype = C.c_uint64
        L.rss_kernel_launches.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _calib(Kinv, R, t):
    return (np.ascontiguousarray(Kinv, np.float32).reshape(9), np.ascontiguousarray(R, np.float32).reshape(9),
            np.ascontiguousarray(t, np.float32).reshape(3))


def parse_config(path):
    """rss_parse_config: the config-derived fields of Info, parsed on the host (no CUDA device needed)."""
    lib = load_library()
    info = Info()
    st = lib.rss_parse_config(os.fsencode(path), C.byref(info))
    if st != 0:
        raise RssError(st, (lib.rss_last_error(None) or b"").decode())
    return info


class Context:
    """One GPU context = the reference Segmenter's constructor state (config + forest), see rss_create."""

    def __init__(self, config_path=DEFAULT_CONFIG, forest_path=None, device=0):
        self._lib = load_library()
        h = C.c_void_p()
        st = self._lib.rss_create(config_path.encode(), forest_path.encode() if forest_path else None, int(device),
                                  C.byref(h))
        if st != 0:
            raise RssError(st, self._lib.rss_last_error(None).decode())
        self.h = h
        self.info = Info()
        self._check(self._lib.rss_get_info(self.h, C.byref(self.info)))
        self.D = self.info.feature_length
        self.classes = list(self.info.class_counts[:self.info.layer_count])
        self.sumC = self.info.total_classes

    def _check(self, st):
        if st != 0:
            raise RssError(st, self._lib.rss_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self._lib.rss_destroy(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def kernel_launches(self):
        return int(self._lib.rss_kernel_launches(self.h))

    def timings(self):
        t = Timings()
        self._check(self._lib.rss_get_timings(self.h, C.byref(t)))
        return {n: getattr(t, n) for n, _ in Timings._fields_}

    def upload_frame(self, rgb, depth):
        rgb = np.ascontiguousarray(rgb, np.uint8)
        depth = np.ascontiguousarray(depth, np.uint16)
        H, W = depth.shape
        self._check(self._lib.rss_upload_frame(self.h, _ptr(rgb, C.c_uint8), _ptr(depth, C.c_uint16), W, H))

    def profile_enable(self, on=True):
        self._check(self._lib.rss_profile_enable(self.h, 1 if on else 0))

    def profile_report(self):
        """{kernel name: (total device ms, launches)} accumulated since profile_enable."""
        out = {}
        name = C.create_string_buffer(128)
        ms = C.c_double(0)
        n = C.c_uint64(0)
        i = 0
        while self._lib.rss_profile_get(self.h, i, name, 128, C.byref(ms), C.byref(n)) == 0:
            out[name.value.decode()] = (ms.value, n.value)
            i += 1
        return out

    # ---- FeatureExtractor::extract
    def extract_features(self, rgb, depth, Kinv, R, t, stride, dmin, dmax, extract_type=NO_LABEL, labels=None,
                         want_feats=True):
        rgb = np.ascontiguousarray(rgb, np.uint8)
        depth = np.ascontiguousarray(depth, np.uint16)
        H, W = depth.shape
        Kinv, R, t = _calib(Kinv, R, t)
        cap = (-(-W // stride)) * (-(-H // stride))
        feats = np.empty((cap, self.D), np.float32) if want_feats else None
        xs = np.empty(cap, np.int32)
        ys = np.empty(cap, np.int32)
        nl, lab, out_lab = 0, None, None
        if labels is not None:
            lab = np.ascontiguousarray(labels, np.int8)
            nl = lab.shape[0]
            out_lab = np.empty((cap, nl), np.int32)
        n = C.c_int(0)
        self._check(self._lib.rss_extract_features(
            self.h, _ptr(rgb, C.c_uint8), _ptr(depth, C.c_uint16), W, H, int(stride), _ptr(Kinv, C.c_float),
            _ptr(R, C.c_float), _ptr(t, C.c_float), C.c_float(dmin), C.c_float(dmax), int(extract_type),
            _ptr(lab, C.c_int8), nl, _ptr(feats, C.c_float), _ptr(xs, C.c_int32), _ptr(ys, C.c_int32),
            _ptr(out_lab, C.c_int32), C.byref(n)))
        n = n.value
        out = [feats[:n] if want_feats else None, xs[:n], ys[:n]]
        if out_lab is not None:
            out.append(out_lab[:n])
        return tuple(out)

    def frame_intermediates(self, W, H, lab=True, xyz=True, normals=True):
        P = self.info.patch_size
        a = np.empty((H + 2 * P, W + 2 * P, 3), np.uint8) if lab else None
        b = np.empty((H, W, 3), np.float32) if xyz else None
        c = np.empty((H, W, 3), np.float32) if normals else None
        self._check(self._lib.rss_frame_intermediates(self.h, _ptr(a, C.c_uint8), _ptr(b, C.c_float), _ptr(c, C.c_float)))
        return a, b, c

    # ---- RandomForest::multiClassLogPosterior
    def load_forest(self, path):
        """RandomForest::read (classifier.cpp:222-235) into this context; refreshes the model-derived info."""
        self._check(self._lib.rss_load_forest(self.h, os.fsencode(path)))
        self._check(self._lib.rss_get_info(self.h, C.byref(self.info)))
        self.classes = list(self.info.class_counts[:self.info.layer_count])
        self.sumC = self.info.total_classes

    def forest_predict(self, feats=None, n=None):
        if feats is not None:
            feats = np.ascontiguousarray(feats, np.float32)
            n = feats.shape[0]
            assert feats.shape[1] == self.D
        T = self.info.num_trees
        leaf = np.empty((T, n), np.int32)
        post = np.empty((n, self.sumC), np.float32)
        self._check(self._lib.rss_forest_predict(self.h, _ptr(feats, C.c_float), int(n), _ptr(leaf, C.c_int32),
                                                 _ptr(post, C.c_float)))
        return leaf, post

    # ---- Segmenter::processFramesFromQueueInternalRF body
    def segment_frame(self, rgb, depth, Kinv, R, t, fill=0.0, want_host=True):
        rgb = np.ascontiguousarray(rgb, np.uint8)
        depth = np.ascontiguousarray(depth, np.uint16)
        H, W = depth.shape
        Kinv, R, t = _calib(Kinv, R, t)
        out = np.empty(self.sumC * H * W, np.float32) if want_host else None
        self._check(self._lib.rss_segment_frame(self.h, _ptr(rgb, C.c_uint8), _ptr(depth, C.c_uint16), W, H,
                                                _ptr(Kinv, C.c_float), _ptr(R, C.c_float), _ptr(t, C.c_float),
                                                C.c_float(fill), _ptr(out, C.c_float)))
        return out

    # ---- GPU forest training (rss_forest_train)
    def forest_train(self, feats, labels, class_counts, out_path, num_trees=4, max_depth=30, min_split_examples=50,
                     min_child_split_examples=1, num_features=0, use_bootstrap=True, num_bootstrap_examples=0,
                     smoothing=1.0, seed=1):
        """feats [n][D] f32, labels [n][L] i32 -> libforest .dat at out_path; returns TrainStats."""
        feats = np.ascontiguousarray(feats, np.float32)
        labels = np.ascontiguousarray(labels, np.int32)
        if labels.ndim == 1:
            labels = labels[:, None]
        n, D = feats.shape
        L = labels.shape[1]
        cc = np.ascontiguousarray(class_counts, np.int32)
        prm = TrainParams(num_trees, max_depth, min_split_examples, min_child_split_examples, num_features,
                          1 if use_bootstrap else 0, num_bootstrap_examples, smoothing, seed)
        stats = TrainStats()
        self._check(self._lib.rss_forest_train(self.h, _ptr(feats, C.c_float), n, D, _ptr(labels, C.c_int32), L,
                                               _ptr(cc, C.c_int), C.byref(prm), os.fsencode(out_path), C.byref(stats)))
        return stats

    # ---- srv/SingleFrameSegmentation.srv payloads (rss_service_single_frame)
    def service_single_frame(self, rgb, depth3d, Kinv, R, t):
        """rgb [H][W][3] u8 and the node's rectified cloud [H][W][3] f32 (src/segmenter.cpp:463-488) -> label_distribution."""
        rgb = np.ascontiguousarray(rgb, np.uint8)
        depth3d = np.ascontiguousarray(depth3d, np.float32)
        H, W = depth3d.shape[:2]
        Kinv, R, t = _calib(Kinv, R, t)
        out = np.empty(self.sumC * H * W, np.float32)
        self._check(self._lib.rss_service_single_frame(self.h, _ptr(rgb, C.c_uint8), _ptr(depth3d, C.c_float), W, H,
                                                       _ptr(Kinv, C.c_float), _ptr(R, C.c_float), _ptr(t, C.c_float),
                                                       _ptr(out, C.c_float)))
        return out

    def segment_keyframe(self, rgb, depth, Kinv, R, t, params, W=None, H=None, want_Q=False, want_labels=True):
        """rgb/depth may be None to reuse the frame already resident on the device (then pass W, H)."""
        if rgb is not None:
            rgb = np.ascontiguousarray(rgb, np.uint8)
            depth = np.ascontiguousarray(depth, np.uint16)
            H, W = depth.shape
        Kinv, R, t = _calib(Kinv, R, t)
        L = self.info.layer_count
        labels = np.empty((L, H * W), np.uint8) if want_labels else None
        Q = np.empty(self.sumC * H * W, np.float32) if want_Q else None
        self._check(self._lib.rss_segment_keyframe(self.h, _ptr(rgb, C.c_uint8), _ptr(depth, C.c_uint16), W, H,
                                                   _ptr(Kinv, C.c_float), _ptr(R, C.c_float), _ptr(t, C.c_float),
                                                   C.byref(params), _ptr(labels, C.c_uint8), _ptr(Q, C.c_float)))
        return (labels, Q) if want_Q else labels

    def posteriors_keep(self, slot):
        """Keep the posteriors of the last segment_frame in device slot `slot` for the map worker (rss_posteriors_keep)."""
        self._check(self._lib.rss_posteriors_keep(self.h, int(slot)))

    def keyframe_graph(self, enable=True):
        """Turn the CUDA-graph replay of segment_keyframe's device part on / off for this context (rss_keyframe_graph)."""
        self._check(self._lib.rss_keyframe_graph(self.h, int(bool(enable))))

    def keyframe_lattice_info(self, k):
        """(feature dimension, vertex count) of lattice k of the last segment_keyframe call."""
        d, v = C.c_int(0), C.c_int(0)
        self._check(self._lib.rss_keyframe_lattice_info(self.h, int(k), C.byref(d), C.byref(v)))
        return d.value, v.value

    def crf(self, N, M):
        return DenseCRF(self, N, M)


class DenseCRF:
    """DenseCRF (third-party/densecrf/include/densecrf.h) on the device.  M: int or list of per-layer label counts.
    Matrices use the reference's column-major convention: a numpy array of shape (N, M)."""

    def __init__(self, ctx, N, M):
        self.ctx = ctx
        self._lib = ctx._lib
        self.N = int(N)
        self.M = [int(M)] if np.isscalar(M) else [int(m) for m in M]
        h = C.c_void_p()
        arr = (C.c_int * len(self.M))(*self.M)
        ctx._check(self._lib.rss_crf_create_layers(ctx.h, self.N, len(self.M), arr, C.byref(h)))
        self.h = h
        self.n_kernels = 0

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self._lib.rss_crf_destroy(self.h)
        self.h = None

    __del__ = close

    def set_unary(self, U, layer=0):
        U = np.ascontiguousarray(U, np.float32)
        assert U.shape == (self.N, self.M[layer])
        self.ctx._check(self._lib.rss_crf_set_unary(self.h, layer, _ptr(U, C.c_float)))

    def add_pairwise(self, feats, potts_w, norm_type=NORMALIZE_SYMMETRIC):
        feats = np.ascontiguousarray(feats, np.float32)
        assert feats.shape[0] == self.N
        self.ctx._check(self._lib.rss_crf_add_pairwise(self.h, _ptr(feats, C.c_float), feats.shape[1],
                                                       C.c_float(potts_w), int(norm_type)))
        self.n_kernels += 1

    def add_pairwise_gaussian(self, W, H, sx, sy, potts_w):
        self.ctx._check(self._lib.rss_crf_add_pairwise_gaussian(self.h, W, H, C.c_float(sx), C.c_float(sy),
                                                                C.c_float(potts_w)))
        self.n_kernels += 1

    def add_pairwise_bilateral(self, W, H, sx, sy, sr, sg, sb, im, potts_w):
        im = np.ascontiguousarray(im, np.uint8)
        self.ctx._check(self._lib.rss_crf_add_pairwise_bilateral(self.h, W, H, C.c_float(sx), C.c_float(sy),
                                                                 C.c_float(sr), C.c_float(sg), C.c_float(sb),
                                                                 _ptr(im, C.c_uint8), C.c_float(potts_w)))
        self.n_kernels += 1

    def add_pairwise_xyzrgb(self, xyz, rgb, wxyz, wrgb, potts_w):
        xyz = np.ascontiguousarray(xyz, np.float32)
        rgb = np.ascontiguousarray(rgb, np.float32)
        self.ctx._check(self._lib.rss_crf_add_pairwise_xyzrgb(self.h, _ptr(xyz, C.c_float), _ptr(rgb, C.c_float),
                                                              C.c_float(wxyz), C.c_float(wrgb), C.c_float(potts_w)))
        self.n_kernels += 1

    def gradient(self, iters, gt, layer=0, robust=0.0):
        """DenseCRF::gradient with the log-likelihood objective: (objective, d objective / d Potts weight of every term)."""
        gt = np.ascontiguousarray(gt, np.int32)
        obj = C.c_double(0.0)
        g = np.zeros(max(1, self.n_kernels), np.float32)
        self.ctx._check(self._lib.rss_crf_gradient(self.h, int(layer), int(iters), _ptr(gt, C.c_int32), C.c_float(robust),
                                                   C.byref(obj), _ptr(g, C.c_float)))
        return obj.value, g[:self.n_kernels]

    def path(self):
        """(fused, sorted) of the next inference: rss_crf_path."""
        f, s = C.c_int(0), C.c_int(0)
        self.ctx._check(self._lib.rss_crf_path(self.h, C.byref(f), C.byref(s)))
        return bool(f.value), bool(s.value)

    def lattice_size(self, k=0):
        v = C.c_int(0)
        self.ctx._check(self._lib.rss_crf_lattice_size(self.h, k, C.byref(v)))
        return v.value

    def filter(self, x, k=0):
        x = np.ascontiguousarray(x, np.float32)
        assert x.shape == (self.N, sum(self.M))
        out = np.empty_like(x)
        self.ctx._check(self._lib.rss_crf_filter(self.h, k, _ptr(x, C.c_float), _ptr(out, C.c_float)))
        return out

    def inference(self, iters, layer=-1, unknown=None, want_Q=True, want_labels=False):
        nl = len(self.M)
        if layer < 0:
            Q = [np.empty((self.N, m), np.float32) for m in self.M] if want_Q else None
            Qbuf = np.empty(self.N * sum(self.M), np.float32) if want_Q else None
            lab = np.empty((nl, self.N), np.uint8) if want_labels else None
        else:
            Qbuf = np.empty((self.N, self.M[layer]), np.float32) if want_Q else None
            lab = np.empty(self.N, np.uint8) if want_labels else None
        unk = None
        if unknown is not None:
            u = [unknown] if np.isscalar(unknown) else list(unknown)
            unk = (C.c_int * len(u))(*u)
        self.ctx._check(self._lib.rss_crf_inference(self.h, int(layer), int(iters), _ptr(Qbuf, C.c_float),
                                                    _ptr(lab, C.c_uint8), unk))
        if want_Q and layer < 0:
            o = 0
            for l, m in enumerate(self.M):
                Q[l] = Qbuf[o:o + self.N * m].reshape(self.N, m)
                o += self.N * m
            Qbuf = Q[0] if nl == 1 else Q
        if want_Q and want_labels:
            return Qbuf, lab
        return Qbuf if want_Q else lab

    def set_cloud(self, xyz, rgb):
        """Upload the local map's points once; they stay resident (rss_crf_set_cloud)."""
        xyz = np.ascontiguousarray(xyz, np.float32)
        rgb = np.ascontiguousarray(rgb, np.float32)
        self.ctx._check(self._lib.rss_crf_set_cloud(self.h, _ptr(xyz, C.c_float), _ptr(rgb, C.c_float)))

    def project_accumulate(self, W, H, K, R, t, zmin, zmax, slot=-1, want_index=False):
        """Device projector + unary accumulation from device-resident posteriors (rss_crf_project_accumulate)."""
        K, R, t = _calib(K, R, t)
        idx = np.empty((H, W), np.int32) if want_index else None
        self.ctx._check(self._lib.rss_crf_project_accumulate(self.h, self.ctx.h, int(slot), W, H, _ptr(K, C.c_float),
                                                             _ptr(R, C.c_float), _ptr(t, C.c_float), C.c_float(zmin),
                                                             C.c_float(zmax), _ptr(idx, C.c_int32)))
        return idx

    def add_pairwise_cloud(self, wxyz, wrgb, potts_w):
        self.ctx._check(self._lib.rss_crf_add_pairwise_cloud(self.h, C.c_float(wxyz), C.c_float(wrgb), C.c_float(potts_w)))
        self.n_kernels += 1

    def clear_pairwise(self):
        self.ctx._check(self._lib.rss_crf_clear_pairwise(self.h))
        self.n_kernels = 0

    def unary_reset(self):
        self.ctx._check(self._lib.rss_crf_unary_reset(self.h))

    def unary_accumulate(self, index_image, posteriors=None):
        idx = np.ascontiguousarray(index_image, np.int32).reshape(-1)
        post = np.ascontiguousarray(posteriors, np.float32) if posteriors is not None else None
        self.ctx._check(self._lib.rss_crf_unary_accumulate(self.h, self.ctx.h, _ptr(idx, C.c_int32), idx.size,
                                                           _ptr(post, C.c_float)))
