"""Synthetic, seeded workloads for the per-keyframe path.

The reference's dataset (`root_dir` in resources/config.json:2) and model download
(resources/get_rf_model.sh) are unavailable offline, so every test and benchmark runs on these
generated 640x480 RGB-D frames, calibrations, label images and point clouds.  Pure numpy; no
dependency on the CUDA library or on the oracle.
"""
import numpy as np


def calibration(W=640, H=480, tilt_deg=12.0):
    """Pin-hole intrinsics (fx=fy=525*W/640, principal point at the centre) and a camera->base
    extrinsic: camera z (forward) -> base x, camera x (right) -> base -y, camera y (down) -> base -z,
    pitched down by `tilt_deg`, mounted 1.2 m above the floor.  Returns float32 (Kinv, R, t), row-major,
    i.e. the reference's Calibration::_intrinsic_inverse, _extrinsic.linear(), _extrinsic.translation()
    (include/calibration.h:20-22)."""
    f = 525.0 * W / 640.0
    K = np.array([[f, 0, W / 2.0], [0, f, H / 2.0], [0, 0, 1.0]], np.float64)
    Kinv = np.linalg.inv(K)
    base = np.array([[0, 0, 1.0], [-1.0, 0, 0], [0, -1.0, 0]])
    a = np.deg2rad(tilt_deg)
    tilt = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    R = tilt @ base
    t = np.array([0.0, 0.0, 1.2])
    return Kinv.astype(np.float32), R.astype(np.float32), t.astype(np.float32)


def frame(seed=0, W=640, H=480):
    """One RGB-D frame: rgb u8 (H,W,3), depth u16 (H,W) in millimetres."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:H, 0:W].astype(np.float64)
    ph = rng.uniform(0, 2 * np.pi, size=(3, 3))
    fr = rng.uniform(0.004, 0.02, size=(3, 2))
    rgb = np.empty((H, W, 3), np.float64)
    for c in range(3):
        rgb[..., c] = 128 + 70 * np.sin(fr[c, 0] * x + ph[c, 0]) * np.cos(fr[c, 1] * y + ph[c, 1]) + 40 * np.sin(
            0.011 * (x + y) + ph[c, 2])
    rgb += rng.uniform(-8, 8, size=rgb.shape)
    rgb = np.clip(np.rint(rgb), 0, 255).astype(np.uint8)

    p = rng.uniform(0, 2 * np.pi, size=2)
    d = 1000.0 * (1 + 3 * (0.5 + 0.5 * np.sin(0.004 * x * 640.0 / W + p[0]) * np.cos(0.006 * y * 480.0 / H + p[1])))
    # far band (exercises up-scaled patches, S < 11) and one depth discontinuity
    far = (x > 0.80 * W) & (y < 0.45 * H)
    d[far] = 7700 + (15000 - 7700) * ((x[far] - 0.80 * W) / (0.20 * W))
    step = (x > 0.35 * W) & (x < 0.55 * W) & (y > 0.55 * H)
    d[step] += 900.0
    # a band closer than depth_min and one beyond depth_max
    d[(y > 0.93 * H) & (x < 0.25 * W)] = 320.0
    d[(y < 0.04 * H) & (x > 0.9 * W)] = 15800.0
    d += rng.uniform(-2, 2, size=d.shape)
    d = np.clip(np.rint(d), 0, 65535).astype(np.uint16)
    # sensor holes: a few blobs plus sparse single-pixel drop-outs
    for _ in range(10):
        cx, cy, rad = rng.uniform(0, W), rng.uniform(0, H), rng.uniform(3, 14) * W / 640.0
        d[(x - cx) ** 2 + (y - cy) ** 2 < rad * rad] = 0
    d[rng.random((H, W)) < 0.002] = 0
    return rgb, d


def label_thresholds(feats, D_color=363):
    """Quantile thresholds (computed once over the training set) used by labels_from_features."""
    c = D_color // 2 - (D_color // 2) % 3
    ang = feats[:, D_color + 2]
    return {
        "L": np.quantile(feats[:, c], [0.25, 0.5, 0.75]),
        "a": float(np.median(feats[:, c + 1])),
        "h": float(np.quantile(feats[:, D_color + 1], 0.85)),
        "ang": np.quantile(ang[ang >= 0], [0.33, 0.66]) if (ang >= 0).any() else np.array([0.5, 1.1]),
        "d": np.quantile(feats[:, D_color], [0.45, 0.92]),
    }


def labels_from_features(feats, thr, D_color=363):
    """Synthetic training labels for the two layers of resources/config.json:51-78 (material 0..7,
    object 0..8) as deterministic functions of a sample's own feature vector, so that a forest can
    learn them: material from the centre Lab pixel + height, object from normal angle + depth."""
    c = D_color // 2 - (D_color // 2) % 3  # centre patch pixel, channel L
    Lc, ac = feats[:, c], feats[:, c + 1]
    depth, height, ang = feats[:, D_color], feats[:, D_color + 1], feats[:, D_color + 2]
    mat = (np.digitize(Lc, thr["L"]) * 2 + (ac > thr["a"])).astype(np.int32)  # 0..7
    mat = np.where(height > thr["h"], 7 - mat, mat)
    a = np.where(ang < 0, 3, np.digitize(ang, thr["ang"]))  # 0..3 (3 = no normal)
    obj = (a * 2 + (depth > thr["d"][0])).astype(np.int32)  # 0..7
    obj = np.where(depth > thr["d"][1], 8, obj)
    return np.stack([mat, obj], axis=1).astype(np.int32)


def local_map(seed=0, n_points=200_000):
    """Points on the floor and walls of a 12 x 5 x 3 m room with 1 cm jitter; rgb in [0,1] smooth in
    position (the cloud the reference gets from fps_mapper, src/segmenter.cpp:559,629-637)."""
    rng = np.random.default_rng(seed)
    n = n_points
    which = rng.integers(0, 5, size=n)
    u, v = rng.random(n), rng.random(n)
    xyz = np.empty((n, 3), np.float64)
    lx, ly, lz = 12.0, 5.0, 3.0
    for k, (a, b, c) in enumerate([(u * lx, v * ly, 0 * u), (u * lx, 0 * u, v * lz), (u * lx, 0 * u + ly, v * lz),
                                   (0 * u, u * ly, v * lz), (0 * u + lx, u * ly, v * lz)]):
        m = which == k
        xyz[m, 0], xyz[m, 1], xyz[m, 2] = a[m], b[m], c[m]
    xyz += rng.normal(0, 0.01, size=xyz.shape)
    rgb = np.stack([0.5 + 0.4 * np.sin(0.9 * xyz[:, 0] + 0.3), 0.5 + 0.4 * np.cos(1.3 * xyz[:, 1]),
                    0.5 + 0.4 * np.sin(1.7 * xyz[:, 2] + xyz[:, 0])], axis=1)
    rgb = np.clip(rgb + rng.normal(0, 0.02, size=rgb.shape), 0, 1)
    return xyz.astype(np.float32), rgb.astype(np.float32)


def unary_from_labels(labels, M, seed=0, conf=0.6):
    """Noisy -log probability unaries (N, M) for CRF tests (like densecrf examples/common.cpp)."""
    rng = np.random.default_rng(seed)
    N = labels.shape[0]
    p = np.full((N, M), (1 - conf) / (M - 1), np.float64)
    p[np.arange(N), labels] = conf
    p *= rng.uniform(0.6, 1.4, size=p.shape)
    p /= p.sum(axis=1, keepdims=True)
    return (-np.log(p)).astype(np.float32)


def map_keyframe_pose(k, n):
    """Pose (R, t: camera -> map) of key frame k of n inside the room of local_map(): the camera walks along the room's
    long axis at 1.5 m height and looks at alternating walls, slightly downwards."""
    a = (k + 0.5) / n
    t = np.array([1.5 + 9.0 * a, 2.5, 1.5], np.float64)
    yaw = (np.pi / 2 if k % 2 == 0 else -np.pi / 2) + 0.3 * np.sin(3.0 * a)
    pitch = 0.15
    # camera axes in the map frame: z forward, x right, y down
    fwd = np.array([np.cos(yaw) * np.cos(pitch), np.sin(yaw) * np.cos(pitch), -np.sin(pitch)])
    right = np.array([np.sin(yaw), -np.cos(yaw), 0.0])
    down = np.cross(fwd, right)
    R = np.stack([right, down, fwd], axis=1)
    return R.astype(np.float32), t.astype(np.float32)


def intrinsics(W=640, H=480):
    """K (row-major 3x3) of the pinhole camera that calibration() inverts."""
    f = 525.0 * W / 640.0
    return np.array([[f, 0, W / 2.0], [0, f, H / 2.0], [0, 0, 1]], np.float32)
