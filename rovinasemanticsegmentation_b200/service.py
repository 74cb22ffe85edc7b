"""The node's external single-frame service, served by the CUDA frame path.

Reference: srv/SingleFrameSegmentation.srv, the client in src/segmenter.cpp:446-514 (external_semantics switch,
:101-103) and the stub server scripts/single_frame_segmentation_server.py:12-52, whose request / response handling this
module mirrors: request = sensor_msgs/Image rgb ("rgb8") + sensor_msgs/Image depth ("32FC3": the rectified world-frame
cloud, NaN where the raw depth is invalid), response = float32[] label_distribution in [layer][y][x][class] order.

ROS (rospy, cv_bridge) is not part of this image: `segment_frame` works on any object with the sensor_msgs/Image fields
(height, width, encoding, step, data), `main()` wires it to rospy when rospy can be imported.  The other three services
of the node (srv/IdsSrv.srv, LocalMapSegmentationSrv.srv, SegmentationInformationSrv.srv; src/segmenter.cpp:722-792)
are answered from a result store filled by the map worker."""
import json
import threading

import numpy as np

from . import Context, DEFAULT_CONFIG


class ImageMsg:
    """sensor_msgs/Image stand-in (same field names)."""

    def __init__(self, height, width, encoding, step, data):
        self.height, self.width, self.encoding, self.step, self.data = height, width, encoding, step, data


def image_to_array(msg):
    """cv_bridge.imgmsg_to_cv2 for the two encodings of the service (single_frame_segmentation_server.py:13-14)."""
    enc = msg.encoding.lower()
    if enc == "rgb8":
        dt, ch = np.uint8, 3
    elif enc == "32fc3":
        dt, ch = np.float32, 3
    else:
        raise ValueError("unsupported encoding %r (the node sends rgb8 and 32FC3)" % msg.encoding)
    raw = np.frombuffer(bytes(msg.data) if not isinstance(msg.data, (bytes, bytearray, memoryview, np.ndarray)) else msg.data,
                        np.uint8)
    row = msg.width * ch * np.dtype(dt).itemsize
    if msg.step < row or raw.size < msg.step * msg.height:
        raise ValueError("image message shorter than height * step")
    rows = raw[:msg.step * msg.height].reshape(msg.height, msg.step)[:, :row]
    return np.ascontiguousarray(rows).view(dt).reshape(msg.height, msg.width, ch)


def rectified_cloud(depth_mm, Kinv, R, t):
    """The "depth" image the node sends (src/segmenter.cpp:463-488), restated in float32 numpy for tests and examples:
    mat = (d x, d y, d) with d = depth / 1000.f, NaN where d < 0.5 or d > 15; rect = (R * Kinv) * mat + t."""
    depth_mm = np.asarray(depth_mm, np.uint16)
    H, W = depth_mm.shape
    d = depth_mm.astype(np.float32) / np.float32(1000.0)
    bad = (d < np.float32(0.5)) | (d > np.float32(15.0))
    xs = np.arange(W, dtype=np.float32)[None, :].repeat(H, 0)
    ys = np.arange(H, dtype=np.float32)[:, None].repeat(W, 1)
    mat = np.stack([d * xs, d * ys, d], 0).reshape(3, -1).astype(np.float32)
    M = (np.asarray(R, np.float32).reshape(3, 3) @ np.asarray(Kinv, np.float32).reshape(3, 3)).astype(np.float32)
    rect = (M @ mat + np.asarray(t, np.float32).reshape(3, 1)).astype(np.float32)
    rect[:, bad.reshape(-1)] = np.nan
    return np.ascontiguousarray(rect.T.reshape(H, W, 3))


class SingleFrameSegmentationResponse:
    def __init__(self, label_distribution):
        self.label_distribution = label_distribution


class SegmentationServer:
    """Owns one CUDA context and answers the node's four services."""

    def __init__(self, config_path=DEFAULT_CONFIG, forest_path=None, Kinv=None, R=None, t=None, device=0):
        self.ctx = Context(config_path, forest_path, device)
        self.Kinv, self.R, self.t = Kinv, R, t
        # layer information, parsed like single_frame_segmentation_server.py:62-69 / src/segmenter.cpp:70-99
        with open(config_path) as f:
            cfg = json.load(f)
        self.layer_names, self.class_counts, self.class_names, self.class_colors = [], [], [], []
        for coding in cfg["color_codings"]:  # src/segmenter.cpp:73-98: classes with label >= 0, in file order
            self.layer_names.append(coding["name"])
            n = 0
            for entry in coding["coding"]:
                if int(entry["label"]) >= 0:
                    self.class_names.append(entry["name"])
                    self.class_colors.extend(int(c) for c in entry["color"][:3])
                    n += 1
            self.class_counts.append(n)
        self._maps = {}  # local_map_id -> [layer][point] uint8 labels
        self._lock = threading.Lock()

    # ---- /semantic_segmentation/SingleFrameSegmentation
    def segment_frame(self, req):
        rgb = image_to_array(req.rgb)
        cloud = image_to_array(req.depth)
        if rgb.shape[:2] != cloud.shape[:2]:
            raise ValueError("rgb and depth sizes differ")
        if self.Kinv is None:
            raise RuntimeError("the server needs the camera calibration the node rectifies with")
        out = self.ctx.service_single_frame(rgb, cloud, self.Kinv, self.R, self.t)
        return SingleFrameSegmentationResponse(out)

    # ---- map-side services (src/segmenter.cpp:722-792)
    def store_map_result(self, local_map_id, labels):
        with self._lock:
            self._maps[int(local_map_id)] = np.ascontiguousarray(labels, np.uint8)

    def stored_semantics_ids(self):  # IdsSrv
        with self._lock:
            return sorted(self._maps)

    def local_map_segmentation(self, local_map_id, segmentation_layers):  # LocalMapSegmentationSrv
        idx = [self.layer_names.index(l) for l in segmentation_layers if l in self.layer_names]
        if len(idx) != len(segmentation_layers):
            return None  # the node returns false (:748-750)
        with self._lock:
            labels = self._maps.get(int(local_map_id))
        if labels is None:
            return None
        return int(local_map_id), np.concatenate([labels[l] for l in idx]) if idx else np.zeros(0, np.uint8)

    def segmentation_information(self):  # SegmentationInformationSrv
        return {"layer_names": list(self.layer_names), "class_counts": list(self.class_counts),
                "class_names": list(self.class_names), "class_colors": list(self.class_colors)}

    def close(self):
        self.ctx.close()


def main():
    try:
        import rospy
        from semantic_segmentation.srv import SingleFrameSegmentation, SingleFrameSegmentationResponse as RosResponse
    except ImportError as e:  # this image has no ROS
        raise SystemExit("rospy / the semantic_segmentation package are not importable here (%s); "
                         "use SegmentationServer.segment_frame directly" % e)
    rospy.init_node("single_frame_segmentation_server")
    ns = rospy.get_name() + "/"
    calib = rospy.get_param(ns + "calibration")  # {"Kinv": [9], "R": [9], "t": [3]}
    srv = SegmentationServer(rospy.get_param(ns + "config_file"), rospy.get_param(ns + "forest_file"),
                             calib["Kinv"], calib["R"], calib["t"], int(rospy.get_param(ns + "cuda_device", 0)))
    rospy.Service("/semantic_segmentation/SingleFrameSegmentation", SingleFrameSegmentation,
                  lambda req: RosResponse(srv.segment_frame(req).label_distribution))
    print("SingleFrameSegmentation server ready!")
    rospy.spin()


if __name__ == "__main__":
    main()
