"""keypoint_tracker::track mirror (zenslam_core/source/tracking/keypoint_tracker.cpp:41-105): the per-frame caller of
the hot path, with the reference's own bookkeeping -- index-keyed keypoint maps, occupancy-aware detection, stereo
tracking of the keypoints the other camera does not have yet.

Every computation goes through the same seams the C++ adapter binds (keypoint_detector, pyr_lk); this file is host
glue only.  What the reference does around it and this backend leaves on the CPU is injected, not re-implemented:
  * landmark projection for the initial flow (keypoint_tracker.cpp:361-373) -> `predicted_points` callable,
  * assign_landmark_indices (:55,71) -> `assign_landmarks` callable (zenslam_b200.matching.assign_landmark_indices fits),
  * filter_epipolar's cv::findFundamentalMat RANSAC (:293-341) -> `epipolar_filter` callable; without one the
    option tracking.filter_epipolar must be off (the RANSAC gate is out of scope).
"""
from __future__ import annotations

from dataclasses import dataclass, field

from .detection import keypoint_detector_grid, keypoint_detector_parallel, keypoint_detector_simple
from .options import slam_options
from .tracking import pyr_lk, track_keypoints


class keypoint_map(dict):
    """zenslam::map<keypoint> : std::map<size_t, keypoint> (types/map.h:24-100): ordered by index, add() keeps an
    existing entry unless overwrite is set."""

    def add(self, items, overwrite: bool = False):
        for kp in (items if isinstance(items, (list, tuple)) else [items]):
            if overwrite or kp.index not in self:
                self[kp.index] = kp

    def values_sorted(self) -> list:
        return [self[k] for k in sorted(self)]

    def values_unmatched(self, other) -> list:
        return [self[k] for k in sorted(self) if k not in other]

    def values_matched(self, other) -> list:
        return [self[k] for k in sorted(self) if k in other]


@dataclass
class stereo_frame:
    """the pieces of frame::processed / frame::estimated that track() reads: level-0 images of the two pyramids
    (the device rebuilds cv::buildOpticalFlowPyramid from them), the undistorted images and the keypoint maps"""
    undistorted: tuple
    keypoints: tuple = field(default_factory=lambda: (keypoint_map(), keypoint_map()))

    @property
    def pyramids(self):
        return self.undistorted


class keypoint_tracker:
    def __init__(self, options: slam_options, ctx, backend: pyr_lk, detector=None, predicted_points=None,
                 assign_landmarks=None, epipolar_filter=None):
        self._options, self._pyr_lk = options, backend
        det = options.detection
        if detector is not None:
            self._detector = detector
        elif det.algorithm == "SIMPLE":           # keypoint_tracker.cpp:27-38
            self._detector = keypoint_detector_simple(det, ctx)
        elif det.algorithm == "GRID":
            self._detector = keypoint_detector_grid(det, ctx)
        else:
            self._detector = keypoint_detector_parallel(det, ctx)
        self._predicted_points, self._assign, self._epipolar = predicted_points, assign_landmarks, epipolar_filter
        if options.tracking.filter_epipolar and epipolar_filter is None:
            raise NotImplementedError("tracking.filter_epipolar needs an epipolar_filter callable: the reference's gate is "
                                      "cv::findFundamentalMat RANSAC on the CPU (keypoint_tracker.cpp:301-308)")

    def _track(self, img_0, img_1, keypoints: list, camera_index=None) -> list:
        predicted = None
        if camera_index is not None and self._predicted_points is not None:
            predicted = self._predicted_points(keypoints, camera_index)
        return track_keypoints(self._pyr_lk, img_0, img_1, keypoints, self._options.tracking, predicted)

    def track(self, frame_0: stereo_frame, frame_1: stereo_frame):
        """-> (keypoints_0, keypoints_1) of frame_1 (keypoint_tracker.cpp:41-105)"""
        kp0, kp1 = keypoint_map(), keypoint_map()
        # temporal tracks of both cameras (:47-51); the overload with a pose passes OPTFLOW_USE_INITIAL_FLOW with the
        # keypoint's own position where no landmark is known, which is what the plain call starts from too
        kp0.add(self._track(frame_0.pyramids[0], frame_1.pyramids[0], frame_0.keypoints[0].values_sorted(), 0))
        kp1.add(self._track(frame_0.pyramids[1], frame_1.pyramids[1], frame_0.keypoints[1].values_sorted(), 1))
        # new keypoints in the cells the tracked ones leave free (:53-57)
        detected_0 = self._detector.detect_keypoints(frame_1.undistorted[0], kp0)
        if self._assign is not None:
            self._assign(detected_0)
        kp0.add(detected_0)
        # left keypoints the right camera does not have yet: stereo track L -> R (:59-67)
        kp1.add(self._track(frame_1.pyramids[0], frame_1.pyramids[1], kp0.values_unmatched(kp1)))
        detected_1 = self._detector.detect_keypoints(frame_1.undistorted[1], kp1)
        if self._assign is not None:
            self._assign(detected_1)
        kp1.add(detected_1)
        # and the other way round (:73-83)
        kp0.add(self._track(frame_1.pyramids[1], frame_1.pyramids[0], kp1.values_unmatched(kp0)))
        if self._options.tracking.filter_epipolar:
            m0, m1 = kp0.values_matched(kp1), kp1.values_matched(kp0)
            keep = self._epipolar(m0, m1)
            f0, f1 = keypoint_map(), keypoint_map()
            for a, b, k in zip(m0, m1, keep):
                if k:
                    f0.add(a); f1.add(b)
            return f0, f1
        return kp0, kp1


class device_keypoint_tracker:
    """zs_tracker: the same track() flow with the keypoint maps, the previous pyramids and keypoint::index_next kept on the
    device -- one C-ABI call per stereo frame instead of seven -- for `sequences` independent stereo sequences in lock-step
    (1 = the reference's own use).  GRID / FAST / ORB, filter_epipolar off (see the module docstring for what stays on
    the host)."""

    def __init__(self, options: slam_options, ctx, width: int, height: int, first_index: int = 0, capacity: int = 0,
                 sequences: int = 1, landmark_capacity: int = 0):
        import ctypes as C

        from ._lib import TrackerOptions, check, lib
        det, trk = options.detection, options.tracking
        if det.algorithm not in ("GRID", "PARALLEL_GRID") or det.feature_detector != "FAST" or det.descriptor != "ORB":
            raise NotImplementedError("the device tracker implements algorithm GRID / PARALLEL_GRID with feature FAST and descriptor ORB")
        if trk.filter_epipolar:
            raise NotImplementedError("tracking.filter_epipolar is a CPU RANSAC in the reference; switch it off or filter the result")
        self._ctx, self.width, self.height, self.sequences = ctx, width, height, sequences
        o = TrackerOptions(width, height, det.cell_size[0], det.cell_size[1], det.fast_threshold, trk.klt_window_size[0],
                           trk.klt_window_size[1], trk.klt_max_level, trk.klt_threshold, capacity, first_index, sequences,
                           1 if det.algorithm == "PARALLEL_GRID" else 0, landmark_capacity, trk.landmark_match_radius,
                           trk.landmark_match_distance)
        h = C.c_void_p()
        check(lib().zs_tracker_create(ctx._h, C.byref(o), C.byref(h)))
        self._h = h
        self.cap = lib().zs_tracker_capacity(h)
        self.next_index = [first_index] * sequences

    def close(self):
        if getattr(self, "_h", None):
            from ._lib import lib
            lib().zs_tracker_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_predictions(self, camera: int, predictions: dict, sequence: int = 0):
        """{keypoint index: (x, y)} -- initial flow of the next temporal track for those keypoints (landmark projections,
        keypoint_tracker.cpp:361-373); consumed by the next track()"""
        import ctypes as C

        import numpy as np

        from ._lib import check, lib
        keys = sorted(predictions)
        idx = np.array(keys, np.int32); xy = np.array([predictions[k] for k in keys], np.float32).reshape(-1, 2)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        check(lib().zs_tracker_set_predictions(self._h, sequence, camera, p(idx), p(xy), len(keys)))

    def add_landmarks(self, index, xyz, descriptors, sequence: int = 0) -> int:
        """`system.points3d += ...` (slam_thread.cpp:210): landmarks whose index the device store already holds are skipped,
        the others appended in the order given.  index (n,) int, xyz (n, 3) world coordinates, descriptors (n, 32) u8.
        -> number of landmarks added.  Needs landmark_capacity > 0 at construction."""
        import ctypes as C

        import numpy as np

        from ._lib import check, lib
        idx = np.ascontiguousarray(index, np.int32); xyz = np.ascontiguousarray(xyz, np.float64).reshape(-1, 3)
        desc = np.ascontiguousarray(descriptors, np.uint8).reshape(-1, 32)
        assert len(idx) == len(xyz) == len(desc)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        added = C.c_int(0)
        check(lib().zs_tracker_landmarks_add_host(self._h, sequence, p(idx), p(xyz), p(desc), len(idx), C.byref(added)))
        return added.value

    def landmarks_size(self, sequence: int = 0) -> int:
        from ._lib import lib
        return int(lib().zs_tracker_landmarks_size(self._h, sequence))

    def set_camera_center(self, center, sequence: int = 0):
        """frame_0.pose.translation(): centre of the radius search of the next step's assign_landmark_indices
        (keypoint_tracker.cpp:56,72,213)"""
        import ctypes as C

        import numpy as np

        from ._lib import check, lib
        c = np.ascontiguousarray(center, np.float64).reshape(3)
        check(lib().zs_tracker_set_camera_center(self._h, sequence, c.ctypes.data_as(C.c_void_p)))

    def track_all(self, left, right):
        """left / right: (sequences, H, W) u8 -> [(keypoints_0, keypoints_1)] per sequence (keypoint_map each); the
        per-sequence index counters are in self.next_index afterwards"""
        import numpy as np
        S = self.sequences
        left = np.ascontiguousarray(left, np.uint8).reshape(S, self.height, self.width)
        right = np.ascontiguousarray(right, np.uint8).reshape(S, self.height, self.width)
        return self._run(left, right, host=True)

    def track_device(self, left, right):
        """the same step with (sequences, H, W) u8 CUDA tensors: enqueued on the context's stream, no host copy;
        call download() when the maps are wanted"""
        import ctypes as C

        from ._lib import check, lib
        assert left.is_cuda and right.is_cuda and left.is_contiguous() and right.is_contiguous()
        check(lib().zs_tracker_track(self._h, C.c_void_p(left.data_ptr()), C.c_void_p(right.data_ptr()), self.width,
                                     self.width * self.height))

    def download(self):
        """-> [(keypoints_0, keypoints_1)] per sequence: the maps after the last step"""
        return self._run(None, None, host=False)

    def filter_epipolar(self, fundamental, threshold: float, sequence: int = 0):
        """keypoint_tracker::filter_epipolar with a caller-supplied F (3x3): the maps of the last step keep only the
        keypoints present in both cameras with |pt0^T F pt1| < threshold; the next step tracks from the filtered maps"""
        import ctypes as C

        import numpy as np

        from ._lib import check, lib
        F = np.ascontiguousarray(fundamental, np.float64).reshape(3, 3)
        check(lib().zs_tracker_filter_epipolar(self._h, sequence, F.ctypes.data_as(C.c_void_p), float(threshold)))

    def submit(self, left, right):
        """pipelined host path (zs_tracker_submit_host): enqueue one step from (sequences, H, W) u8 pinned tensors / arrays;
        up to two steps may be in flight, wait() returns them oldest first.  The frames must stay alive until then."""
        import ctypes as C

        import numpy as np

        from ._lib import TrackerResults, check, lib
        torch = __import__("torch")
        if not hasattr(self, "_slots"):
            S, cap = self.sequences, self.cap
            pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
            self._slots, self._queue, self._step = [], [], 0
            for _ in range(2):
                d = dict(n=pin((S, 2), torch.int32), nxt=pin((S,), torch.int32), idx=[pin((S, cap), torch.int32) for _ in range(2)],
                         xy=[pin((S, cap, 2), torch.float32) for _ in range(2)], resp=[pin((S, cap), torch.float32) for _ in range(2)],
                         desc=[pin((S, cap, 32), torch.uint8) for _ in range(2)])
                r = TrackerResults(); r.cap = cap; r.n = d["n"].data_ptr(); r.next_index = d["nxt"].data_ptr()
                for c in range(2):
                    r.index[c] = d["idx"][c].data_ptr(); r.xy[c] = d["xy"][c].data_ptr(); r.response[c] = d["resp"][c].data_ptr()
                    r.desc[c] = d["desc"][c].data_ptr()
                d["r"] = r
                self._slots.append(d)
        if len(self._queue) == 2:
            raise RuntimeError("two steps are already in flight: call wait() first")
        as_t = lambda a: a if hasattr(a, "data_ptr") else torch.from_numpy(np.ascontiguousarray(a, np.uint8))
        lt, rt = as_t(left).contiguous(), as_t(right).contiguous()
        d = self._slots[self._step & 1]
        check(lib().zs_tracker_submit_host(self._h, C.c_void_p(lt.data_ptr()), C.c_void_p(rt.data_ptr()), self.width,
                                           self.width * self.height, C.byref(d["r"])))
        self._queue.append((d, lt, rt))
        self._step += 1

    def wait(self):
        """-> [(keypoints_0, keypoints_1)] per sequence of the oldest step in flight"""
        from ._lib import check, lib
        from .types import keypoint
        d, _, _ = self._queue.pop(0)
        check(lib().zs_tracker_wait(self._h))
        n, nxt = d["n"].numpy(), d["nxt"].numpy()
        self.next_index = [int(v) for v in nxt]
        out = []
        for s in range(self.sequences):
            maps = []
            for c in range(2):
                idx, xy, resp, desc = d["idx"][c].numpy(), d["xy"][c].numpy(), d["resp"][c].numpy(), d["desc"][c].numpy()
                m = keypoint_map()
                for i in range(int(n[s, c])):
                    m[int(idx[s, i])] = keypoint(pt=(float(xy[s, i, 0]), float(xy[s, i, 1])), response=float(resp[s, i]),
                                                 index=int(idx[s, i]), descriptor=desc[s, i].copy())
                maps.append(m)
            out.append((maps[0], maps[1]))
        return out

    def _run(self, left, right, host):
        import ctypes as C

        import numpy as np

        from ._lib import TrackerResults, check, lib
        from .types import keypoint
        S = self.sequences
        cap = self.cap
        n = np.zeros((S, 2), np.int32); nxt = np.zeros(S, np.int32)
        idx = [np.empty((S, cap), np.int32) for _ in range(2)]; xy = [np.empty((S, cap, 2), np.float32) for _ in range(2)]
        resp = [np.empty((S, cap), np.float32) for _ in range(2)]; desc = [np.empty((S, cap, 32), np.uint8) for _ in range(2)]
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        r = TrackerResults()
        r.cap = cap; r.n = p(n).value; r.next_index = p(nxt).value
        for c in range(2):
            r.index[c] = p(idx[c]).value; r.xy[c] = p(xy[c]).value; r.response[c] = p(resp[c]).value; r.desc[c] = p(desc[c]).value
        if host:
            check(lib().zs_tracker_track_host(self._h, p(left), p(right), self.width, self.width * self.height, C.byref(r)))
        else:
            check(lib().zs_tracker_download(self._h, C.byref(r)))
        self.next_index = [int(v) for v in nxt]
        out = []
        for s in range(S):
            maps = []
            for c in range(2):
                m = keypoint_map()
                for i in range(int(n[s, c])):
                    m[int(idx[c][s, i])] = keypoint(pt=(float(xy[c][s, i, 0]), float(xy[c][s, i, 1])), response=float(resp[c][s, i]),
                                                    index=int(idx[c][s, i]), descriptor=desc[c][s, i].copy())
                maps.append(m)
            out.append((maps[0], maps[1]))
        return out

    def track(self, left, right):
        """one sequence: -> (keypoints_0, keypoints_1) like keypoint_tracker.track; advances keypoint.index_next"""
        from .types import keypoint
        assert self.sequences == 1
        k0, k1 = self.track_all(left, right)[0]
        keypoint.index_next = self.next_index[0]
        return k0, k1
