"""keypoint_detector mirrors (zenslam_core/include/zenslam/detection/keypoint_detector.h:8-14).

Same names, argument meaning and results as the reference's detectors; the work happens in
libzenslam_cuda.so through the host-pointer C-ABI calls a C++ adapter would bind.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, lib
from .options import detection_options
from .runtime import Context, Pyramid, fast_detect, orb_compute
from .types import keypoint


class keypoint_detector:
    """Interface: detect_keypoints(image, keypoints_existing) -> list[keypoint] (keypoint_detector.h:13)."""

    def detect_keypoints(self, image: np.ndarray, keypoints_existing) -> list:
        raise NotImplementedError


def _require_fast_orb(options: detection_options):
    # keypoint_detector_grid.cpp:12-36: FAST -> ORB is the path this backend implements
    if options.feature_detector != "FAST" or options.descriptor != "ORB":
        raise NotImplementedError("the CUDA detector implements feature FAST + descriptor ORB (the reference default)")


def _c_div(a: int, b: int) -> int:
    """integer division as C++ does it (toward zero), not Python's floor"""
    q = abs(a) // b
    return q if a >= 0 else -q


def _emit(xs, ys, resp, desc) -> list:
    out = []
    for i in range(len(xs)):        # keypoint_detector_grid.cpp:142-147: sequential global indices
        out.append(keypoint(pt=(float(xs[i]), float(ys[i])), size=7.0, angle=-1.0, response=float(resp[i]),
                            octave=0, class_id=-1, index=keypoint.index_next, descriptor=desc[i]))
        keypoint.index_next += 1
    return out


class keypoint_detector_grid(keypoint_detector):
    """keypoint_detector_grid (zenslam_core/source/detection/keypoint_detector_grid.cpp:9-150)."""

    _entry = "zs_detect_keypoints_grid_host"

    def __init__(self, options: detection_options, ctx: Context):
        _require_fast_orb(options)
        self._options, self._ctx = options, ctx

    def detect_keypoints(self, image: np.ndarray, keypoints_existing=None) -> list:
        image = np.ascontiguousarray(image, np.uint8)
        h, w = image.shape
        cw, ch = self._options.cell_size
        gw, gh = w // cw, h // ch
        if gw * gh == 0:
            return []
        occ = None
        if keypoints_existing:
            # keypoint_detector_grid.cpp:48-63: occupied[int(pt.x)/cw][int(pt.y)/ch]
            occ = np.zeros((gh, gw), np.uint8)
            values = keypoints_existing.values() if hasattr(keypoints_existing, "values") else keypoints_existing
            for kp in values:
                # C++ semantics: the float truncates toward zero, and so does the integer division -- a keypoint tracked to
                # x in (-cell, 0) lands in column 0 (checked against the reference's own code: tests/test_oracle_vs_reference_glue.py)
                gx, gy = _c_div(int(kp.pt[0]), cw), _c_div(int(kp.pt[1]), ch)
                if 0 <= gx < gw and 0 <= gy < gh:
                    occ[gy, gx] = 1
        cells = gw * gh
        xs = np.empty(cells, np.float32); ys = np.empty(cells, np.float32); resp = np.empty(cells, np.float32)
        desc = np.empty((cells, 32), np.uint8)
        n = C.c_int(0)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        check(getattr(lib(), self._entry)(self._ctx._h, p(image), w, h, w, cw, ch, int(self._options.fast_threshold),
                                                  p(occ) if occ is not None else None, p(xs), p(ys), p(resp), p(desc),
                                                  C.byref(n)))
        k = n.value
        return _emit(xs[:k], ys[:k], resp[:k], desc[:k].copy())


class keypoint_detector_parallel(keypoint_detector_grid):
    """keypoint_detector_parallel (zenslam_core/source/detection/keypoint_detector_parallel.cpp:40-193): the grid
    detector's cells, then cv::cornerSubPix (win 5x5, 30 its, eps 0.01) on every selected corner, then ORB::compute
    at cvRound(pt).  The reference's per-cell std::async threads are an execution detail, not a result."""

    _entry = "zs_detect_keypoints_parallel_host"


class keypoint_detector_simple(keypoint_detector):
    """keypoint_detector_simple (zenslam_core/source/detection/keypoint_detector_simple.cpp:8-63): full-frame
    FAST with a mask that is 0 inside discs of radius min(cell)/2 around existing keypoints."""

    def __init__(self, options: detection_options, ctx: Context, cap: int = 1 << 17):
        # `feature: ORB` = cv::ORB::create(500, 1.2f, 8, 31, 0, 2, HARRIS_SCORE, 31, fast_threshold) as the detector
        # (keypoint_detector_simple.cpp:17); the descriptor stays cv::ORB::create() (:27)
        if options.feature_detector not in ("FAST", "ORB") or options.descriptor != "ORB":
            raise NotImplementedError("the CUDA detector implements feature FAST or ORB with descriptor ORB")
        self._options, self._ctx, self._cap = options, ctx, cap
        self._pyr = None

    def _detect_orb(self, image, mask) -> list:
        h, w = image.shape
        cap = 500 + 32 * 8 + 64
        x = np.empty(cap, np.float32); y = np.empty(cap, np.float32); size = np.empty(cap, np.float32)
        ang = np.empty(cap, np.float32); resp = np.empty(cap, np.float32); octv = np.empty(cap, np.int32)
        desc = np.empty((cap, 32), np.uint8)
        n = C.c_int(0)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        check(lib().zs_detect_keypoints_orb_host(self._ctx._h, p(image), w, h, w, p(mask) if mask is not None else None, w,
                                                 500, 1.2, 8, 31, 31, int(self._options.fast_threshold), p(x), p(y), p(size),
                                                 p(ang), p(resp), p(octv), p(desc), cap, C.byref(n)))
        out = []
        for i in range(n.value):
            out.append(keypoint(pt=(float(x[i]), float(y[i])), size=float(size[i]), angle=float(ang[i]),
                                response=float(resp[i]), octave=int(octv[i]), class_id=-1, index=keypoint.index_next,
                                descriptor=desc[i].copy()))
            keypoint.index_next += 1
        return out

    def _mask(self, h, w, keypoints_existing):
        if not keypoints_existing:
            return None
        mask = np.full((h, w), 255, np.uint8)
        r = min(self._options.cell_size) // 2
        yy, xx = np.mgrid[-r:r + 1, -r:r + 1]
        disc = (xx * xx + yy * yy) <= r * r
        values = keypoints_existing.values() if hasattr(keypoints_existing, "values") else keypoints_existing
        for kp in values:
            cx, cy = int(np.rint(np.float32(kp.pt[0]))), int(np.rint(np.float32(kp.pt[1])))   # cv::Point(Point2f) rounds
            y0, y1, x0, x1 = max(cy - r, 0), min(cy + r + 1, h), max(cx - r, 0), min(cx + r + 1, w)
            if y0 >= y1 or x0 >= x1:
                continue
            sub = disc[y0 - (cy - r):y1 - (cy - r), x0 - (cx - r):x1 - (cx - r)]
            mask[y0:y1, x0:x1][sub] = 0
        return mask

    def detect_keypoints(self, image: np.ndarray, keypoints_existing=None) -> list:
        image = np.ascontiguousarray(image, np.uint8)
        h, w = image.shape
        if self._options.feature_detector == "ORB":
            return self._detect_orb(image, self._mask(h, w, keypoints_existing))
        if self._pyr is None or (self._pyr.width, self._pyr.height) != (w, h):
            self._pyr = Pyramid(self._ctx, w, h, 1, (16, 16), 0)
        pyr = self._pyr
        pyr.upload(image, 0)
        pyr.build(0, 1)
        mask = self._mask(h, w, keypoints_existing)
        xy, resp, n = fast_detect(pyr, 0, 1, self._options.fast_threshold, None if mask is None else mask[None], self._cap)
        if int(n[0]) > self._cap:
            raise RuntimeError("keypoint_detector_simple: %d corners exceed the capacity %d" % (int(n[0]), self._cap))
        oxy, oresp, _, on, desc = orb_compute(pyr, 0, 1, xy, resp, n)
        k = int(on[0])
        oxy = oxy[0, :k].cpu().numpy(); oresp = oresp[0, :k].cpu().numpy(); desc = desc[0, :k].cpu().numpy()
        return _emit(oxy[:, 0], oxy[:, 1], oresp, desc)
