"""The reference's hot path driven through the REAL dependency (Python cv2) -- TEST INFRASTRUCTURE.

The reference (vinodkhare/zenslam) is C++ glue over OpenCV and cannot be compiled here (no C++
OpenCV / VTK / Ceres).  This module restates that glue in Python over genuine
``cv2`` calls, so every arithmetic result comes from OpenCV itself:

* it pins ``zs_oracle.c`` (``tests/golden/make_golden.py`` writes fixtures from these functions);
* ``bench.py`` times :func:`stereo_frame` as the CPU baseline (``cpu_baseline`` / ``--impl reference``).

Each function cites the reference code it mirrors.  Imports cv2 lazily so that the oracle
package itself stays importable without it.
"""
from __future__ import annotations

import numpy as np


def _cv2():
    import cv2
    return cv2


def grid_detect(img, cell=(16, 16), threshold=10, occupied=None):
    """keypoint_detector_grid::detect_keypoints, detection half
    (zenslam_core/source/detection/keypoint_detector_grid.cpp:39-120)."""
    cv2 = _cv2()
    fast = cv2.FastFeatureDetector_create(int(threshold))
    h, w = img.shape
    gw, gh = w // cell[0], h // cell[1]
    xs, ys, sc = [], [], []
    for gy in range(gh):
        for gx in range(gw):
            if occupied is not None and occupied[gy, gx]:
                continue
            x0, y0 = gx * cell[0], gy * cell[1]
            roi = img[y0:y0 + min(cell[1], h - y0), x0:x0 + min(cell[0], w - x0)]
            kps = fast.detect(roi, None)
            if not kps:
                continue  # the ORB-detect fallback (:92-95) cannot fire on cells <= 62 px
            best = 0
            for i in range(1, len(kps)):       # std::ranges::max_element: first maximum
                if kps[i].response > kps[best].response:
                    best = i
            xs.append(kps[best].pt[0] + x0); ys.append(kps[best].pt[1] + y0); sc.append(kps[best].response)
    return np.array(xs, np.float32), np.array(ys, np.float32), np.array(sc, np.float32)


def orb_compute(img, xs, ys, angles=None):
    """_describer->compute(image, keypoints_cv, descriptors) with cv::ORB::create()
    (keypoint_detector_grid.cpp:28,138).  Returns (kept xs, kept ys, desc)."""
    cv2 = _cv2()
    orb = cv2.ORB_create()
    kps = [cv2.KeyPoint(float(x), float(y), 7.0, -1.0 if angles is None else float(angles[i]), 0.0, 0, -1)
           for i, (x, y) in enumerate(zip(xs, ys))]
    kps2, desc = orb.compute(img, kps)
    if desc is None:
        desc = np.zeros((0, 32), np.uint8)
    kx = np.array([k.pt[0] for k in kps2], np.float32)
    ky = np.array([k.pt[1] for k in kps2], np.float32)
    return kx, ky, desc


def corner_subpix(img, xs, ys):
    """cv::cornerSubPix(image, pts, Size(5,5), Size(-1,-1), {EPS+COUNT, 30, 0.01}) of the PARALLEL_GRID detector
    (keypoint_detector_parallel.cpp:160-170)."""
    cv2 = _cv2()
    if len(xs) == 0:
        return xs, ys
    c = np.stack([xs, ys], 1).astype(np.float32).reshape(-1, 1, 2).copy()
    cv2.cornerSubPix(img, c, (5, 5), (-1, -1), (cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 30, 0.01))
    return np.ascontiguousarray(c[:, 0, 0]), np.ascontiguousarray(c[:, 0, 1])


def detect_keypoints(img, cell=(16, 16), threshold=10, occupied=None, parallel_grid=False):
    """Full keypoint_detector_grid::detect_keypoints (or, with parallel_grid, keypoint_detector_parallel::detect_keypoints,
    keypoint_detector_parallel.cpp:40-193) -> (x, y, response, desc)."""
    xs, ys, sc = grid_detect(img, cell, threshold, occupied)
    if parallel_grid:
        xs, ys = corner_subpix(img, xs, ys)
    kx, ky, desc = orb_compute(img, xs, ys)
    # responses of the survivors: ORB::compute keeps order, so match by position
    rx, ry = np.rint(xs), np.rint(ys)          # ORB's border filter tests cvRound(pt) (round half to even, like np.rint)
    keep = (rx >= 31) & (rx < img.shape[1] - 31) & (ry >= 31) & (ry < img.shape[0] - 31)
    return kx, ky, sc[keep], desc


def match_knn_ratio(d0, d1, ratio=0.8, norm="hamming"):
    """matcher::match_keypoints KNN branch before the RANSAC gate (matcher.cpp:60-75)."""
    cv2 = _cv2()
    bf = cv2.BFMatcher(cv2.NORM_HAMMING if norm == "hamming" else cv2.NORM_L2, False)
    out = []
    if len(d0) and len(d1):
        for knn in bf.knnMatch(d0, d1, 2):
            if len(knn) == 2 and knn[0].distance < ratio * knn[1].distance:
                out.append((knn[0].queryIdx, knn[0].trainIdx, knn[0].distance))
    return out


def match_cross(d0, d1, norm="hamming"):
    """matcher::match_keypoints BRUTE branch (matcher.cpp:76-80; matching_utils.cpp:90-94)."""
    cv2 = _cv2()
    bf = cv2.BFMatcher(cv2.NORM_HAMMING if norm == "hamming" else cv2.NORM_L2, True)
    if not len(d0) or not len(d1):
        return []
    return [(m.queryIdx, m.trainIdx, m.distance) for m in bf.match(d0, d1)]


def lk(img0, img1, p0, p1_init=None, win=(31, 31), max_level=3, max_iters=99, eps=0.001, min_eig=1e-4):
    """pyr_lk::calc_optical_flow_pyr_lk as the reference calls it (keypoint_tracker.cpp:142-170,379-407).
    Python's binding accepts raw images only; OpenCV builds the identical pyramid internally."""
    cv2 = _cv2()
    p0 = np.ascontiguousarray(p0, np.float32).reshape(-1, 1, 2)
    if len(p0) == 0:
        return np.zeros((0, 2), np.float32), np.zeros(0, np.uint8), np.zeros(0, np.float32)
    flags = cv2.OPTFLOW_LK_GET_MIN_EIGENVALS
    nxt = None
    if p1_init is not None:
        flags |= cv2.OPTFLOW_USE_INITIAL_FLOW
        nxt = np.ascontiguousarray(p1_init, np.float32).reshape(-1, 1, 2).copy()
    p1, st, err = cv2.calcOpticalFlowPyrLK(
        img0, img1, p0, nxt, winSize=tuple(win), maxLevel=int(max_level),
        criteria=(cv2.TERM_CRITERIA_COUNT | cv2.TERM_CRITERIA_EPS, int(max_iters), float(eps)),
        flags=flags, minEigThreshold=float(min_eig))
    return p1.reshape(-1, 2), st.reshape(-1), err.reshape(-1)


def track_fb(img0, img1, p0, p1_init=None, win=(31, 31), max_level=3, klt_threshold=1.0):
    """keypoint_tracker::track_keypoints: forward + backward LK and the FB gate
    (keypoint_tracker.cpp:129-197).  Returns (p1, keep mask)."""
    p1, st, _ = lk(img0, img1, p0, p1_init, win, max_level)
    pb, sb, _ = lk(img1, img0, p1, None, win, max_level)
    d = pb - np.asarray(p0, np.float32).reshape(-1, 2)
    nrm = np.sqrt(d[:, 0].astype(np.float64) ** 2 + d[:, 1].astype(np.float64) ** 2)
    keep = (st != 0) & (sb != 0) & (nrm < klt_threshold)
    return p1, keep


def stereo_frame(prev_l, prev_r, cur_l, cur_r, prev_kp_l, prev_kp_r, opts):
    """One stereo frame of the hot path in the reference's call pattern:
    2 pyramids, 2 grid detections + ORB, 1 stereo kNN-ratio match, 4 forward+backward KLT pairs.
    prev_kp_* are (n,2) float32 keypoint positions detected on the previous frame."""
    cv2 = _cv2()
    win, ml = tuple(opts.klt_window_size), int(opts.klt_max_level)
    cv2.buildOpticalFlowPyramid(cur_l, win, ml)          # utils::pyramid (utils_opencv.cpp:525-530)
    cv2.buildOpticalFlowPyramid(cur_r, win, ml)
    pg = bool(getattr(opts, "parallel_grid", False))
    xl, yl, rl, dl = detect_keypoints(cur_l, opts.cell_size, opts.fast_threshold, None, pg)
    xr, yr, rr, dr = detect_keypoints(cur_r, opts.cell_size, opts.fast_threshold, None, pg)
    matches = match_knn_ratio(dl, dr, opts.matcher_ratio)
    kl = np.stack([xl, yl], 1) if len(xl) else np.zeros((0, 2), np.float32)
    kr = np.stack([xr, yr], 1) if len(xr) else np.zeros((0, 2), np.float32)
    t_l = track_fb(prev_l, cur_l, prev_kp_l, None, win, ml, opts.klt_threshold)
    t_r = track_fb(prev_r, cur_r, prev_kp_r, None, win, ml, opts.klt_threshold)
    s_lr = track_fb(cur_l, cur_r, kl, None, win, ml, opts.klt_threshold)
    s_rl = track_fb(cur_r, cur_l, kr, None, win, ml, opts.klt_threshold)
    return dict(kp_l=kl, kp_r=kr, resp_l=rl, resp_r=rr, desc_l=dl, desc_r=dr, matches=matches,
                temporal_l=t_l, temporal_r=t_r, stereo_lr=s_lr, stereo_rl=s_rl)


def overhead_estimate(img, opts, reps=3):
    """What `stereo_frame` spends per stereo frame on work the C++ reference does not do (bench.py reports it next to the CPU
    baseline so that the GPU/CPU ratio can be bounded):
      * pyramid rebuilds -- Python's calcOpticalFlowPyrLK takes raw images only, so each of the 8 LK calls rebuilds both image
        pyramids and the Scharr planes of its source image, while the reference hands over the two pyramids processor.cpp built;
      * interpreter time of the per-cell loop of `grid_detect` (the reference's loop is C++).
    Returns seconds per stereo frame: dict(pyramid_rebuild=..., python_cell_loop=...)."""
    import time
    cv2 = _cv2()
    win, ml = tuple(opts.klt_window_size), int(opts.klt_max_level)
    ta = tb = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); cv2.buildOpticalFlowPyramid(img, win, ml, None, True); ta = min(ta, time.perf_counter() - t0)
        t0 = time.perf_counter(); cv2.buildOpticalFlowPyramid(img, win, ml, None, False); tb = min(tb, time.perf_counter() - t0)
    pyr = 8 * 2 * tb + 8 * max(0.0, ta - tb)
    fast = cv2.FastFeatureDetector_create(int(opts.fast_threshold))
    h, w = img.shape
    cw, ch = opts.cell_size
    loop = 1e9
    for _ in range(reps):
        inner = 0.0
        t0 = time.perf_counter()
        for gy in range(h // ch):
            for gx in range(w // cw):
                roi = img[gy * ch:(gy + 1) * ch, gx * cw:(gx + 1) * cw]
                t1 = time.perf_counter()
                fast.detect(roi, None)
                inner += time.perf_counter() - t1
        loop = min(loop, time.perf_counter() - t0 - inner)
    return dict(pyramid_rebuild=pyr, python_cell_loop=2 * loop)
