"""pyr_lk plug-in mirror (zenslam_core/include/zenslam/tracking/pyr_lk.h:10-29) and the KLT glue of
keypoint_tracker::track_keypoints (zenslam_core/source/tracking/keypoint_tracker.cpp:129-197, 343-434).

`create_cuda_pyr_lk()` is the sibling of `metal::create_metal_pyr_lk()`
(zenslam_metal/source/pyr_lk_factory.cpp:41-49): it returns None when the backend is unavailable so the
caller can fall back to its own CPU implementation; the CUDA backend itself never falls back.
"""
from __future__ import annotations

import ctypes as C
import dataclasses

import numpy as np

from ._lib import LK_GET_MIN_EIGENVALS, LK_USE_INITIAL_FLOW, LkParams, check, lib
from .options import tracking_options
from .runtime import Context, is_available


class pyr_lk:
    """Abstract KLT backend: same argument list as the reference's virtual (pyr_lk.h:15-26)."""

    def calc_optical_flow_pyr_lk(self, prev_pyramid, next_pyramid, prev_points, next_points, win_size, max_level,
                                 criteria, flags, min_eig_threshold=1e-4):
        """-> (next_points (n,2) f32, status (n,) u8, err (n,) f32)"""
        raise NotImplementedError


def _level0(pyramid) -> np.ndarray:
    # the reference passes cv::buildOpticalFlowPyramid output [img0, deriv0, img1, ...]; level 0 determines the rest
    img = pyramid[0] if isinstance(pyramid, (list, tuple)) else pyramid
    return np.ascontiguousarray(img, np.uint8)


class cuda_pyr_lk(pyr_lk):
    def __init__(self, ctx: Context):
        self._ctx = ctx

    def calc_optical_flow_pyr_lk(self, prev_pyramid, next_pyramid, prev_points, next_points, win_size, max_level,
                                 criteria=(99, 0.001), flags=LK_GET_MIN_EIGENVALS, min_eig_threshold=1e-4):
        prev, nxt = _level0(prev_pyramid), _level0(next_pyramid)
        assert prev.shape == nxt.shape
        h, w = prev.shape
        pp = np.ascontiguousarray(prev_points, np.float32).reshape(-1, 2)
        n = len(pp)
        if flags & LK_USE_INITIAL_FLOW:
            npts = np.ascontiguousarray(next_points, np.float32).reshape(-1, 2).copy()
            assert len(npts) == n
        else:
            npts = np.zeros((n, 2), np.float32)
        status = np.zeros(n, np.uint8); err = np.zeros(n, np.float32)
        prm = LkParams(win_size[0], win_size[1], int(max_level), int(criteria[0]), float(criteria[1]), int(flags),
                       float(min_eig_threshold))
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        check(lib().zs_calc_optical_flow_pyr_lk_host(self._ctx._h, p(prev), p(nxt), w, h, w, p(pp), p(npts), n, p(status),
                                                     p(err), C.byref(prm)))
        return npts, status, err


    def track_keypoints_fused(self, pyramid_0, pyramid_1, points_0, predicted_1, win_size, max_level, klt_threshold,
                              criteria=(99, 0.001), min_eig_threshold=1e-4):
        """zs_track_keypoints_host: forward + backward LK + FB gate in one device call -> (points_1, status, keep)"""
        a, b = _level0(pyramid_0), _level0(pyramid_1)
        h, w = a.shape
        p0 = np.ascontiguousarray(points_0, np.float32).reshape(-1, 2)
        n = len(p0)
        pred = None if predicted_1 is None else np.ascontiguousarray(predicted_1, np.float32).reshape(-1, 2)
        p1 = np.zeros((n, 2), np.float32); status = np.zeros(n, np.uint8); keep = np.zeros(n, np.uint8)
        prm = LkParams(win_size[0], win_size[1], int(max_level), int(criteria[0]), float(criteria[1]), LK_GET_MIN_EIGENVALS,
                       float(min_eig_threshold))
        p = lambda x: x.ctypes.data_as(C.c_void_p)
        check(lib().zs_track_keypoints_host(self._ctx._h, p(a), p(b), w, h, w, p(p0), p(pred) if pred is not None else None, n,
                                            C.byref(prm), float(klt_threshold), p(p1), p(status), None, p(keep)))
        return p1, status, keep


def create_cuda_pyr_lk(ctx: Context | None = None):
    """Factory; None when no sm_100 device is usable (cf. pyr_lk_factory.cpp:43-46)."""
    if not is_available():
        return None
    return cuda_pyr_lk(ctx if ctx is not None else Context())


def track_keypoints(backend: pyr_lk, pyramid_0, pyramid_1, keypoints_0: list, tracking: tracking_options,
                    predicted_points=None, fused: bool = False) -> list:
    """keypoint_tracker::track_keypoints: forward LK (initial flow from `predicted_points` when given --
    the temporal overload, keypoint_tracker.cpp:361-391), backward LK, keep i iff both statuses are set and
    ||p0_back - p0|| < klt_threshold; survivors copy the keypoint with pt replaced (keypoint_tracker.cpp:188-196)."""
    if not keypoints_0:
        return []
    p0 = np.array([kp.pt for kp in keypoints_0], np.float32)
    if fused and isinstance(backend, cuda_pyr_lk):
        # one device call instead of the two pyr_lk calls + host gate below: same results (tests compare both routes)
        p1, _, keep = backend.track_keypoints_fused(pyramid_0, pyramid_1, p0, predicted_points, tracking.klt_window_size,
                                                    tracking.klt_max_level, tracking.klt_threshold)
        return [dataclasses.replace(kp, pt=(float(p1[i, 0]), float(p1[i, 1]))) for i, kp in enumerate(keypoints_0) if keep[i]]
    crit = (99, 0.001)
    if predicted_points is not None:
        p1, st, _ = backend.calc_optical_flow_pyr_lk(pyramid_0, pyramid_1, p0, predicted_points, tracking.klt_window_size,
                                                     tracking.klt_max_level, crit,
                                                     LK_GET_MIN_EIGENVALS | LK_USE_INITIAL_FLOW)
    else:
        p1, st, _ = backend.calc_optical_flow_pyr_lk(pyramid_0, pyramid_1, p0, None, tracking.klt_window_size,
                                                     tracking.klt_max_level, crit, LK_GET_MIN_EIGENVALS)
    pb, sb, _ = backend.calc_optical_flow_pyr_lk(pyramid_1, pyramid_0, p1, None, tracking.klt_window_size,
                                                 tracking.klt_max_level, crit, LK_GET_MIN_EIGENVALS)
    out = []
    for i, kp in enumerate(keypoints_0):
        dx, dy = np.float32(pb[i, 0] - p0[i, 0]), np.float32(pb[i, 1] - p0[i, 1])
        nrm = float(np.sqrt(np.float64(dx) * np.float64(dx) + np.float64(dy) * np.float64(dy)))
        if st[i] and sb[i] and nrm < tracking.klt_threshold:
            out.append(dataclasses.replace(kp, pt=(float(p1[i, 0]), float(p1[i, 1]))))
    return out


def track_keylines(backend: pyr_lk, pyramid_0, pyramid_1, keylines_0, tracking: tracking_options) -> list:
    """utils::track_keylines (zenslam_core/source/tracking/tracking_utils.cpp:14-143): both endpoints of every keyline
    go through the same pyr_lk seam -- forward 0 -> 1, backward 1 -> 0 from the forward results, (99, 0.001),
    OPTFLOW_LK_GET_MIN_EIGENVALS; a keyline survives iff all four statuses are set and both forward-backward errors
    (cv::norm narrowed to float) are below klt_threshold; its endpoints, midpoint, length and angle are rewritten."""
    values = list(keylines_0.values()) if hasattr(keylines_0, "values") else list(keylines_0)
    if not values:
        return []
    f32 = np.float32
    s0 = np.array([(k.startPointX, k.startPointY) for k in values], f32)
    e0 = np.array([(k.endPointX, k.endPointY) for k in values], f32)
    crit = (99, 0.001)
    args = (tracking.klt_window_size, tracking.klt_max_level, crit, LK_GET_MIN_EIGENVALS)
    s1, st_sf, _ = backend.calc_optical_flow_pyr_lk(pyramid_0, pyramid_1, s0, None, *args)
    e1, st_ef, _ = backend.calc_optical_flow_pyr_lk(pyramid_0, pyramid_1, e0, None, *args)
    sb, st_sb, _ = backend.calc_optical_flow_pyr_lk(pyramid_1, pyramid_0, s1, None, *args)
    eb, st_eb, _ = backend.calc_optical_flow_pyr_lk(pyramid_1, pyramid_0, e1, None, *args)

    def fb_error(back, orig):
        dx, dy = f32(back[0] - orig[0]), f32(back[1] - orig[1])
        return f32(np.sqrt(np.float64(dx) * np.float64(dx) + np.float64(dy) * np.float64(dy)))

    out = []
    for i, kl in enumerate(values):
        if not (st_sf[i] and st_ef[i] and st_sb[i] and st_eb[i]):
            continue
        if not (float(fb_error(sb[i], s0[i])) < tracking.klt_threshold and float(fb_error(eb[i], e0[i])) < tracking.klt_threshold):
            continue
        dx, dy = f32(e1[i, 0] - s1[i, 0]), f32(e1[i, 1] - s1[i, 1])
        out.append(dataclasses.replace(
            kl, startPointX=float(s1[i, 0]), startPointY=float(s1[i, 1]), endPointX=float(e1[i, 0]), endPointY=float(e1[i, 1]),
            pt=(float(f32(f32(s1[i, 0] + e1[i, 0]) * f32(0.5))), float(f32(f32(s1[i, 1] + e1[i, 1]) * f32(0.5)))),
            lineLength=float(np.sqrt(f32(f32(dx * dx) + f32(dy * dy)))),
            angle=float(f32(f32(np.arctan2(dy, dx)) * f32(180.0)) / f32(np.pi))))
    return out
