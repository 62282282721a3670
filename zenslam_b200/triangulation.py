"""triangulator mirror (zenslam_core/source/mapping/triangulator.cpp:31-188), keypoint part: epipolar filter,
cv::triangulatePoints, reprojection / depth / parallax gates -- one GPU thread per matched pair.
Keylines, colours and the point3d_cloud container are outside this backend."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from ._lib import TriangulationParams, check, lib
from .runtime import Context


@dataclass
class triangulation_options:
    """slam.triangulation keys used by this path (all_options.h:35-45)"""
    reprojection_threshold: float = 1.0
    min_depth: float = 1.0
    max_depth: float = 50.0
    filter_epipolar: bool = True
    epipolar_threshold: float = 0.01


@dataclass
class point3d:
    x: float
    y: float
    z: float
    index: int
    descriptor: np.ndarray | None = None


class triangulator:
    """calibration: projection_matrix[0], projection_matrix[1] (3x4), fundamental_matrix[0] (3x3),
    cameras[1].pose_in_cam0.translation() (3,)"""

    def __init__(self, ctx: Context, projection_0, projection_1, fundamental, translation_1_in_0, options: triangulation_options | None = None):
        self._ctx = ctx
        self._P0 = np.ascontiguousarray(projection_0, np.float64).reshape(3, 4)
        self._P1 = np.ascontiguousarray(projection_1, np.float64).reshape(3, 4)
        self._F = None if fundamental is None else np.ascontiguousarray(fundamental, np.float64).reshape(3, 3)
        self._t = np.ascontiguousarray(translation_1_in_0, np.float64).reshape(3)
        self._o = options or triangulation_options()

    def triangulate_points(self, pts0: np.ndarray, pts1: np.ndarray, with_diag: bool = False):
        """-> xyz (n,3) f64 for every pair, keep (n,) bool [, diag (n,4)]"""
        pts0 = np.ascontiguousarray(pts0, np.float32).reshape(-1, 2); pts1 = np.ascontiguousarray(pts1, np.float32).reshape(-1, 2)
        n = len(pts0)
        assert len(pts1) == n
        xyz = np.zeros((n, 3), np.float64); keep = np.zeros(n, np.uint8); diag = np.zeros((n, 4), np.float64) if with_diag else None
        prm = TriangulationParams(1 if self._o.filter_epipolar else 0, self._o.epipolar_threshold, self._o.reprojection_threshold,
                                  self._o.min_depth, self._o.max_depth)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        check(lib().zs_triangulate_keypoints_host(self._ctx._h, p(self._P0), p(self._P1), p(self._F) if self._F is not None else None,
                                                  p(self._t), p(pts0), p(pts1), n, C.byref(prm), p(xyz), p(keep),
                                                  p(diag) if diag is not None else None))
        return (xyz, keep.astype(bool), diag) if with_diag else (xyz, keep.astype(bool))

    def triangulate_keypoints(self, keypoints_0: dict, keypoints_1: dict) -> list:
        """map<keypoint> x map<keypoint> -> surviving point3d list (triangulator.cpp:39-132); pairs are the keypoints
        whose index is present in both maps, in ascending index order (map::values_matched)"""
        idx = sorted(i for i in keypoints_0 if i in keypoints_1)
        if not idx:
            return []
        p0 = np.array([keypoints_0[i].pt for i in idx], np.float32); p1 = np.array([keypoints_1[i].pt for i in idx], np.float32)
        xyz, keep = self.triangulate_points(p0, p1)
        return [point3d(float(xyz[k, 0]), float(xyz[k, 1]), float(xyz[k, 2]), i, keypoints_0[i].descriptor)
                for k, i in enumerate(idx) if keep[k]]
