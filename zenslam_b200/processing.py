"""Pre-processing mirror of zenslam::processor (zenslam_core/source/processor.cpp:25-55), image path only:
convert_color(BGR2GRAY) -> optional CLAHE(4.0) -> rectify (cv::remap with the calibration's CV_32FC1 maps).
The IMU pre-integration that runs beside it in the reference is outside this backend."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, lib
from .runtime import Context


class processor:
    """process_image(image, camera) -> undistorted gray image (frame::processed::undistorted[camera]).

    `maps`: per camera (map_x, map_y) float32 arrays as produced by cv::initUndistortRectifyMap(..., CV_32FC1)
    (calibration.cpp:60-70), or None for no rectification.  clahe_clip_limit mirrors cv::createCLAHE(4.0)
    (processor.h:38)."""

    def __init__(self, ctx: Context, clahe_enabled: bool = False, maps=None, clahe_clip_limit: float = 4.0):
        self._ctx, self._clahe, self._clip = ctx, bool(clahe_enabled), float(clahe_clip_limit)
        self._maps = None
        if maps is not None:
            self._maps = [(np.ascontiguousarray(mx, np.float32), np.ascontiguousarray(my, np.float32)) for mx, my in maps]

    def process_image(self, image: np.ndarray, camera: int = 0) -> np.ndarray:
        image = np.ascontiguousarray(image, np.uint8)
        channels = 1 if image.ndim == 2 else image.shape[2]
        h, w = image.shape[:2]
        out = np.empty((h, w), np.uint8)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        mx = my = None
        if self._maps is not None:
            mx, my = self._maps[camera]
            assert mx.shape == (h, w) == my.shape
        check(lib().zs_process_image_host(self._ctx._h, p(image), channels, w, h, w * channels, 1 if self._clahe else 0, self._clip,
                                          p(mx) if mx is not None else None, p(my) if my is not None else None, p(out)))
        return out
