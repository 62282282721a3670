"""ctypes access to oracle/_ref/libzs_ref_glue.so -- the REFERENCE's own detector classes (keypoint_detector_grid / _parallel /
_simple, compiled unmodified by oracle/build_ref.py) running on the OpenCV stand-in backed by the C oracle.  TEST INFRASTRUCTURE."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libzs_ref_glue.so")
SIMPLE, GRID, PARALLEL_GRID = 0, 1, 2
_lib = None


def available() -> bool:
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB)
        _lib.zref_detect_keypoints.restype = C.c_int
        _lib.zref_detect_keypoints.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                               C.POINTER(C.c_longlong)] + [C.c_void_p] * 7 + [C.c_int]
    return _lib


def detect_keypoints(algorithm: int, img: np.ndarray, cell=(16, 16), fast_threshold=10, existing=None, index_next=0):
    """zenslam::keypoint_detector_{simple,grid,parallel}::detect_keypoints(image, keypoints_existing) of the reference.
    existing: (n, 3) rows of (index, x, y).  -> dict(xy, response, size, angle, octave, index, desc, index_next)"""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    ex = np.zeros((0, 3), np.float32) if existing is None else np.ascontiguousarray(existing, np.float32).reshape(-1, 3)
    cap = max(16, w * h // 4)
    xy = np.zeros((cap, 2), np.float32); resp = np.zeros(cap, np.float32); size = np.zeros(cap, np.float32); ang = np.zeros(cap, np.float32)
    octv = np.zeros(cap, np.int32); idx = np.zeros(cap, np.int64); desc = np.zeros((cap, 32), np.uint8)
    nxt = C.c_longlong(int(index_next))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    n = lib().zref_detect_keypoints(int(algorithm), p(img), w, h, w, int(cell[0]), int(cell[1]), int(fast_threshold), p(ex), len(ex),
                                    C.byref(nxt), p(xy), p(resp), p(size), p(ang), p(octv), p(idx), p(desc), cap)
    if n < 0:
        raise RuntimeError("capacity %d too small for %d keypoints" % (cap, -1 - n))
    return dict(xy=xy[:n].copy(), response=resp[:n].copy(), size=size[:n].copy(), angle=ang[:n].copy(), octave=octv[:n].copy(),
                index=idx[:n].copy(), desc=desc[:n].copy(), index_next=int(nxt.value))
