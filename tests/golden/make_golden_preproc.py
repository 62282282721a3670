"""Golden fixture for the pre-processing path of processor::process (processor.cpp:25-55): cv2.cvtColor(BGR2GRAY),
cv2.createCLAHE(4.0).apply, cv2.remap(INTER_LINEAR) with float maps.  Run here:  python tests/golden/make_golden_preproc.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import cv2  # noqa: E402

from zenslam_b200 import synthetic as syn  # noqa: E402


def main():
    W, H = 188, 122                              # not a multiple of the 8x8 CLAHE grid: exercises the REFLECT_101 extension
    rng = np.random.default_rng(2024)
    tex = syn.crop(syn.base_texture(W, H, 808), W, H, 0, 0)
    bgr = np.stack([np.clip(tex.astype(np.int32) * s // 100 + o + rng.integers(-6, 7, tex.shape), 0, 255)
                    for s, o in ((45, 10), (60, 5), (38, 25))], -1).astype(np.uint8)      # dim, coloured
    gray = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
    clahe = cv2.createCLAHE(4.0).apply(gray)
    # a radial undistortion-like map from a real OpenCV call (what calibration.cpp:60-70 does)
    K = np.array([[0.8 * W, 0, W / 2 - 3.5], [0, 0.8 * W, H / 2 + 2.25], [0, 0, 1]], np.float64)
    dist = np.array([-0.28, 0.09, 0.0012, -0.0007, 0.0], np.float64)
    mx, my = cv2.initUndistortRectifyMap(K, dist, None, K, (W, H), cv2.CV_32FC1)
    mx[3, :7] = np.nan; my[5, 3] = np.inf; mx[7, 2] = -1e12; my[9, 4] = 3e9                  # hostile map entries
    out = dict(bgr=bgr, map_x=mx, map_y=my, cv_gray=gray, cv_clahe=clahe,
               cv_remap_gray=cv2.remap(gray, mx, my, cv2.INTER_LINEAR),
               cv_remap_clahe=cv2.remap(clahe, mx, my, cv2.INTER_LINEAR),
               cv_clahe_2_4x6=cv2.createCLAHE(2.0, (4, 6)).apply(gray), cv2_version=cv2.__version__)
    np.savez_compressed(os.path.join(HERE, "preproc.npz"), **out)
    print("wrote preproc.npz")


if __name__ == "__main__":
    main()
