"""Golden fixture for cv::cornerSubPix as the PARALLEL_GRID detector calls it (keypoint_detector_parallel.cpp:160-170).

Run here (cv2 4.13.0 importable):  python tests/golden/make_golden_subpix.py  -> tests/golden/subpix.npz
`cv_subpix`     cv2.cornerSubPix with IPP switched OFF (the plain OpenCV code path the C oracle restates; a vcpkg OpenCV
                -- what the reference links -- is built without IPP by default)
`cv_subpix_ipp` the same call with IPP on (opencv-python's default): ~10 % of the points differ by up to ~3e-3 px
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import cv2  # noqa: E402

from oracle import cv2_ref  # noqa: E402
from zenslam_b200 import synthetic as syn  # noqa: E402


def main():
    W, H = 256, 192
    L = syn.crop(syn.base_texture(W, H, 4242), W, H, 0, 0)
    x, y, _ = cv2_ref.grid_detect(L, (16, 16), 10)
    pts = np.stack([x, y], 1).astype(np.float32)
    # a few hand-placed points: image corners / edges (border-replicating sampling), flat-ish spots, sub-pixel starts
    extra = np.array([[0, 0], [3, 43], [255, 191], [250.5, 3.25], [6, 6], [7, 7], [128.4, 96.7], [12.75, 180.1]], np.float32)
    pts = np.concatenate([pts, extra])
    crit = (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_COUNT, 30, 0.01)
    out = dict(L=L, pts=pts, cv2_version=cv2.__version__)
    for name, ipp in (("cv_subpix", False), ("cv_subpix_ipp", True)):
        cv2.ipp.setUseIPP(ipp)
        ref = pts.copy().reshape(-1, 1, 2)
        cv2.cornerSubPix(L, ref, (5, 5), (-1, -1), crit)
        out[name] = ref.reshape(-1, 2)
    cv2.ipp.setUseIPP(True)
    # the whole PARALLEL_GRID detector: grid corners -> cornerSubPix -> ORB::compute
    cv2.ipp.setUseIPP(False)
    ref = np.stack([x, y], 1).astype(np.float32).reshape(-1, 1, 2)
    cv2.cornerSubPix(L, ref, (5, 5), (-1, -1), crit)
    cv2.ipp.setUseIPP(True)
    ref = ref.reshape(-1, 2)
    kx, ky, desc = cv2_ref.orb_compute(L, ref[:, 0].copy(), ref[:, 1].copy())
    out["cv_par_kx"], out["cv_par_ky"], out["cv_par_desc"] = kx, ky, desc
    np.savez_compressed(os.path.join(HERE, "subpix.npz"), **out)
    print("wrote subpix.npz:", len(pts), "points,", int((np.abs(out["cv_subpix"] - out["cv_subpix_ipp"]).max(1) > 0).sum()),
          "differ between the IPP and plain paths; parallel detector keeps", len(kx))


if __name__ == "__main__":
    main()
