"""Golden fixture for the multi-scale ORB detector (`feature: ORB`):
cv::ORB::create(500, 1.2f, 8, 31, 0, 2, HARRIS_SCORE, 31, fast_threshold)->detect(image, kps, mask) followed by
cv::ORB::create()->compute(image, kps, desc) -- keypoint_detector_simple.cpp:17,27,49,54.

Run here (cv2 4.13.0 importable):  python tests/golden/make_golden_orb.py  -> tests/golden/orb_detect.npz
OpenCV's KeyPointsFilter::retainBest permutes the keypoints with std::nth_element, so the order of cv2's output is
standard-library specific; the fixture stores the keypoint SET in canonical order (octave, y, x).  Also stored:
cv2.resize(..., INTER_LINEAR_EXACT) of the first image at the first two pyramid sizes and a table of cv2.fastAtan2.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import cv2  # noqa: E402

from zenslam_b200 import synthetic as syn  # noqa: E402


def detect(img, mask, thr):
    kps = cv2.ORB_create(500, 1.2, 8, 31, 0, 2, cv2.ORB_HARRIS_SCORE, 31, thr).detect(img, mask)
    kps2, desc = cv2.ORB_create().compute(img, kps)
    assert len(kps2) == len(kps)
    order = [i for _, _, _, i in sorted((k.octave, k.pt[1], k.pt[0], i) for i, k in enumerate(kps2))]
    return dict(x=np.array([kps2[i].pt[0] for i in order], np.float32), y=np.array([kps2[i].pt[1] for i in order], np.float32),
                size=np.array([kps2[i].size for i in order], np.float32),
                angle=np.array([kps2[i].angle for i in order], np.float32),
                response=np.array([kps2[i].response for i in order], np.float32),
                octave=np.array([kps2[i].octave for i in order], np.int32), desc=desc[order])


def main():
    cv2.setNumThreads(1)
    out = dict(cv2_version=cv2.__version__)
    cases = [("a", 376, 240, 5001, 10, False), ("b", 376, 240, 5001, 10, True), ("c", 320, 200, 5002, 20, True)]
    for name, w, h, seed, thr, masked in cases:
        img, _ = syn.stereo_pair(w, h, seed)
        mask = None
        if masked:
            rng = np.random.default_rng(seed)
            mask = np.full((h, w), 255, np.uint8)
            for _ in range(120):
                cv2.circle(mask, (int(rng.integers(0, w)), int(rng.integers(0, h))), 8, 0, -1)
            out[name + "_mask"] = mask
        out[name + "_img"] = img
        out[name + "_thr"] = thr
        for k, v in detect(img, mask, thr).items():
            out[name + "_" + k] = v
        print(name, w, h, "thr", thr, "masked" if masked else "", "->", len(out[name + "_x"]), "keypoints, octaves",
              np.bincount(out[name + "_octave"]).tolist())
    img = out["a_img"]
    out["resize_313x200"] = cv2.resize(img, (313, 200), interpolation=cv2.INTER_LINEAR_EXACT)
    out["resize_261x167"] = cv2.resize(out["resize_313x200"], (261, 167), interpolation=cv2.INTER_LINEAR_EXACT)
    rng = np.random.default_rng(11)
    yx = rng.integers(-300000, 300000, (4096, 2)).astype(np.float32)
    yx[:8] = [[0, 1], [1, 0], [0, -1], [-1, 0], [5, 5], [-5, 5], [0, 0], [3, -7]]
    out["atan_yx"] = yx
    out["atan_deg"] = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in yx], np.float32)
    np.savez_compressed(os.path.join(HERE, "orb_detect.npz"), **out)


if __name__ == "__main__":
    main()
