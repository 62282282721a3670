#!/usr/bin/env python
"""Micro-benchmark of the multi-scale ORB detector (`feature: ORB`): cv::ORB::create(500, 1.2f, 8, 31, 0, 2, HARRIS_SCORE, 31,
thr)->detect + cv::ORB::create()->compute on 752x480 frames, device-resident, against cv2 on one host core.

    python tools/bench_orb_detect.py [--batch 16] [--reps 20]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--width", type=int, default=752)
    ap.add_argument("--height", type=int, default=480)
    a = ap.parse_args()
    import torch
    from zenslam_b200 import synthetic as syn
    from zenslam_b200.runtime import Context, OrbDetector
    ctx = Context(0)
    imgs = np.stack([syn.stereo_pair(a.width, a.height, 9100 + i)[0] for i in range(a.batch)])
    out = {}
    for b in sorted({1, a.batch}):
        det = OrbDetector(ctx, a.width, a.height, b, fast_threshold=10)
        d = torch.from_numpy(imgs[:b]).cuda()
        for _ in range(3):
            r = det.detect_and_compute(d)
        torch.cuda.synchronize()
        l0 = ctx.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            r = det.detect_and_compute(d)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        out["batch_%d" % b] = {"ms_per_call": ms, "images_per_s": b / ms * 1e3, "keypoints_per_image": float(r["n"].float().mean()),
                               "launches_per_call": (ctx.launches - l0) / a.reps}
        det.close()
    try:
        import cv2
        cv2.setNumThreads(1)
        orb, desc = cv2.ORB_create(500, 1.2, 8, 31, 0, 2, cv2.ORB_HARRIS_SCORE, 31, 10), cv2.ORB_create()
        t0 = time.perf_counter()
        for i in range(min(8, a.batch)):
            k = orb.detect(imgs[i], None)
            desc.compute(imgs[i], k)
        out["cv2_one_core"] = {"images_per_s": min(8, a.batch) / (time.perf_counter() - t0), "version": cv2.__version__}
    except Exception as e:
        out["cv2_one_core"] = {"error": str(e)[:100]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
