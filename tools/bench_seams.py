#!/usr/bin/env python
"""The hot path the way the reference's per-frame `slam_thread` can actually call it: ONE stereo frame at a time through the
host-pointer entries that sit behind the three seams --

    2 x zs_detect_keypoints_grid_host   (keypoint_detector::detect_keypoints, keypoint_tracker.cpp:53,69)
    4 x zs_track_keypoints_host         (keypoint_tracker::track_keypoints: 2 temporal + 2 stereo, :47,50,60-67,76-83)
    1 x zs_match_host                   (matcher::match_keypoints, KNN + ratio)

every call with host buffers in and out (H2D / D2H inside), strictly sequential, no batching.

    python tools/bench_seams.py [--frames 40] [--width 752 --height 480]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def measure(ctx, w=752, h=480, cell=(16, 16), thr=10, win=(31, 31), max_level=3, klt_thr=1.0, ratio=0.8, parallel_grid=False,
            frames=40, warm=8, seed=8800):
    from zenslam_b200 import synthetic as syn
    from zenslam_b200._lib import LK_GET_MIN_EIGENVALS, LkParams, check, lib
    L = lib()
    seq, _ = syn.stereo_sequence(w, h, frames + warm + 1, seed, subpixel=True)
    cap = max(1, (w // cell[0]) * (h // cell[1]))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    detect = L.zs_detect_keypoints_parallel_host if parallel_grid else L.zs_detect_keypoints_grid_host
    prm = LkParams(win[0], win[1], max_level, 99, 0.001, LK_GET_MIN_EIGENVALS, 1e-4)

    def kp_buffers():
        return dict(x=np.empty(cap, np.float32), y=np.empty(cap, np.float32), r=np.empty(cap, np.float32),
                    d=np.empty((cap, 32), np.uint8), n=C.c_int(0))

    def det(img, b):
        check(detect(ctx._h, p(img), w, h, w, cell[0], cell[1], thr, None, p(b["x"]), p(b["y"]), p(b["r"]), p(b["d"]), C.byref(b["n"])))
        n = b["n"].value
        return np.ascontiguousarray(np.stack([b["x"][:n], b["y"][:n]], 1))

    out_pts = np.empty((cap, 2), np.float32); keep = np.empty(cap, np.uint8)

    def track(a, b, pts):
        n = len(pts)
        if n:
            check(L.zs_track_keypoints_host(ctx._h, p(a), p(b), w, h, w, p(pts), None, n, C.byref(prm), float(klt_thr), p(out_pts),
                                            None, None, p(keep)))
        return int(keep[:n].sum())

    qi = np.empty(cap, np.int32); ti = np.empty(cap, np.int32); md = np.empty(cap, np.float32); nm = C.c_int(0)
    bl, br = kp_buffers(), kp_buffers()
    prev = None
    t_stage = dict(detect=0.0, track=0.0, match=0.0)
    kept = matches = kps = 0
    t0 = None
    for t in range(frames + warm + 1):
        if t == warm + 1:
            ctx.synchronize()
            t0 = time.perf_counter(); t_stage = dict(detect=0.0, track=0.0, match=0.0); kept = matches = kps = 0
        Lf, Rf = np.ascontiguousarray(seq[t, 0]), np.ascontiguousarray(seq[t, 1])
        a = time.perf_counter()
        kl, kr = det(Lf, bl), det(Rf, br)
        b = time.perf_counter()
        if prev is not None:
            kept += track(prev[0], Lf, prev[2]) + track(prev[1], Rf, prev[3])
        kept += track(Lf, Rf, kl) + track(Rf, Lf, kr)
        c = time.perf_counter()
        if len(kl) and len(kr):
            check(L.zs_match_host(ctx._h, p(bl["d"]), len(kl), p(br["d"]), len(kr), 32, 0, 0, float(ratio), p(qi), p(ti), p(md), C.byref(nm)))
            matches += nm.value
        d = time.perf_counter()
        t_stage["detect"] += b - a; t_stage["track"] += c - b; t_stage["match"] += d - c
        kps += len(kl) + len(kr)
        prev = (Lf, Rf, kl, kr)
    dt = (time.perf_counter() - t0) / frames
    return {"stereo_frames_per_s": 1.0 / dt, "ms_per_stereo_frame": dt * 1e3, "frames": frames,
            "calls_per_frame": "2 x zs_detect_keypoints_%s_host + 4 x zs_track_keypoints_host + 1 x zs_match_host, host buffers, "
                               "sequential" % ("parallel" if parallel_grid else "grid"),
            "ms_detect": t_stage["detect"] / frames * 1e3, "ms_track": t_stage["track"] / frames * 1e3,
            "ms_match": t_stage["match"] / frames * 1e3, "keypoints_per_image": kps / (2.0 * frames),
            "tracks_kept_per_frame": kept / float(frames), "matches_per_frame": matches / float(frames),
            "h2d_bytes_per_frame": 2 * w * h + 8 * w * h, "note": "the LK entries see each frame through the content-keyed "
            "pyramid cache (zs_host.cu): 4 of the 8 image uploads per frame are hits"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=40)
    ap.add_argument("--width", type=int, default=752)
    ap.add_argument("--height", type=int, default=480)
    a = ap.parse_args()
    from zenslam_b200.runtime import Context
    print(json.dumps(measure(Context(0), a.width, a.height, frames=a.frames)))


if __name__ == "__main__":
    main()
