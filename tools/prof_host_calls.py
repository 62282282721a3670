#!/usr/bin/env python
"""Per-call latency of the host-pointer entries the C++ adapter binds, measured through ctypes on one
B200 at 752 x 480 with ~1 118 keypoints: the LK entries with both frames in the pyramid cache / one new frame per call (what
the reference's call pattern produces) / two new frames, the grid detector, the matcher.

    python tools/prof_host_calls.py        -> one JSON line (microseconds per call)
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from zenslam_b200 import synthetic as syn
    from zenslam_b200._lib import LK_GET_MIN_EIGENVALS, LkParams, check, lib
    from zenslam_b200.runtime import Context
    ctx = Context(0)
    L = lib()
    w, h, cap = 752, 480, 1410
    seq, _ = syn.stereo_sequence(w, h, 40, 8800, subpixel=True)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    prm = LkParams(31, 31, 3, 99, 0.001, LK_GET_MIN_EIGENVALS, 1e-4)
    x = np.empty(cap, np.float32); y = np.empty(cap, np.float32); r = np.empty(cap, np.float32)
    d = np.empty((cap, 32), np.uint8); d2 = np.empty((cap, 32), np.uint8); n = C.c_int(0); n2 = C.c_int(0)
    frames = [np.ascontiguousarray(seq[i, 0]) for i in range(40)]
    right = np.ascontiguousarray(seq[0, 1])

    def detect(img, desc, cnt):
        check(L.zs_detect_keypoints_grid_host(ctx._h, p(img), w, h, w, 16, 16, 10, None, p(x), p(y), p(r), p(desc), C.byref(cnt)))

    detect(right, d2, n2)
    detect(frames[0], d, n)
    pts = np.ascontiguousarray(np.stack([x[:n.value], y[:n.value]], 1))
    out = np.empty((cap, 2), np.float32); keep = np.empty(cap, np.uint8); st = np.empty(cap, np.uint8); err = np.empty(cap, np.float32)
    qi = np.empty(cap, np.int32); ti = np.empty(cap, np.int32); md = np.empty(cap, np.float32); nm = C.c_int(0)

    def track(a, b, P):
        check(L.zs_track_keypoints_host(ctx._h, p(a), p(b), w, h, w, p(P), None, len(P), C.byref(prm), 1.0, p(out), None, None, p(keep)))

    def lk(a, b, P):
        check(L.zs_calc_optical_flow_pyr_lk_host(ctx._h, p(a), p(b), w, h, w, p(P), p(out), len(P), p(st), p(err), C.byref(prm)))

    def timeit(fn, reps=40):
        for i in range(6):
            fn(i)
        ctx.synchronize()
        t0 = time.perf_counter()
        for i in range(reps):
            fn(6 + i)
        return round((time.perf_counter() - t0) / reps * 1e6, 1)

    res = {"unit": "us per call", "keypoints": int(n.value), "frame": "%dx%d" % (w, h),
           "track_fb_both_frames_cached": timeit(lambda i: track(frames[0], frames[1], pts)),
           "track_fb_both_frames_cached_8_points": timeit(lambda i: track(frames[0], frames[1], pts[:8])),
           "track_fb_one_new_frame": timeit(lambda i: track(frames[i % 40], frames[(i + 1) % 40], pts)),
           "track_fb_two_new_frames": timeit(lambda i: track(frames[(2 * i) % 40], frames[(2 * i + 1) % 40], pts)),
           "lk_both_frames_cached": timeit(lambda i: lk(frames[0], frames[1], pts)),
           "lk_one_new_frame": timeit(lambda i: lk(frames[i % 40], frames[(i + 1) % 40], pts)),
           "lk_two_new_frames": timeit(lambda i: lk(frames[(2 * i) % 40], frames[(2 * i + 1) % 40], pts)),
           "detect_grid": timeit(lambda i: detect(frames[i % 40], d, n)),
           "match_knn_hamming": timeit(lambda i: check(L.zs_match_host(ctx._h, p(d), n.value, p(d2), n2.value, 32, 0, 0, 0.8, p(qi), p(ti), p(md), C.byref(nm))))}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
