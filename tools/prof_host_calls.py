import ctypes as C, time, sys, os
import numpy as np
sys.path.insert(0, "/root/repo")
from zenslam_b200 import synthetic as syn
from zenslam_b200._lib import LK_GET_MIN_EIGENVALS, LkParams, check, lib
from zenslam_b200.runtime import Context
ctx = Context(0); L = lib()
w, h = 752, 480
seq, _ = syn.stereo_sequence(w, h, 40, 8800, subpixel=True)
p = lambda a: a.ctypes.data_as(C.c_void_p)
prm = LkParams(31, 31, 3, 99, 0.001, LK_GET_MIN_EIGENVALS, 1e-4)
cap = 1410
x = np.empty(cap, np.float32); y = np.empty(cap, np.float32); r = np.empty(cap, np.float32); d = np.empty((cap, 32), np.uint8); n = C.c_int(0)
img0 = np.ascontiguousarray(seq[0, 0])
check(L.zs_detect_keypoints_grid_host(ctx._h, p(img0), w, h, w, 16, 16, 10, None, p(x), p(y), p(r), p(d), C.byref(n)))
pts = np.ascontiguousarray(np.stack([x[:n.value], y[:n.value]], 1))
out = np.empty((cap, 2), np.float32); keep = np.empty(cap, np.uint8)
frames = [np.ascontiguousarray(seq[i, 0]) for i in range(40)]
def track(a, b, P):
    check(L.zs_track_keypoints_host(ctx._h, p(a), p(b), w, h, w, p(P), None, len(P), C.byref(prm), 1.0, p(out), None, None, p(keep)))
def timeit(fn, reps=30):
    for i in range(5): fn(i)
    ctx.synchronize(); t0 = time.perf_counter()
    for i in range(reps): fn(5 + i)
    return (time.perf_counter() - t0) / reps * 1e6
print("track, both frames cached (hits), n=%d: %.1f us" % (len(pts), timeit(lambda i: track(frames[0], frames[1], pts))))
print("track, both frames cached, n=8: %.1f us" % timeit(lambda i: track(frames[0], frames[1], pts[:8])))
print("track, one new frame per call, n=%d: %.1f us" % (len(pts), timeit(lambda i: track(frames[i % 40], frames[(i + 1) % 40], pts))))
print("track, two new frames per call, n=8: %.1f us" % timeit(lambda i: track(frames[(2 * i) % 40], frames[(2 * i + 1) % 40], pts[:8])))
print("detect grid: %.1f us" % timeit(lambda i: check(L.zs_detect_keypoints_grid_host(ctx._h, p(frames[i % 40]), w, h, w, 16, 16, 10, None, p(x), p(y), p(r), p(d), C.byref(n)))))
