#!/usr/bin/env python
"""Throughput of the stateful keypoint_tracker::track flow on the device (zs_tracker): S independent 752x480 stereo
sequences tracked in lock-step, one zs_tracker_track_host call per time step, host images in / both keypoint maps out.

    python tools/bench_tracker.py [--sequences 1 8 32] [--frames 24]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def measure(ctx, sequences, frames=24, w=752, h=480):
    """-> {"sequences_S": {...}}: blocking host call, device-resident and pipelined host rates of zs_tracker for every S"""
    from zenslam_b200 import detection_options, slam_options, synthetic as syn, tracking_options
    from zenslam_b200._lib import TrackerResults, check, lib
    from zenslam_b200.keypoint_tracker import device_keypoint_tracker
    F = frames
    opts = slam_options(detection=detection_options(), tracking=tracking_options(filter_epipolar=False))
    base = [syn.stereo_sequence(w, h, F, 4100 + s, subpixel=True)[0] for s in range(4)]      # four distinct sequences, reused
    out = {}
    p = lambda x: x.ctypes.data_as(C.c_void_p)
    for S in sequences:
        trk = device_keypoint_tracker(opts, ctx, w, h, sequences=S)
        cap = trk.cap
        n = np.zeros((S, 2), np.int32); nxt = np.zeros(S, np.int32)
        idx = [np.empty((S, cap), np.int32) for _ in range(2)]; xy = [np.empty((S, cap, 2), np.float32) for _ in range(2)]
        resp = [np.empty((S, cap), np.float32) for _ in range(2)]; desc = [np.empty((S, cap, 32), np.uint8) for _ in range(2)]
        r = TrackerResults(); r.cap = cap; r.n = p(n).value; r.next_index = p(nxt).value
        for c in range(2):
            r.index[c] = p(idx[c]).value; r.xy[c] = p(xy[c]).value; r.response[c] = p(resp[c]).value; r.desc[c] = p(desc[c]).value
        L = [np.ascontiguousarray(np.stack([base[s % 4][t, 0] for s in range(S)])) for t in range(F)]
        R = [np.ascontiguousarray(np.stack([base[s % 4][t, 1] for s in range(S)])) for t in range(F)]
        warm = F // 3
        for t in range(warm):
            check(lib().zs_tracker_track_host(trk._h, p(L[t]), p(R[t]), w, w * h, C.byref(r)))
        t0 = time.perf_counter()
        for t in range(warm, F):
            check(lib().zs_tracker_track_host(trk._h, p(L[t]), p(R[t]), w, w * h, C.byref(r)))
        dt = (time.perf_counter() - t0) / (F - warm)
        # frames resident on the device, maps left there (one download at the end): the device-side rate of the same flow
        import torch
        Ld = [torch.from_numpy(x).cuda() for x in L]; Rd = [torch.from_numpy(x).cuda() for x in R]
        trk2 = device_keypoint_tracker(opts, ctx, w, h, sequences=S)
        for t in range(warm):
            check(lib().zs_tracker_track(trk2._h, C.c_void_p(Ld[t].data_ptr()), C.c_void_p(Rd[t].data_ptr()), w, w * h))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(warm, F):
            check(lib().zs_tracker_track(trk2._h, C.c_void_p(Ld[t].data_ptr()), C.c_void_p(Rd[t].data_ptr()), w, w * h))
        e1.record()
        torch.cuda.synchronize()
        ddt = e0.elapsed_time(e1) / (F - warm) / 1e3
        # pipelined host path: pinned frames, two steps in flight (copy-in | step | copy-out overlap)
        Lp = [torch.from_numpy(x).pin_memory() for x in L]; Rp = [torch.from_numpy(x).pin_memory() for x in R]
        trk3 = device_keypoint_tracker(opts, ctx, w, h, sequences=S)
        pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
        pb = []
        for k in range(2):
            d = dict(n=pin((S, 2), torch.int32), nxt=pin((S,), torch.int32), idx=[pin((S, cap), torch.int32) for _ in range(2)],
                     xy=[pin((S, cap, 2), torch.float32) for _ in range(2)], resp=[pin((S, cap), torch.float32) for _ in range(2)],
                     desc=[pin((S, cap, 32), torch.uint8) for _ in range(2)])
            rr = TrackerResults(); rr.cap = cap; rr.n = d["n"].data_ptr(); rr.next_index = d["nxt"].data_ptr()
            for c in range(2):
                rr.index[c] = d["idx"][c].data_ptr(); rr.xy[c] = d["xy"][c].data_ptr(); rr.response[c] = d["resp"][c].data_ptr()
                rr.desc[c] = d["desc"][c].data_ptr()
            d["r"] = rr
            pb.append(d)
        sub = lambda t: check(lib().zs_tracker_submit_host(trk3._h, C.c_void_p(Lp[t].data_ptr()), C.c_void_p(Rp[t].data_ptr()), w, w * h,
                                                           C.byref(pb[t & 1]["r"])))
        for t in range(warm):
            sub(t)
            if t >= 1:
                check(lib().zs_tracker_wait(trk3._h))
        check(lib().zs_tracker_wait(trk3._h))
        t0 = time.perf_counter()
        sub(warm)
        for t in range(warm + 1, F):
            sub(t)
            check(lib().zs_tracker_wait(trk3._h))
        check(lib().zs_tracker_wait(trk3._h))
        pdt = (time.perf_counter() - t0) / (F - warm)
        assert np.array_equal(pb[(F - 1) & 1]["n"].numpy(), n)
        trk3.close()
        n2 = np.zeros((S, 2), np.int32)
        r2 = TrackerResults(); r2.cap = cap; r2.n = p(n2).value
        check(lib().zs_tracker_download(trk2._h, C.byref(r2)))
        assert np.array_equal(n2, n)
        out["sequences_%d" % S] = {"ms_per_step": dt * 1e3, "stereo_frames_per_s": S / dt, "keypoints_per_camera": float(n.mean()),
                                   "device_resident_ms_per_step": ddt * 1e3, "device_resident_stereo_frames_per_s": S / ddt,
                                   "pipelined_ms_per_step": pdt * 1e3, "pipelined_stereo_frames_per_s": S / pdt}
        trk.close(); trk2.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sequences", type=int, nargs="+", default=[1, 8, 32])
    ap.add_argument("--frames", type=int, default=24)
    ap.add_argument("--width", type=int, default=752)
    ap.add_argument("--height", type=int, default=480)
    a = ap.parse_args()
    from zenslam_b200.runtime import Context
    print(json.dumps(measure(Context(0), a.sequences, a.frames, a.width, a.height)))


if __name__ == "__main__":
    main()
