"""GPU: the reference-interface mirrors (detector / matcher / pyr_lk) and the batched front-end vs the oracle."""
import numpy as np
import pytest

import oracle
from zenslam_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from zenslam_b200.runtime import Context
    c = Context()
    yield c
    c.close()


def oracle_detect(img, cell, thr, occ=None):
    x, y, s = oracle.grid_detect(img, cell, thr, occ)
    kept, desc = oracle.orb_compute(img, x, y)
    return np.stack([x[kept], y[kept]], 1).astype(np.float32), s[kept].astype(np.float32), desc


def test_keypoint_detector_grid_mirror(ctx):
    from zenslam_b200 import detection_options, keypoint
    from zenslam_b200.detection import keypoint_detector_grid
    L, R = syn.stereo_pair(752, 480, 1001)
    det = keypoint_detector_grid(detection_options(), ctx)
    keypoint.index_next = 100
    kps = det.detect_keypoints(L, {})
    xy, resp, desc = oracle_detect(L, (16, 16), 10)
    assert len(kps) == len(xy) and [k.index for k in kps] == list(range(100, 100 + len(kps)))
    assert np.array_equal(np.array([k.pt for k in kps], np.float32), xy)
    assert np.array_equal(np.array([k.response for k in kps], np.float32), resp)
    assert np.array_equal(np.stack([k.descriptor for k in kps]), desc)
    assert all(k.size == 7.0 and k.angle == -1.0 and k.octave == 0 and k.class_id == -1 for k in kps)
    # occupancy from existing keypoints (keypoint_detector_grid.cpp:48-63)
    existing = {k.index: k for k in kps[::3]}
    kps2 = det.detect_keypoints(L, existing)
    occ = np.zeros((480 // 16, 752 // 16), np.uint8)
    for k in existing.values():
        occ[int(k.pt[1]) // 16, int(k.pt[0]) // 16] = 1
    xy2, _, desc2 = oracle_detect(L, (16, 16), 10, occ)
    assert np.array_equal(np.array([k.pt for k in kps2], np.float32), xy2)
    assert np.array_equal(np.stack([k.descriptor for k in kps2]), desc2)


def test_keypoint_detector_simple_mirror(ctx):
    from zenslam_b200 import detection_options
    from zenslam_b200.detection import keypoint_detector_simple
    L, _ = syn.stereo_pair(320, 240, 1002)
    det = keypoint_detector_simple(detection_options(fast_threshold=20), ctx)
    kps = det.detect_keypoints(L, {})
    x, y, s = oracle.fast_detect(L, 20)
    kept, desc = oracle.orb_compute(L, x, y)
    assert len(kps) == len(kept)
    assert np.array_equal(np.array([k.pt for k in kps], np.float32), np.stack([x[kept], y[kept]], 1).astype(np.float32))
    assert np.array_equal(np.stack([k.descriptor for k in kps]), desc)
    # mask: discs of radius min(cell)/2 around existing keypoints (keypoint_detector_simple.cpp:41-45)
    existing = kps[::5]
    kps2 = det.detect_keypoints(L, existing)
    r = 8
    drop = np.zeros(len(x), bool)
    for k in existing:
        cx, cy = int(round(k.pt[0])), int(round(k.pt[1]))
        drop |= (x - cx) ** 2 + (y - cy) ** 2 <= r * r
    kept2, desc2 = oracle.orb_compute(L, x[~drop], y[~drop])
    assert len(kps2) == len(kept2)
    assert np.array_equal(np.stack([k.descriptor for k in kps2]), desc2) if len(kept2) else True


@pytest.mark.parametrize("mode", ["KNN", "BRUTE"])
def test_matcher_mirror(ctx, mode):
    from zenslam_b200 import detection_options, keypoint, slam_options
    from zenslam_b200.detection import keypoint_detector_grid
    from zenslam_b200.matching import matcher
    L, R = syn.stereo_pair(752, 480, 1003)
    det = keypoint_detector_grid(detection_options(), ctx)
    keypoint.index_next = 0
    k0 = det.detect_keypoints(L, {}); k1 = det.detect_keypoints(R, {})
    m = matcher(slam_options(matcher=mode, matcher_ratio=0.8), True, ctx)
    got = m.match_keypoints(k0, k1)
    d0 = np.stack([k.descriptor for k in k0]); d1 = np.stack([k.descriptor for k in k1])
    if mode == "KNN":
        idx, dist = oracle.match_hamming_knn2(d0, d1)
        q, t, d = oracle.ratio_test(idx, dist.astype(np.float32), 0.8)
    else:
        q, t, d = oracle.match_hamming_cross(d0, d1)
    assert len(got) == len(q) and len(got) > 100
    assert [g.queryIdx for g in got] == [k0[i].index for i in q]
    assert [g.trainIdx for g in got] == [k1[i].index for i in t]
    assert np.array_equal(np.array([g.distance for g in got], np.float32), d.astype(np.float32))
    # map overload: keypoints whose index exists in the other set are skipped (matcher.cpp:21-53)
    m0 = {k.index: k for k in k0}; m1 = {k.index: k for k in k1}
    shared = k0[5]
    m1[shared.index] = shared
    got2 = m.match_keypoints(m0, m1)
    assert all(g.queryIdx != shared.index and g.trainIdx != shared.index for g in got2)
    assert m.match_keypoints([], k1) == [] and m.match_keypoints({}, m1) == []


def test_pyr_lk_mirror_and_fb(ctx):
    from zenslam_b200 import keypoint, tracking_options
    from zenslam_b200.tracking import create_cuda_pyr_lk, track_keypoints
    lk = create_cuda_pyr_lk(ctx)
    assert lk is not None
    seq, _ = syn.stereo_sequence(752, 480, 2, 1004, subpixel=True)
    A, B = seq[0, 0], seq[1, 0]
    rng = np.random.default_rng(2)
    pts = np.stack([rng.uniform(-5, 757, 500), rng.uniform(-5, 485, 500)], 1).astype(np.float32)
    p1, st, err = lk.calc_optical_flow_pyr_lk([A], [B], pts, None, (31, 31), 3, (99, 0.001), 8, 1e-4)
    PA, PB = oracle.Pyramid(A, (31, 31), 3), oracle.Pyramid(B, (31, 31), 3)
    o1, os_, oe = oracle.lk_track(PA, PB, pts, None)
    assert np.array_equal(st, os_) and np.array_equal(p1, o1) and np.array_equal(err, oe)
    kps = [keypoint(pt=(float(x), float(y)), index=i) for i, (x, y) in enumerate(pts)]
    tracked = track_keypoints(lk, A, B, kps, tracking_options())
    ob, osb, _ = oracle.lk_track(PB, PA, o1, None)
    keep = oracle.fb_check(pts, ob, os_, osb, 1.0)
    assert [k.index for k in tracked] == list(np.nonzero(keep)[0])
    assert np.array_equal(np.array([k.pt for k in tracked], np.float32), o1[keep])
    # the fused host entry (forward + backward + gate in one device call) gives the same survivors and positions,
    # with and without an initial flow
    fused = track_keypoints(lk, A, B, kps, tracking_options(), fused=True)
    assert [k.index for k in fused] == [k.index for k in tracked] and [k.pt for k in fused] == [k.pt for k in tracked]
    pred = (pts + rng.uniform(-4, 4, pts.shape)).astype(np.float32)
    t2 = track_keypoints(lk, A, B, kps, tracking_options(), predicted_points=pred)
    f2 = track_keypoints(lk, A, B, kps, tracking_options(), predicted_points=pred, fused=True)
    assert len(t2) > 100 and [k.index for k in f2] == [k.index for k in t2] and [k.pt for k in f2] == [k.pt for k in t2]
    # empty input (keypoint_tracker.cpp:122)
    assert track_keypoints(lk, A, B, [], tracking_options()) == []
    e1, es, ee = lk.calc_optical_flow_pyr_lk(A, B, np.zeros((0, 2), np.float32), None, (31, 31), 3)
    assert e1.shape == (0, 2)


def test_keypoint_tracker_track_flow_vs_oracle(ctx):
    """keypoint_tracker::track (keypoint_tracker.cpp:41-105) over several frames, state carried from frame to frame:
    temporal tracks, occupancy-aware detection, stereo tracks of the keypoints the other camera lacks, index-keyed maps.
    The CUDA-backed mirror and the same host glue over the oracle (C restatement of OpenCV) must produce identical maps."""
    import dataclasses
    from zenslam_b200 import detection_options, keypoint, slam_options, tracking_options
    from zenslam_b200.keypoint_tracker import keypoint_tracker, stereo_frame
    from zenslam_b200.tracking import create_cuda_pyr_lk, pyr_lk

    class oracle_pyr_lk(pyr_lk):
        def calc_optical_flow_pyr_lk(self, prev_pyramid, next_pyramid, prev_points, next_points, win_size, max_level,
                                     criteria=(99, 0.001), flags=8, min_eig_threshold=1e-4):
            P0, P1 = oracle.Pyramid(prev_pyramid, win_size, max_level), oracle.Pyramid(next_pyramid, win_size, max_level)
            return oracle.lk_track(P0, P1, np.asarray(prev_points, np.float32), next_points, win_size, max_level)

    class oracle_detector:
        def __init__(self, opt):
            self.opt = opt

        def detect_keypoints(self, image, existing):
            cw, ch = self.opt.cell_size
            h, w = image.shape
            occ = np.zeros((h // ch, w // cw), np.uint8)
            for kp in (existing.values() if existing else []):
                gx, gy = int(int(kp.pt[0]) / cw), int(int(kp.pt[1]) / ch)   # C++ truncation (toward zero) twice
                if 0 <= gx < occ.shape[1] and 0 <= gy < occ.shape[0]:
                    occ[gy, gx] = 1
            x, y, s = oracle.grid_detect(image, (cw, ch), self.opt.fast_threshold, occ)
            kept, desc = oracle.orb_compute(image, x, y)
            out = []
            for i, k in enumerate(kept):
                out.append(keypoint(pt=(float(x[k]), float(y[k])), response=float(s[k]), index=keypoint.index_next, descriptor=desc[i]))
                keypoint.index_next += 1
            return out

    w, h, frames = 376, 240, 4
    seq, _ = syn.stereo_sequence(w, h, frames, 1020, subpixel=True)
    opts = slam_options(matcher="KNN", detection=detection_options(), tracking=tracking_options(filter_epipolar=False))

    def run(tracker):
        keypoint.index_next = 0
        prev = stereo_frame((seq[0, 0], seq[0, 1]))
        out = []
        for t in range(frames):
            cur = stereo_frame((seq[t, 0], seq[t, 1]))
            k0, k1 = tracker.track(prev, cur)
            out.append((k0, k1))
            prev = dataclasses.replace(cur, keypoints=(k0, k1))
        return out

    got = run(keypoint_tracker(opts, ctx, create_cuda_pyr_lk(ctx)))
    ref = run(keypoint_tracker(opts, ctx, oracle_pyr_lk(), detector=oracle_detector(opts.detection)))
    for t in range(frames):
        for cam in range(2):
            g, r = got[t][cam], ref[t][cam]
            assert sorted(g) == sorted(r), (t, cam)
            assert all(g[i].pt == r[i].pt and np.array_equal(g[i].descriptor, r[i].descriptor) for i in g), (t, cam)
    n_last = len(got[-1][0])
    shared = len(set(got[-1][0]) & set(got[-1][1]))
    carried = len(set(got[-1][0]) & set(got[0][0]))
    assert n_last > 200 and shared > 100 and carried > 50, (n_last, shared, carried)     # tracks survive, stereo pairs exist
    with pytest.raises(NotImplementedError):
        keypoint_tracker(slam_options(), ctx, create_cuda_pyr_lk(ctx))                  # filter_epipolar on, no RANSAC callable


@pytest.mark.parametrize("w,h,cell,algorithm", [(376, 240, (16, 16), "GRID"), (640, 400, (32, 32), "GRID"),
                                                 (376, 240, (16, 16), "PARALLEL_GRID")])
def test_device_tracker_equals_track_mirror(ctx, w, h, cell, algorithm):
    """zs_tracker (maps, previous pyramids and index counter on the device, one call per stereo frame) against the
    keypoint_tracker.track mirror, which is itself checked against the oracle: same index sets, positions, responses and
    descriptors in both cameras on every frame, same keypoint::index_next"""
    import dataclasses
    from zenslam_b200 import detection_options, keypoint, slam_options, tracking_options
    from zenslam_b200.keypoint_tracker import device_keypoint_tracker, keypoint_tracker, stereo_frame
    from zenslam_b200.tracking import create_cuda_pyr_lk
    frames = 6
    seq, _ = syn.stereo_sequence(w, h, frames, 1040 + w, subpixel=True)
    opts = slam_options(matcher="KNN", detection=detection_options(cell_size=cell, algorithm=algorithm),
                        tracking=tracking_options(filter_epipolar=False))
    keypoint.index_next = 0
    host = keypoint_tracker(opts, ctx, create_cuda_pyr_lk(ctx))
    prev = stereo_frame((seq[0, 0], seq[0, 1]))
    want = []
    for t in range(frames):
        cur = stereo_frame((seq[t, 0], seq[t, 1]))
        k0, k1 = host.track(prev, cur)
        want.append((k0, k1, keypoint.index_next))
        prev = dataclasses.replace(cur, keypoints=(k0, k1))
    keypoint.index_next = 0
    dev_trk = device_keypoint_tracker(opts, ctx, w, h)
    for t in range(frames):
        g0, g1 = dev_trk.track(seq[t, 0], seq[t, 1])
        for cam, (g, r) in enumerate(((g0, want[t][0]), (g1, want[t][1]))):
            assert list(g) == sorted(r), (t, cam, len(g), len(r))
            for i in g:
                assert g[i].pt == r[i].pt and g[i].response == r[i].response and np.array_equal(g[i].descriptor, r[i].descriptor), (t, cam, i)
        assert keypoint.index_next == want[t][2]
    assert len(set(g0) & set(want[0][0])) > 30 and len(set(g0) & set(g1)) > 100           # tracks survive, stereo pairs exist
    dev_trk.close()


def test_device_tracker_batched_sequences(ctx):
    """three independent stereo sequences tracked in lock-step by one zs_tracker: every sequence must equal its own
    single-sequence run (maps, positions, descriptors, index counters), including per-sequence predictions"""
    from zenslam_b200 import detection_options, keypoint, slam_options, tracking_options
    from zenslam_b200.keypoint_tracker import device_keypoint_tracker
    w, h, frames, S = 376, 240, 5, 3
    seqs = [syn.stereo_sequence(w, h, frames, 1070 + s, subpixel=True)[0] for s in range(S)]
    opts = slam_options(matcher="KNN", detection=detection_options(), tracking=tracking_options(filter_epipolar=False))

    def pred_for(maps, s):
        return {i: (k.pt[0] + 0.75 * (s + 1), k.pt[1] + 0.5) for i, k in maps.items() if i % 4 == s}

    want = []
    for s in range(S):
        keypoint.index_next = 0
        one = device_keypoint_tracker(opts, ctx, w, h)
        per_frame, last = [], ({}, {})
        for t in range(frames):
            one.set_predictions(0, pred_for(last[0], s))
            last = one.track(seqs[s][t, 0], seqs[s][t, 1])
            per_frame.append((last, one.next_index[0]))
        want.append(per_frame)
        one.close()
    multi = device_keypoint_tracker(opts, ctx, w, h, sequences=S)
    last = [({}, {})] * S
    for t in range(frames):
        for s in range(S):
            multi.set_predictions(0, pred_for(last[s][0], s), sequence=s)
        got = multi.track_all(np.stack([seqs[s][t, 0] for s in range(S)]), np.stack([seqs[s][t, 1] for s in range(S)]))
        for s in range(S):
            (w0, w1), nxt = want[s][t]
            for g, r in ((got[s][0], w0), (got[s][1], w1)):
                assert list(g) == list(r), (t, s)
                assert all(g[i].pt == r[i].pt and np.array_equal(g[i].descriptor, r[i].descriptor) for i in g), (t, s)
            assert multi.next_index[s] == nxt
        last = got
    multi.close()
    # frames resident on the device, maps downloaded once at the end: the same final maps (no predictions this time)
    import torch
    a = device_keypoint_tracker(opts, ctx, w, h, sequences=S)
    b = device_keypoint_tracker(opts, ctx, w, h, sequences=S)
    for t in range(frames):
        L = np.stack([seqs[s][t, 0] for s in range(S)]); R = np.stack([seqs[s][t, 1] for s in range(S)])
        ha = a.track_all(L, R)
        b.track_device(torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda())
    hb = b.download()
    for s in range(S):
        for cam in range(2):
            assert list(ha[s][cam]) == list(hb[s][cam]) and all(ha[s][cam][i].pt == hb[s][cam][i].pt for i in ha[s][cam])
    assert a.next_index == b.next_index
    a.close(); b.close()


def test_device_tracker_filter_epipolar(ctx):
    """filter_epipolar (keypoint_tracker.cpp:293-341) with a caller-supplied F: the device filter after every step against
    the mirror's epipolar_filter callable evaluating |pt0^T F pt1| in the same double arithmetic; the filtered maps are
    what the next frame tracks from in both implementations"""
    import dataclasses
    from zenslam_b200 import detection_options, keypoint, slam_options, tracking_options
    from zenslam_b200.keypoint_tracker import device_keypoint_tracker, keypoint_tracker, stereo_frame
    from zenslam_b200.tracking import create_cuda_pyr_lk
    w, h, frames = 376, 240, 5
    seq, _ = syn.stereo_sequence(w, h, frames, 1110, subpixel=True)
    # rectified stereo: F = [t]_x with t along x, so that pt0^T F pt1 = y1 - y0 (plus a little skew to make every term matter)
    F = np.array([[0.0, 1e-6, 0.0], [-1e-6, 0.0, -1.0], [0.0, 1.0, 0.002]], np.float64)
    thr = 0.0027                                      # err = 0.002 + 6e-6 y on this pair (disparity 6 px): the upper half passes

    def epipolar_filter(m0, m1):
        keep = []
        for a, b in zip(m0, m1):
            x0, y0 = np.float64(np.float32(a.pt[0])), np.float64(np.float32(a.pt[1]))
            x1, y1 = np.float64(np.float32(b.pt[0])), np.float64(np.float32(b.pt[1]))
            r = [(x0 * F[0, j] + y0 * F[1, j]) + F[2, j] for j in range(3)]
            keep.append(abs((r[0] * x1 + r[1] * y1) + r[2]) < thr)
        return keep

    opts_h = slam_options(matcher="KNN", detection=detection_options(), tracking=tracking_options(filter_epipolar=True))
    keypoint.index_next = 0
    host = keypoint_tracker(opts_h, ctx, create_cuda_pyr_lk(ctx), epipolar_filter=epipolar_filter)
    prev = stereo_frame((seq[0, 0], seq[0, 1]))
    want = []
    for t in range(frames):
        cur = stereo_frame((seq[t, 0], seq[t, 1]))
        k0, k1 = host.track(prev, cur)
        want.append((k0, k1))
        prev = dataclasses.replace(cur, keypoints=(k0, k1))
    opts_d = slam_options(matcher="KNN", detection=detection_options(), tracking=tracking_options(filter_epipolar=False))
    keypoint.index_next = 0
    dev_trk = device_keypoint_tracker(opts_d, ctx, w, h)
    dropped = 0
    for t in range(frames):
        before = dev_trk.track(seq[t, 0], seq[t, 1])
        dev_trk.filter_epipolar(F, thr)
        g = dev_trk.download()[0]
        dropped += len(before[0]) - len(g[0])
        for cam in range(2):
            assert list(g[cam]) == sorted(want[t][cam]), (t, cam, len(g[cam]), len(want[t][cam]))
            assert all(g[cam][i].pt == want[t][cam][i].pt for i in g[cam])
        assert list(g[0]) == list(g[1])                       # only keypoints present in both cameras survive
    assert dropped > 0 and len(g[0]) > 20
    dev_trk.close()


def test_device_tracker_pipelined_equals_blocking(ctx):
    """zs_tracker_submit_host / zs_tracker_wait (two steps in flight on three streams, frames staged, maps snapshotted)
    must return, step for step, what the blocking call returns"""
    import ctypes as C
    import torch
    from zenslam_b200 import detection_options, slam_options, tracking_options
    from zenslam_b200._lib import TrackerResults, check, lib
    from zenslam_b200.keypoint_tracker import device_keypoint_tracker
    w, h, frames, S = 376, 240, 7, 2
    seqs = [syn.stereo_sequence(w, h, frames, 1090 + s, subpixel=True)[0] for s in range(S)]
    L = [np.ascontiguousarray(np.stack([seqs[s][t, 0] for s in range(S)])) for t in range(frames)]
    R = [np.ascontiguousarray(np.stack([seqs[s][t, 1] for s in range(S)])) for t in range(frames)]
    opts = slam_options(matcher="KNN", detection=detection_options(), tracking=tracking_options(filter_epipolar=False))
    ref = device_keypoint_tracker(opts, ctx, w, h, sequences=S)
    want = [ref.track_all(L[t], R[t]) for t in range(frames)]
    ref.close()
    trk = device_keypoint_tracker(opts, ctx, w, h, sequences=S)
    cap = trk.cap
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    Lp = [torch.from_numpy(x).pin_memory() for x in L]; Rp = [torch.from_numpy(x).pin_memory() for x in R]
    bufs = []
    for t in range(frames):
        b = dict(n=np.zeros((S, 2), np.int32), nxt=np.zeros(S, np.int32), idx=[np.zeros((S, cap), np.int32) for _ in range(2)],
                 xy=[np.zeros((S, cap, 2), np.float32) for _ in range(2)], desc=[np.zeros((S, cap, 32), np.uint8) for _ in range(2)])
        r = TrackerResults(); r.cap = cap; r.n = p(b["n"]).value; r.next_index = p(b["nxt"]).value
        for c in range(2):
            r.index[c] = p(b["idx"][c]).value; r.xy[c] = p(b["xy"][c]).value; r.desc[c] = p(b["desc"][c]).value
        b["r"] = r
        bufs.append(b)
    done = 0

    def check_step(t):
        b = bufs[t]
        for s in range(S):
            for c in range(2):
                m = want[t][s][c]
                n = int(b["n"][s, c])
                assert n == len(m) and b["idx"][c][s, :n].tolist() == list(m), (t, s, c)
                assert np.array_equal(b["xy"][c][s, :n], np.array([m[i].pt for i in m], np.float32).reshape(-1, 2))
                assert np.array_equal(b["desc"][c][s, :n], np.stack([m[i].descriptor for i in m]))

    for t in range(frames):
        check(lib().zs_tracker_submit_host(trk._h, C.c_void_p(Lp[t].data_ptr()), C.c_void_p(Rp[t].data_ptr()), w, w * h, C.byref(bufs[t]["r"])))
        assert 1 <= lib().zs_tracker_in_flight(trk._h) <= 2
        if t >= 1:
            check(lib().zs_tracker_wait(trk._h)); check_step(done); done += 1
    while lib().zs_tracker_in_flight(trk._h):
        check(lib().zs_tracker_wait(trk._h)); check_step(done); done += 1
    assert done == frames
    trk.close()
    # the same through the Python mirror's submit / wait
    py = device_keypoint_tracker(opts, ctx, w, h, sequences=S)
    got = []
    for t in range(frames):
        py.submit(Lp[t], Rp[t])
        if t >= 1:
            got.append(py.wait())
    got.append(py.wait())
    for t in range(frames):
        for s in range(S):
            for c in range(2):
                assert list(got[t][s][c]) == list(want[t][s][c]) and all(got[t][s][c][i].pt == want[t][s][c][i].pt for i in got[t][s][c])
    py.close()


def test_device_tracker_with_predicted_initial_flow(ctx):
    """temporal tracks that start from host-supplied predictions (landmark projections, keypoint_tracker.cpp:361-373):
    zs_tracker_set_predictions against the mirror's predicted_points callable, frame by frame"""
    import dataclasses
    from zenslam_b200 import detection_options, keypoint, slam_options, tracking_options
    from zenslam_b200.keypoint_tracker import device_keypoint_tracker, keypoint_tracker, stereo_frame
    from zenslam_b200.tracking import create_cuda_pyr_lk
    w, h, frames = 376, 240, 5
    seq, _ = syn.stereo_sequence(w, h, frames, 1055, subpixel=True)
    opts = slam_options(matcher="KNN", detection=detection_options(), tracking=tracking_options(filter_epipolar=False))

    def predict(kp, cam):                        # a third of the keypoints carry a prediction, offset from their position
        return (kp.pt[0] + 1.5 * (cam + 1), kp.pt[1] - 1.0) if kp.index % 3 == 0 else None

    def predicted_points(keypoints, cam):
        return np.array([predict(k, cam) or k.pt for k in keypoints], np.float32)

    keypoint.index_next = 0
    host = keypoint_tracker(opts, ctx, create_cuda_pyr_lk(ctx), predicted_points=predicted_points)
    prev = stereo_frame((seq[0, 0], seq[0, 1]))
    want = []
    for t in range(frames):
        cur = stereo_frame((seq[t, 0], seq[t, 1]))
        k0, k1 = host.track(prev, cur)
        want.append((k0, k1))
        prev = dataclasses.replace(cur, keypoints=(k0, k1))
    keypoint.index_next = 0
    dev_trk = device_keypoint_tracker(opts, ctx, w, h)
    last = ({}, {})
    for t in range(frames):
        for cam in range(2):
            dev_trk.set_predictions(cam, {i: predict(k, cam) for i, k in last[cam].items() if predict(k, cam)})
        g = dev_trk.track(seq[t, 0], seq[t, 1])
        for cam in range(2):
            assert list(g[cam]) == sorted(want[t][cam]), (t, cam)
            assert all(g[cam][i].pt == want[t][cam][i].pt for i in g[cam]), (t, cam)
        last = g
    # the predictions changed something: without them the second frame differs
    keypoint.index_next = 0
    plain = device_keypoint_tracker(opts, ctx, w, h)
    p = [plain.track(seq[t, 0], seq[t, 1]) for t in range(2)][1]
    assert any(p[0][i].pt != want[1][0][i].pt for i in set(p[0]) & set(want[1][0])) or set(p[0]) != set(want[1][0])
    dev_trk.close(); plain.close()


def test_lk_host_pyramid_cache(ctx):
    """the LK host entries keep the pyramids of the frames they saw (content-keyed): repeated frames hit, a frame whose
    bytes changed in place misses, and results never depend on the cache state"""
    import ctypes as C
    from zenslam_b200._lib import lib
    from zenslam_b200.tracking import create_cuda_pyr_lk
    lk = create_cuda_pyr_lk(ctx)
    seq, _ = syn.stereo_sequence(376, 240, 5, 1030, subpixel=True)
    rng = np.random.default_rng(9)
    pts = np.stack([rng.uniform(5, 370, 200), rng.uniform(5, 235, 200)], 1).astype(np.float32)

    def stats():
        h, m = C.c_uint64(), C.c_uint64()
        lib().zs_lk_cache_stats(ctx._h, C.byref(h), C.byref(m))
        return h.value, m.value

    def check(a, b):
        p1, st, err = lk.calc_optical_flow_pyr_lk([a], [b], pts, None, (21, 21), 2, (99, 0.001), 8, 1e-4)
        o1, os_, oe = oracle.lk_track(oracle.Pyramid(a, (21, 21), 2), oracle.Pyramid(b, (21, 21), 2), pts, None, (21, 21), 2)
        assert np.array_equal(p1, o1) and np.array_equal(st, os_) and np.array_equal(err, oe)

    frames = [np.ascontiguousarray(seq[t, 0]) for t in range(5)]
    check(frames[0], frames[1])                       # new window geometry: both miss
    h0, m0 = stats()
    check(frames[1], frames[0])                       # both hit
    check(frames[1], frames[2])                       # one hit, one miss
    h1, m1 = stats()
    assert (h1 - h0, m1 - m0) == (3, 1)
    buf = frames[2]                                   # same buffer, new content: must not be served from the cache
    buf[:] = frames[4]
    check(frames[1], buf)
    h2, m2 = stats()
    assert (h2 - h1, m2 - m1) == (1, 1)
    for t in range(5):                                # more distinct frames than slots: eviction keeps results right
        check(frames[t % 5], frames[(t + 3) % 5])
    check(frames[0], frames[0])                       # the same frame twice occupies two slots


def test_track_keylines_mirror(ctx):
    """utils::track_keylines (tracking_utils.cpp:14-143): both endpoints through the pyr_lk seam, forward + backward,
    all four statuses and both FB errors gate the keyline; survivors get new endpoints, midpoint, length, angle"""
    from zenslam_b200 import tracking_options
    from zenslam_b200.tracking import create_cuda_pyr_lk, track_keylines
    from zenslam_b200.types import keyline
    lk = create_cuda_pyr_lk(ctx)
    seq, _ = syn.stereo_sequence(752, 480, 2, 1010, subpixel=True)
    A, B = seq[0, 0], seq[1, 0]
    rng = np.random.default_rng(4)
    s = np.stack([rng.uniform(-10, 760, 300), rng.uniform(-10, 490, 300)], 1).astype(np.float32)
    e = (s + rng.uniform(-60, 60, (300, 2))).astype(np.float32)
    lines = {i: keyline(startPointX=float(s[i, 0]), startPointY=float(s[i, 1]), endPointX=float(e[i, 0]),
                        endPointY=float(e[i, 1]), index=i) for i in range(300)}
    got = track_keylines(lk, A, B, lines, tracking_options())
    PA, PB = oracle.Pyramid(A, (31, 31), 3), oracle.Pyramid(B, (31, 31), 3)
    s1, ss, _ = oracle.lk_track(PA, PB, s, None); e1, se, _ = oracle.lk_track(PA, PB, e, None)
    sb, ssb, _ = oracle.lk_track(PB, PA, s1, None); eb, seb, _ = oracle.lk_track(PB, PA, e1, None)
    keep = oracle.fb_check(s, sb, ss, ssb, 1.0) & oracle.fb_check(e, eb, se, seb, 1.0)
    assert 50 < keep.sum() < 300
    assert [k.index for k in got] == list(np.nonzero(keep)[0])
    assert np.array_equal(np.array([(k.startPointX, k.startPointY) for k in got], np.float32), s1[keep])
    assert np.array_equal(np.array([(k.endPointX, k.endPointY) for k in got], np.float32), e1[keep])
    k0 = got[0]
    assert abs(k0.lineLength - np.hypot(k0.endPointX - k0.startPointX, k0.endPointY - k0.startPointY)) < 1e-3
    assert abs(k0.pt[0] - 0.5 * (k0.startPointX + k0.endPointX)) < 1e-4
    assert track_keylines(lk, A, B, {}, tracking_options()) == []


# three batches each: the first runs eagerly, the second is captured into a CUDA graph, the third replays it.  21 x 21
# takes the register-template kernel, 45 x 45 the generic kernel (> 48 KB of dynamic shared memory: its attribute change
# is not capturable, so that front-end must fall back to eager launches without an error)
@pytest.mark.parametrize("w,h,B,cell,win,ml", [(752, 480, 3, (16, 16), (31, 31), 3), (640, 400, 2, (32, 32), (31, 31), 3),
                                               (400, 300, 2, (16, 16), (21, 21), 2), (400, 300, 1, (16, 16), (45, 45), 2)])
def test_frontend_batches_vs_oracle(ctx, w, h, B, cell, win, ml):
    from zenslam_b200 import detection_options, slam_options, tracking_options
    from zenslam_b200.frontend import StereoFrontend
    opts = slam_options(matcher="KNN", matcher_ratio=0.8, detection=detection_options(cell_size=cell),
                        tracking=tracking_options(klt_window_size=win, klt_max_level=ml))
    fe = StereoFrontend(ctx, w, h, B, opts)
    nb = 3
    seq, _ = syn.stereo_sequence(w, h, nb * B, 5000 + w, subpixel=True)
    prev = None     # (imgL, imgR, kpL, kpR)
    for b in range(nb):
        chunk = seq[b * B:(b + 1) * B]
        res = fe.process(np.ascontiguousarray(chunk[:, 0]), np.ascontiguousarray(chunk[:, 1]))
        for k in range(B):
            L, R = chunk[k, 0], chunk[k, 1]
            xyl, rl, dl = oracle_detect(L, cell, 10)
            xyr, rr, dr = oracle_detect(R, cell, 10)
            nl, nr = res["n_left"][k], res["n_right"][k]
            assert nl == len(xyl) and nr == len(xyr)
            assert np.array_equal(res["kp_left"][k, :nl], xyl) and np.array_equal(res["kp_right"][k, :nr], xyr)
            assert np.array_equal(res["resp_left"][k, :nl], rl) and np.array_equal(res["desc_left"][k, :nl], dl)
            assert np.array_equal(res["desc_right"][k, :nr], dr)
            oi, od = oracle.match_hamming_knn2(dl, dr)
            assert np.array_equal(res["match_idx"][k, :nl], oi)
            assert np.array_equal(res["match_dist"][k, :nl], od.astype(np.float32))
            rq, _, _ = oracle.ratio_test(oi, od.astype(np.float32), 0.8)
            assert np.array_equal(np.nonzero(res["match_pass"][k, :nl])[0], rq)
            PL, PR = oracle.Pyramid(L, win, ml), oracle.Pyramid(R, win, ml)
            jobs = []
            if prev is not None:
                jobs += [(0, prev[0], PL, prev[2]), (1, prev[1], PR, prev[3])]
            else:
                assert res["track_n"][0, k] == 0 and res["track_n"][1, k] == 0
            jobs += [(2, PL, PR, xyl), (3, PR, PL, xyr)]
            for kind, P0, P1, pts in jobs:
                n = len(pts)
                assert res["track_n"][kind, k] == n
                o1, os_, _ = oracle.lk_track(P0, P1, pts, None, win, ml)
                ob, osb, _ = oracle.lk_track(P1, P0, o1, None, win, ml)
                keep = oracle.fb_check(pts, ob, os_, osb, 1.0)
                assert np.array_equal(res["track_pts"][kind, k, :n], o1), (b, k, kind)
                assert np.array_equal(res["track_keep"][kind, k, :n].astype(bool), keep)
            prev = (PL, PR, xyl, xyr)
    fe.close()


@pytest.mark.parametrize("w,h,B", [(752, 480, 2), (600, 360, 3)])     # 600: exercises the unaligned unpack path
def test_frontend_pipelined_equals_blocking(ctx, w, h, B):
    """submit/wait (two batches in flight on three streams) must return exactly what the blocking call returns,
    batch for batch, including the state carried from batch to batch."""
    from zenslam_b200 import slam_options
    from zenslam_b200.frontend import StereoFrontend
    nb = 4
    seq, _ = syn.stereo_sequence(w, h, nb * B, 77 + w, subpixel=True)
    lefts = [np.ascontiguousarray(seq[b * B:(b + 1) * B, 0]) for b in range(nb)]
    rights = [np.ascontiguousarray(seq[b * B:(b + 1) * B, 1]) for b in range(nb)]
    fe = StereoFrontend(ctx, w, h, B, slam_options())
    want = [{k: v.copy() for k, v in fe.process(lefts[b], rights[b]).items()} for b in range(nb)]
    fe.close()
    fe = StereoFrontend(ctx, w, h, B, slam_options())
    got = []
    fe.submit(lefts[0], rights[0])
    for b in range(1, nb):
        fe.submit(lefts[b], rights[b])
        assert fe.in_flight == 2
        got.append({k: v.copy() for k, v in fe.wait().items()})
    got.append({k: v.copy() for k, v in fe.wait().items()})
    assert fe.in_flight == 0
    with pytest.raises(Exception):
        fe.wait()                                   # nothing in flight: error, not a hang
    for b in range(nb):
        nl = want[b]["n_left"]
        assert np.array_equal(got[b]["n_left"], nl) and np.array_equal(got[b]["n_right"], want[b]["n_right"])
        for k in range(B):
            n = nl[k]
            for key in ("kp_left", "desc_left", "match_idx", "match_pass"):
                assert np.array_equal(got[b][key][k, :n], want[b][key][k, :n]), (b, k, key)
            for kind in range(4):
                m = want[b]["track_n"][kind, k]
                assert got[b]["track_n"][kind, k] == m
                assert np.array_equal(got[b]["track_pts"][kind, k, :m], want[b]["track_pts"][kind, k, :m])
                assert np.array_equal(got[b]["track_keep"][kind, k, :m], want[b]["track_keep"][kind, k, :m])
    fe.close()


def test_keypoint_detector_parallel_mirror(ctx, golden):
    """PARALLEL_GRID: grid corners -> cornerSubPix -> ORB::compute at cvRound(pt) (keypoint_detector_parallel.cpp:40-193),
    against the cv2 fixture (IPP off) and against the oracle on a full-size frame"""
    from zenslam_b200 import detection_options, keypoint
    from zenslam_b200.detection import keypoint_detector_parallel
    g = golden("subpix")
    det = keypoint_detector_parallel(detection_options(), ctx)
    kps = det.detect_keypoints(g["L"], {})
    assert np.array_equal(np.array([k.pt for k in kps], np.float32), np.stack([g["cv_par_kx"], g["cv_par_ky"]], 1))
    assert np.array_equal(np.stack([k.descriptor for k in kps]), g["cv_par_desc"])
    L, _ = syn.stereo_pair(752, 480, 1003)
    keypoint.index_next = 7
    kps = det.detect_keypoints(L, {})
    x, y, s = oracle.grid_detect(L, (16, 16), 10)
    ref = oracle.corner_subpix(L, np.stack([x, y], 1).astype(np.float32))
    kept, desc = oracle.orb_compute(L, ref[:, 0].copy(), ref[:, 1].copy())
    assert [k.index for k in kps] == list(range(7, 7 + len(kept)))
    assert np.array_equal(np.array([k.pt for k in kps], np.float32), ref[kept])
    assert np.array_equal(np.array([k.response for k in kps], np.float32), s[kept].astype(np.float32))
    assert np.array_equal(np.stack([k.descriptor for k in kps]), desc)
    assert any(k.pt[0] != int(k.pt[0]) for k in kps)                  # points really are sub-pixel


def test_processor_mirror_golden_and_oracle(ctx, golden):
    """pre-processing on the device (processor.cpp:25-55): against the cv2 fixture and, at full frame size, the oracle"""
    from zenslam_b200.processing import processor
    g = golden("preproc")
    maps = [(g["map_x"], g["map_y"])]
    assert np.array_equal(processor(ctx).process_image(g["bgr"]), g["cv_gray"])
    assert np.array_equal(processor(ctx, clahe_enabled=True).process_image(g["bgr"]), g["cv_clahe"])
    assert np.array_equal(processor(ctx, maps=maps).process_image(g["bgr"]), g["cv_remap_gray"])
    assert np.array_equal(processor(ctx, clahe_enabled=True, maps=maps).process_image(g["bgr"]), g["cv_remap_clahe"])
    assert np.array_equal(processor(ctx, clahe_enabled=True, maps=maps).process_image(g["cv_gray"]), g["cv_remap_clahe"])   # gray input
    # full-size frame, width not a multiple of 4 in the BGR rows, against the oracle
    w, h = 750, 478
    rng = np.random.default_rng(5)
    tex = syn.crop(syn.base_texture(752, 480, 77), 752, 480, 0, 0)[:h, :w]
    bgr = np.stack([tex, np.roll(tex, 3, 1), 255 - tex], -1).astype(np.uint8)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    r2 = ((xx - w / 2) ** 2 + (yy - h / 2) ** 2) / (w * w / 4)
    mx = (w / 2 + (xx - w / 2) * (1 + 0.11 * r2)).astype(np.float32); my = (h / 2 + (yy - h / 2) * (1 + 0.11 * r2)).astype(np.float32)
    want = oracle.remap_linear(oracle.clahe(oracle.bgr2gray(bgr), 4.0), mx, my)
    got = processor(ctx, clahe_enabled=True, maps=[(mx, my)]).process_image(bgr)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("n,m", [(1100, 30000), (37, 5000), (300, 2049)])
def test_assign_landmark_indices_large_map(ctx, n, m):
    """new keypoints against a large landmark map (rectangular Hamming, train side split across blocks),
    cross-check + distance gate of keypoint_tracker.cpp:262-287, against the oracle"""
    from zenslam_b200 import keypoint
    from zenslam_b200.matching import assign_landmark_indices
    rng = np.random.default_rng(n + m)
    lm = rng.integers(0, 256, (m, 32), dtype=np.uint8)
    q = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    src = rng.choice(m, n // 2, replace=False)
    q[:n // 2] = lm[src]                                              # true landmarks, some with a few flipped bits
    flip = rng.integers(0, 32, n // 2)
    q[np.arange(n // 2), flip] ^= rng.integers(0, 4, n // 2).astype(np.uint8)
    lm[(src[:5] + 1) % m] = lm[src[:5]]                               # duplicated landmarks: ties resolved by index
    kps = [keypoint(pt=(0.0, 0.0), index=10_000_000 + i, descriptor=q[i]) for i in range(n)]
    kps[3].descriptor = None                                          # keypoints without descriptors are skipped
    landmark_indices = np.arange(m) * 7 + 3
    got_n = assign_landmark_indices(ctx, kps, lm, landmark_indices, 32.0)
    rows = [i for i in range(n) if i != 3]
    oq, ot, od = oracle.match_hamming_cross(q[rows], lm)
    want = {rows[a]: int(landmark_indices[b]) for a, b, d in zip(oq, ot, od) if d <= 32.0}
    assert got_n == len(want) and got_n >= n // 4
    for i, kp in enumerate(kps):
        assert kp.index == want.get(i, 10_000_000 + i), i


@pytest.mark.parametrize("w,h,cell,name", [(1280, 1024, (32, 32), "C4"), (3840, 2160, (32, 32), "C5")])
def test_frontend_full_size_configs(ctx, w, h, cell, name):
    """BASELINE configs 4 and 5 at their full frame sizes.  Detection, description and the stereo match are checked
    against the oracle outright (they are cheap on the CPU); the 8 KLT calls are checked (a) against the oracle on a
    sample of the points and (b) through size-independent properties: with integer ground-truth motion the forward
    track of every kept point lands within 0.05 px of the true displacement, forward-backward returns to the start,
    and a second run of the same batch reproduces every output bit for bit."""
    from zenslam_b200 import detection_options, slam_options, tracking_options
    from zenslam_b200.frontend import StereoFrontend
    B = 2
    opts = slam_options(matcher="KNN", matcher_ratio=0.8, detection=detection_options(cell_size=cell), tracking=tracking_options())
    seq, offs = syn.stereo_sequence(w, h, B, 4000 + w, subpixel=False)        # integer motion: exact ground truth
    fe = StereoFrontend(ctx, w, h, B, opts)
    L, R = np.ascontiguousarray(seq[:, 0]), np.ascontiguousarray(seq[:, 1])
    res = {k: v.copy() for k, v in fe.process(L, R).items()}
    fe.close()
    fe = StereoFrontend(ctx, w, h, B, opts)
    res2 = fe.process(L, R)
    for k in res:
        assert np.array_equal(res[k], res2[k]), k                                # determinism at full size
    fe.close()
    win, ml = (31, 31), 3
    for k in range(B):
        xyl, rl, dl = oracle_detect(L[k], cell, 10)
        xyr, rr, dr = oracle_detect(R[k], cell, 10)
        nl, nr = res["n_left"][k], res["n_right"][k]
        assert nl == len(xyl) and nr == len(xyr) and nl > 0.7 * (w // cell[0]) * (h // cell[1]) * 0.8
        assert np.array_equal(res["kp_left"][k, :nl], xyl) and np.array_equal(res["kp_right"][k, :nr], xyr)
        assert np.array_equal(res["desc_left"][k, :nl], dl) and np.array_equal(res["desc_right"][k, :nr], dr)
        oi, od = oracle.match_hamming_knn2(dl, dr)
        assert np.array_equal(res["match_idx"][k, :nl], oi) and np.array_equal(res["match_dist"][k, :nl], od.astype(np.float32))
        # stereo L->R: the right view is the left view shifted by the disparity (6 px): true flow = (-6, 0)
        keep = res["track_keep"][2, k, :nl].astype(bool)
        assert keep.mean() > 0.9
        flow = res["track_pts"][2, k, :nl][keep] - xyl[keep]
        assert np.abs(flow - np.array([-6.0, 0.0], np.float32)).max() < 0.05
        # oracle on a sample of the points (every 16th), forward + backward + gate
        PL, PR = oracle.Pyramid(L[k], win, ml), oracle.Pyramid(R[k], win, ml)
        sel = np.arange(0, nl, 16)
        o1, os_, _ = oracle.lk_track(PL, PR, xyl[sel], None, win, ml)
        ob, osb, _ = oracle.lk_track(PR, PL, o1, None, win, ml)
        assert np.array_equal(res["track_pts"][2, k, sel], o1)
        assert np.array_equal(res["track_keep"][2, k, sel].astype(bool), oracle.fb_check(xyl[sel], ob, os_, osb, 1.0))
    # temporal L: frame 0 -> 1 moves by the known integer offset
    d = offs[1] - offs[0]
    n0 = res["n_left"][0]
    keep = res["track_keep"][0, 1, :n0].astype(bool)
    assert res["track_n"][0, 1] == n0 and keep.mean() > 0.9
    flow = res["track_pts"][0, 1, :n0][keep] - res["kp_left"][0, :n0][keep]
    assert np.abs(flow + d.astype(np.float32)).max() < 0.05


def test_triangulator_mirror_tolerance_1e5_relative(ctx, golden):
    """epipolar gate + cv::triangulatePoints + reprojection/depth/parallax gates on the device, against the
    cv2 fixture and the oracle.  Floating point: the 3-D points must agree to float rounding (they are ratios of
    float-rounded SVD outputs), the keep decisions everywhere except within 1e-6 of a threshold.
    TOLERANCE (this is the one tolerance-based row of the path): 3-D points within 1e-5 relative of cv::triangulatePoints,
    >= 98 % of them the very same floats; gate inputs within rtol 1e-5; keep flags identical outside a 1e-6 margin."""
    from zenslam_b200 import keypoint
    from zenslam_b200.triangulation import triangulation_options, triangulator
    g = golden("triangulate")
    tri = triangulator(ctx, g["P0"], g["P1"], g["F"], g["t"])
    xyz, keep, diag = tri.triangulate_points(g["pts0"], g["pts1"], with_diag=True)
    X4 = g["cv_points4d"].astype(np.float64)
    ok = np.abs(X4[3]) > 1e-9
    want = np.where(ok, X4[:3] / np.where(ok, X4[3], 1.0), 0.0).T
    scale = np.maximum(np.abs(want).max(1, keepdims=True), 1e-6)
    assert (np.abs(xyz - want) / scale).max() < 1e-5
    assert np.mean(np.all(xyz == want, axis=1)) > 0.98                 # almost always the very same floats
    oxyz, okeep, odiag = oracle.triangulate_keypoints(g["P0"], g["P1"], g["F"], g["t"], g["pts0"], g["pts1"])
    n0 = np.linalg.norm(oxyz, axis=1)
    margin = np.min(np.stack([np.abs(np.abs(odiag[:, 0]) - 0.01), np.abs(n0 - 1.0), np.abs(n0 - 50.0), np.abs(odiag[:, 1] - 1.0),
                              np.abs(odiag[:, 2] - 1.0), np.abs(odiag[:, 3] - 0.25), np.abs(oxyz[:, 2])]), 0)
    safe = margin > 1e-6
    assert safe.mean() > 0.99 and np.array_equal(keep[safe], okeep[safe])
    assert np.allclose(diag[safe], odiag[safe], rtol=1e-5, atol=1e-7)
    # map overload: pairs by common index, ascending (triangulator.cpp:39-132)
    k0 = {i * 3: keypoint(pt=tuple(g["pts0"][i]), index=i * 3) for i in range(200)}
    k1 = {i * 3: keypoint(pt=tuple(g["pts1"][i]), index=i * 3) for i in range(0, 200, 2)}
    pts = tri.triangulate_keypoints(k0, k1)
    want_idx = [i * 3 for i in range(0, 200, 2) if keep[i]]
    assert [p.index for p in pts] == want_idx
    # no epipolar filter keeps a superset
    tri2 = triangulator(ctx, g["P0"], g["P1"], None, g["t"], triangulation_options(filter_epipolar=False))
    _, keep2 = tri2.triangulate_points(g["pts0"], g["pts1"])
    assert keep2.sum() >= keep.sum()
    assert tri.triangulate_points(np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32))[0].shape == (0, 3)


@pytest.mark.parametrize("clahe,use_maps,channels", [(True, True, 3), (False, True, 1), (True, False, 3), (False, False, 3)])
def test_frontend_with_preprocessing(ctx, clahe, use_maps, channels):
    """raw camera frames in (SURVEY 8 f1 hooked into the batched front-end): BGR -> gray -> CLAHE -> rectification on the
    device must give exactly the results of the plain front-end fed with frames pre-processed by the oracle"""
    from zenslam_b200 import slam_options
    from zenslam_b200.frontend import StereoFrontend
    w, h, B = 752, 480, 2
    seq, _ = syn.stereo_sequence(w, h, 2 * B, 611, subpixel=True)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    maps = []
    for cam, kk in enumerate((0.06, -0.04)):
        r2 = ((xx - w / 2) ** 2 + (yy - h / 2) ** 2) / (w * w / 4)
        maps.append(((w / 2 + (xx - w / 2) * (1 + kk * r2) + 0.3 * cam).astype(np.float32), (h / 2 + (yy - h / 2) * (1 + kk * r2)).astype(np.float32)))

    def raw(img):                                   # a colour frame whose gray value is not simply one channel
        if channels == 1:
            return img
        return np.stack([img, np.roll(img, 2, 1), (img.astype(np.int32) * 3 // 4 + 20).astype(np.uint8)], -1)

    def pre(img, cam):
        g = oracle.bgr2gray(raw(img)) if channels == 3 else img
        if clahe:
            g = oracle.clahe(g, 4.0)
        if use_maps:
            g = oracle.remap_linear(g, maps[cam][0], maps[cam][1])
        return g

    fe_ref = StereoFrontend(ctx, w, h, B, slam_options())
    fe = StereoFrontend(ctx, w, h, B, slam_options())
    fe.set_preprocess(channels, clahe, 4.0, maps if use_maps else None)
    for b in range(2):
        chunk = seq[b * B:(b + 1) * B]
        Lr = np.ascontiguousarray(np.stack([raw(f) for f in chunk[:, 0]])); Rr = np.ascontiguousarray(np.stack([raw(f) for f in chunk[:, 1]]))
        Lp = np.ascontiguousarray(np.stack([pre(f, 0) for f in chunk[:, 0]])); Rp = np.ascontiguousarray(np.stack([pre(f, 1) for f in chunk[:, 1]]))
        want = {k: v.copy() for k, v in fe_ref.process(Lp, Rp).items()}
        if b == 0:
            got = fe.process(Lr, Rr)
        else:
            fe.submit(Lr, Rr); got = fe.wait()
        assert want["n_left"].min() > 300
        for k in range(B):
            nl, nr = want["n_left"][k], want["n_right"][k]
            assert got["n_left"][k] == nl and got["n_right"][k] == nr
            for key in ("kp_left", "desc_left", "match_idx", "match_pass"):
                assert np.array_equal(got[key][k, :nl], want[key][k, :nl]), (b, k, key)
            assert np.array_equal(got["desc_right"][k, :nr], want["desc_right"][k, :nr])
            for kind in range(4):
                m = want["track_n"][kind, k]
                assert got["track_n"][kind, k] == m
                assert np.array_equal(got["track_pts"][kind, k, :m], want["track_pts"][kind, k, :m])
                assert np.array_equal(got["track_keep"][kind, k, :m], want["track_keep"][kind, k, :m])
    fe.close(); fe_ref.close()
