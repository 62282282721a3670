"""GPU: landmark association -- keypoint_tracker::assign_landmark_indices inside the stateful device flow
(keypoint_tracker.cpp:55,71,199-291) and utils::match_keypoints3d (matching_utils.cpp:132-343) -- against the CPU
restatement oracle/landmarks.py (whose Hamming matcher is pinned to cv2 elsewhere)."""
import dataclasses

import numpy as np
import pytest

import oracle
from oracle import landmarks as olm
from zenslam_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from zenslam_b200.runtime import Context
    c = Context()
    yield c
    c.close()


@pytest.mark.parametrize("radius", [50.0, 0.0])
def test_device_tracker_with_landmark_map(ctx, radius):
    """zs_tracker with a non-empty landmark store against the keypoint_tracker.track mirror whose assign_landmark_indices is
    the oracle's: identical index sets, positions, responses and descriptors in both cameras on every frame.  The landmark
    map grows every frame like system.points3d (slam_thread.cpp:210), holds descriptors of keypoints the NEXT frame will
    detect (so detections do take landmark indices), indices the maps already hold (so map.add skips them), landmarks
    outside the radius and re-submitted indices (so `+=` skips them); the camera centre moves."""
    from zenslam_b200 import detection_options, keypoint, slam_options, tracking_options
    from zenslam_b200.detection import keypoint_detector_grid
    from zenslam_b200.keypoint_tracker import device_keypoint_tracker, keypoint_tracker, stereo_frame
    from zenslam_b200.tracking import create_cuda_pyr_lk
    w, h, frames = 376, 240, 7
    seq, _ = syn.stereo_sequence(w, h, frames, 3100, subpixel=True)
    opts = slam_options(matcher="KNN", detection=detection_options(),
                        tracking=tracking_options(filter_epipolar=False, landmark_match_radius=radius, landmark_match_distance=32.0))
    rng = np.random.default_rng(31)
    det = keypoint_detector_grid(opts.detection, ctx)

    def new_landmarks(t, maps):
        """landmarks to add before frame t is tracked: descriptors of what a full-grid detection finds in frame t (left and
        right), keyed partly by fresh landmark indices and partly by indices the current maps hold"""
        saved = keypoint.index_next
        cand = det.detect_keypoints(seq[t, 0], {})[::2] + det.detect_keypoints(seq[t, 1], {})[1::2]
        keypoint.index_next = saved
        held = sorted(set(maps[0]) | set(maps[1]))
        idx, desc = [], []
        for j, k in enumerate(cand):
            if held and j % 5 == 0:
                idx.append(int(held[(7 * j) % len(held)]))
            else:
                idx.append(500000 + 1000 * t + j)
            desc.append(k.descriptor)
        xyz = rng.normal(0.0, 28.0, (len(idx), 3))                   # some beyond the 50 m radius
        return canon(np.array(idx, np.int64), xyz, np.stack(desc))

    def canon(idx, xyz, desc):
        """what `+=` sees: a std::map holds every index once (first one wins here) and is iterated in key order"""
        _, first = np.unique(idx, return_index=True)
        return idx[first], xyz[first], desc[first]

    # ---- the mirror: CUDA detector / pyr_lk seams + the oracle's assign_landmark_indices
    cloud = olm.landmark_cloud()
    state = {"center": np.zeros(3)}
    assigned_total = [0]

    def assign(detected):
        if not detected:
            return
        got = olm.assign_landmark_indices(np.stack([k.descriptor for k in detected]), cloud, state["center"], radius, 32.0)
        for k, li in zip(detected, got):
            if li >= 0:
                k.index = int(li); assigned_total[0] += 1

    keypoint.index_next = 0
    host = keypoint_tracker(opts, ctx, create_cuda_pyr_lk(ctx), assign_landmarks=assign)
    prev = stereo_frame((seq[0, 0], seq[0, 1]))
    want, adds = [], []
    maps = ({}, {})
    for t in range(frames):
        if t >= 2:
            lm = new_landmarks(t, maps)
            if t == 4:                                                # re-submit a few known indices with other data: += skips them
                lm = canon(np.concatenate([cloud.index[:5], lm[0]]), np.concatenate([cloud.xyz[:5] + 1.0, lm[1]]),
                           np.concatenate([cloud.desc[:5] ^ 255, lm[2]]))
            adds.append((t, lm, cloud.add(*lm)))
        state["center"] = np.array([0.8 * t, -0.5 * t, 0.3 * t])
        cur = stereo_frame((seq[t, 0], seq[t, 1]))
        k0, k1 = host.track(prev, cur)
        want.append((k0, k1, keypoint.index_next))
        prev = dataclasses.replace(cur, keypoints=(k0, k1))
        maps = (k0, k1)
    assert assigned_total[0] >= 10, assigned_total                    # the association really fires
    if radius > 0:
        assert cloud.radius_count(state["center"], radius) < len(cloud)   # ... and the radius really cuts

    # ---- the device tracker
    keypoint.index_next = 0
    dev_trk = device_keypoint_tracker(opts, ctx, w, h, landmark_capacity=4096)
    ai = 0
    for t in range(frames):
        if ai < len(adds) and adds[ai][0] == t:
            assert dev_trk.add_landmarks(*adds[ai][1]) == adds[ai][2]
            ai += 1
        dev_trk.set_camera_center([0.8 * t, -0.5 * t, 0.3 * t])
        g0, g1 = dev_trk.track(seq[t, 0], seq[t, 1])
        for cam, (g, r) in enumerate(((g0, want[t][0]), (g1, want[t][1]))):
            assert list(g) == sorted(r), (t, cam, len(g), len(r), sorted(set(g) ^ set(r))[:10])
            for i in g:
                assert g[i].pt == r[i].pt and g[i].response == r[i].response and np.array_equal(g[i].descriptor, r[i].descriptor), (t, cam, i)
        assert keypoint.index_next == want[t][2]
    assert dev_trk.landmarks_size() == len(cloud)
    assert any(i >= 500000 for i in g0) and any(i >= 500000 for i in g1)      # landmark indices live in the maps
    dev_trk.close()


def test_device_tracker_empty_landmark_store_equals_plain_tracker(ctx):
    """points3d.empty() -> assign_landmark_indices returns at once (keypoint_tracker.cpp:208): a tracker with an empty
    landmark store must produce exactly what the tracker without one produces"""
    from zenslam_b200 import detection_options, keypoint, slam_options, tracking_options
    from zenslam_b200.keypoint_tracker import device_keypoint_tracker
    w, h, frames = 376, 240, 4
    seq, _ = syn.stereo_sequence(w, h, frames, 3200, subpixel=True)
    opts = slam_options(detection=detection_options(), tracking=tracking_options(filter_epipolar=False))
    out = []
    for lm_cap in (0, 1024):
        keypoint.index_next = 0
        trk = device_keypoint_tracker(opts, ctx, w, h, landmark_capacity=lm_cap)
        out.append([trk.track(seq[t, 0], seq[t, 1]) for t in range(frames)])
        trk.close()
    for t in range(frames):
        for cam in range(2):
            a, b = out[0][t][cam], out[1][t][cam]
            assert list(a) == list(b)
            assert all(a[i].pt == b[i].pt and np.array_equal(a[i].descriptor, b[i].descriptor) for i in a)


def _scene(rng, m, n_kp):
    """a cloud in front of / around a camera, keypoints that are noisy projections of some landmarks + strays"""
    from zenslam_b200 import keypoint
    ang = 0.3
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    t = np.array([2.0, -1.0, 0.5])
    K = np.array([[420.0, 0, 376.0], [0, 420.0, 240.0], [0, 0, 1.0]])
    P = np.hstack([K, np.zeros((3, 1))])
    cam = np.stack([rng.uniform(-12, 12, m), rng.uniform(-8, 8, m), rng.uniform(-5, 40, m)], 1)
    xyz = cam @ R.T + t                                       # world = R cam + t
    desc = rng.integers(0, 256, (m, 32), dtype=np.uint8)
    index = rng.permutation(np.arange(1000, 1000 + 3 * m))[:m].astype(np.int64)
    kps = {}
    pick = rng.choice(m, n_kp, replace=False)
    for j, i in enumerate(pick):
        z = max(cam[i, 2], 0.5)
        uv = (K @ (cam[i] / z))[:2] + rng.normal(0, [0.4, 3.0][j % 2], 2)      # half of them beyond a 2 px gate
        d = desc[i].copy()
        d[rng.integers(0, 32, 3)] ^= np.uint8(1 << int(rng.integers(0, 8)))    # a few flipped bits
        # a fifth of the keypoints carry a landmark's index: values_unmatched drops them
        ki = int(index[(i + 1) % m]) if j % 5 == 0 else 900000 + j
        kps[ki] = keypoint(pt=(float(np.float32(uv[0])), float(np.float32(uv[1]))), index=ki, descriptor=d)
    return index, xyz, desc, kps, R, t, P


@pytest.mark.parametrize("frustum", [False, True])
@pytest.mark.parametrize("m,n_kp,radius", [(3000, 900, 30.0), (200, 150, 1000.0), (50, 40, 5.0)])
def test_match_keypoints3d_vs_oracle(ctx, m, n_kp, radius, frustum):
    from zenslam_b200.matching import match_keypoints3d
    rng = np.random.default_rng(m + int(frustum))
    index, xyz, desc, kps, R, t, P = _scene(rng, m, n_kp)
    cloud = olm.landmark_cloud()
    order = np.argsort(index, kind="stable")
    cloud.add(index[order], xyz[order], desc[order])
    size, margin = ((752, 480), 50.0) if frustum else (None, None)
    keys = sorted(kps)
    ol, ok_, oe = olm.match_keypoints3d(cloud, keys, np.array([kps[k].pt for k in keys], np.float32),
                                        np.stack([kps[k].descriptor for k in keys]), R, t, P, radius, 2.0, size, margin)
    got = match_keypoints3d(ctx, cloud.index, cloud.xyz, cloud.desc, kps, R, t, P, radius, 2.0, size, margin if frustum else 50.0)
    assert [g.queryIdx for g in got] == ol.tolist() and [g.trainIdx for g in got] == ok_.tolist()
    # the reference inverts the pose numerically (cv::Affine3d::inv): the error is a float comparison, tolerance 1e-4 px
    assert np.allclose([g.distance for g in got], oe, rtol=0, atol=1e-4)
    if m >= 200:
        assert len(got) > 10
    assert match_keypoints3d(ctx, cloud.index, cloud.xyz, cloud.desc, {}, R, t, P, radius, 2.0) == []
    assert match_keypoints3d(ctx, [], np.zeros((0, 3)), np.zeros((0, 32), np.uint8), kps, R, t, P, radius, 2.0) == []


def test_match_temporal_vs_oracle(ctx):
    """utils::match_temporal (matching_utils.cpp:441-563): unmatched selection, the five-keypoint floor, the cross-checked
    Hamming match on the device and the mask / epipolar / distance <= 5 gate, with findEssentialMat supplied by the caller."""
    from zenslam_b200 import keypoint
    from zenslam_b200.matching import match_temporal
    rng = np.random.default_rng(77)
    n0, n1 = 400, 380
    d0 = rng.integers(0, 256, (n0, 32), dtype=np.uint8)
    d1 = rng.integers(0, 256, (n1, 32), dtype=np.uint8)
    d1[:250] = d0[50:300]
    flips = rng.integers(0, 9, 250)                                     # 0..8 flipped bits: some matches beyond distance 5
    for i, f in enumerate(flips):
        for b in rng.choice(256, f, replace=False):
            d1[i, b // 8] ^= np.uint8(1 << (b % 8))
    keys_0 = np.sort(rng.choice(5000, n0, replace=False)); keys_1 = np.sort(rng.choice(np.arange(5000, 9000), n1, replace=False))
    keys_1[:40] = keys_0[300:340]                                       # shared indices: dropped from both sides
    keys_1 = np.sort(keys_1)
    p0 = rng.uniform(0, 700, (n0, 2)).astype(np.float32); p1 = rng.uniform(0, 700, (n1, 2)).astype(np.float32)
    K = np.array([[450.0, 0, 376.0], [0, 450.0, 240.0], [0, 0, 1.0]])
    E = np.array([[0.0, -0.2, 0.05], [0.2, 0.0, -1.0], [-0.05, 1.0, 0.0]])
    seen = []

    def fem(points_0, points_1):
        seen.append((points_0.copy(), points_1.copy()))
        mask = (np.arange(len(points_0)) % 4 != 1).astype(np.uint8)     # what RANSAC would hand back: some outliers
        return E, mask

    m0 = {int(k): keypoint(pt=(float(p0[i, 0]), float(p0[i, 1])), index=int(k), descriptor=d0[i]) for i, k in enumerate(keys_0)}
    m1 = {int(k): keypoint(pt=(float(p1[i, 0]), float(p1[i, 1])), index=int(k), descriptor=d1[i]) for i, k in enumerate(keys_1)}
    for thr in (1e30, 0.0):
        got = match_temporal(ctx, m0, m1, K, thr, fem)
        want = olm.match_temporal(keys_0, p0, d0, keys_1, p1, d1, K, thr, fem)
        assert [(g.queryIdx, g.trainIdx, g.distance) for g in got] == want
        assert np.array_equal(seen[-1][0], seen[-2][0]) and np.array_equal(seen[-1][1], seen[-2][1])
    assert len(match_temporal(ctx, m0, m1, K, 1e30, fem)) > 20
    few = {k: m0[k] for k in list(m0)[:4]}
    assert match_temporal(ctx, few, m1, K, 1e30, fem) == [] and match_temporal(ctx, m1, few, K, 1e30, fem) == []


def test_match_keylines_and_keyline_landmarks_vs_cv2_restatement(ctx):
    """utils::match_keylines (matching_utils.cpp:345-439) and the keyline landmark association (keyline_tracker.cpp:135-163):
    the device Hamming matcher + the reference's gates against a restatement over the oracle matcher and, where cv2 imports,
    cv2.computeCorrespondEpilines itself"""
    from zenslam_b200 import keyline
    from zenslam_b200.matching import _epilines, assign_keyline_landmark_indices, match_keylines
    rng = np.random.default_rng(91)
    n0, n1 = 180, 170
    F = np.array([[0.0, 0.0, 0.0], [0.0, 0.0, -0.013], [0.0, 0.013, 0.0]])                        # rectified pair: horizontal epipolar lines
    d0 = rng.integers(0, 256, (n0, 32), dtype=np.uint8)
    perm = rng.permutation(n0)[:n1]
    d1 = d0[perm].copy(); d1[::4, 3] ^= 9
    def mk(i, d, s, e):
        return keyline(startPointX=float(s[0]), startPointY=float(s[1]), endPointX=float(e[0]), endPointY=float(e[1]),
                       pt=(float(np.float32(0.5) * np.float32(s[0] + e[0])), float(np.float32(0.5) * np.float32(s[1] + e[1]))), index=i, descriptor=d)
    s0 = rng.uniform(50, 700, (n0, 2)).astype(np.float32); e0 = s0 + rng.uniform(-40, 40, (n0, 2)).astype(np.float32)
    # image-1 lines: the same lines shifted along x (epipolar-consistent for this F up to noise), some pushed off their epiline
    shift = np.stack([rng.uniform(-30, -5, n1), rng.normal(0, 0.3, n1)], 1).astype(np.float32)
    shift[::5, 1] += 6.0
    s1 = s0[perm] + shift; e1 = e0[perm] + shift
    m0 = {10 + i: mk(10 + i, d0[i], s0[i], e0[i]) for i in range(n0)}
    m1 = {900 + i: mk(900 + i, d1[i], s1[i], e1[i]) for i in range(n1)}
    got = match_keylines(ctx, m0, m1, F, 2.0)
    oq, ot, od = oracle.match_hamming_cross(d0, d1)
    try:
        import cv2
    except Exception:
        cv2 = None
    want = []
    for a, b, d in zip(oq, ot, od):
        p0 = np.array([s0[a], e0[a], m0[10 + a].pt], np.float32); p1 = np.array([s1[b], e1[b], m1[900 + b].pt], np.float32)
        if cv2 is not None:
            l1 = cv2.computeCorrespondEpilines(p0.reshape(-1, 1, 2), 1, F).reshape(-1, 3)
            l0 = cv2.computeCorrespondEpilines(p1.reshape(-1, 1, 2), 2, F).reshape(-1, 3)
            assert np.allclose(l1, _epilines(p0, 1, F), rtol=0, atol=1e-5) and np.allclose(l0, _epilines(p1, 2, F), rtol=0, atol=1e-5)
        else:
            l1, l0 = _epilines(p0, 1, F), _epilines(p1, 2, F)
        err0 = np.abs(l0[:, 0] * p0[:, 0] + l0[:, 1] * p0[:, 1] + l0[:, 2]) / np.sqrt(l0[:, 0] ** 2 + l0[:, 1] ** 2)
        err1 = np.abs(l1[:, 0] * p1[:, 0] + l1[:, 1] * p1[:, 1] + l1[:, 2]) / np.sqrt(l1[:, 0] ** 2 + l1[:, 1] ** 2)
        worst = float(max(err0.max(), err1.max()))
        assert abs(worst - 2.0) > 1e-3, "a pair sits on the threshold: change the seed"     # float order differences stay below this
        if worst <= 2.0:
            want.append((10 + int(a), 900 + int(b), float(d)))
    assert [(g.queryIdx, g.trainIdx, g.distance) for g in got] == want
    assert 20 < len(got) < len(oq)                                          # the gate really cut something
    # keyline landmarks: 1-NN without cross check, distance gate
    lm_desc = rng.integers(0, 256, (400, 32), dtype=np.uint8); lm_desc[:60] = d0[:60]; lm_desc[:60, 0] ^= 1
    lm_idx = np.arange(7000, 7400)
    kls = [mk(i, d0[i], s0[i], e0[i]) for i in range(n0)]
    n = assign_keyline_landmark_indices(ctx, kls, lm_desc, lm_idx, 32.0)
    oi, odist = oracle.match_hamming_knn2(d0, lm_desc)
    want_idx = [int(lm_idx[oi[i, 0]]) if odist[i, 0] <= 32 else i for i in range(n0)]
    assert [k.index for k in kls] == want_idx and n == sum(1 for i in range(n0) if odist[i, 0] <= 32) and n >= 60
