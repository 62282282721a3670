"""Live pinning of the C oracle against the real dependency: where Python cv2 is importable (this image ships 4.13.0), seeded
random inputs -- other than the committed fixtures -- go through cv2 and through the oracle and must agree bit for bit.
Skipped as a whole when cv2 is absent.  CPU only."""
import numpy as np
import pytest

import oracle
from zenslam_b200 import synthetic as syn

cv2 = pytest.importorskip("cv2")


def _img(rng, w, h, kind=0):
    img = syn.crop(syn.base_texture(w, h, int(rng.integers(1, 1 << 30))), w, h, 0, 0)
    if kind == 1:
        img = (img // 32 * 32).astype(np.uint8)
    elif kind == 2:
        img = img.copy(); img[h // 4:h // 2, w // 4:w // 2] = 90
    return np.ascontiguousarray(img)


@pytest.mark.parametrize("seed", range(4))
def test_fast_orb_live(seed):
    rng = np.random.default_rng(700 + seed)
    w, h, thr = int(rng.integers(100, 400)), int(rng.integers(100, 300)), int(rng.choice([5, 10, 25]))
    img = _img(rng, w, h, seed % 3)
    kps = cv2.FastFeatureDetector_create(thr, True).detect(img)
    x, y, s = oracle.fast_detect(img, thr)
    assert [(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in kps] == list(zip(x.tolist(), y.tolist(), s.tolist()))
    k2, d = cv2.ORB_create().compute(img, kps)
    kept, desc = oracle.orb_compute(img, x, y)
    assert len(k2) == len(kept) and (len(kept) == 0 or np.array_equal(d, desc))
    assert [k.pt for k in k2] == [(float(x[i]), float(y[i])) for i in kept]


@pytest.mark.parametrize("seed", range(4))
def test_lk_live(seed):
    rng = np.random.default_rng(720 + seed)
    w, h = int(rng.integers(150, 420)), int(rng.integers(120, 320))
    win, ml = [((31, 31), 3), ((21, 21), 2), ((15, 15), 4), ((25, 13), 1)][seed]
    base = syn.base_texture(w, h, int(rng.integers(1, 1 << 30)))
    A = syn.crop(base, w, h, 0, 0); B = syn.crop(base, w, h, float(rng.uniform(-6, 6)), float(rng.uniform(-6, 6)))
    n = 150
    pts = np.stack([rng.uniform(-10, w + 10, n), rng.uniform(-10, h + 10, n)], 1).astype(np.float32)
    crit = (cv2.TERM_CRITERIA_COUNT | cv2.TERM_CRITERIA_EPS, 99, 0.001)
    p1, st, err = cv2.calcOpticalFlowPyrLK(A, B, pts.reshape(-1, 1, 2), None, winSize=win, maxLevel=ml, criteria=crit,
                                           flags=cv2.OPTFLOW_LK_GET_MIN_EIGENVALS, minEigThreshold=1e-4)
    o1, os_, oe = oracle.lk_track(oracle.Pyramid(A, win, ml), oracle.Pyramid(B, win, ml), pts, None, win, ml)
    # cv2's SIMD path accumulates the mismatch vector in float lanes, the oracle (like the CUDA kernel) in exact integers:
    # positions agree to the north-star tolerance of 0.01 px (typically 1e-4), status flags exactly
    st = st.reshape(-1)
    assert np.array_equal(st, os_)
    ok = st > 0
    assert np.abs(p1.reshape(-1, 2) - o1)[ok].max() < 0.01
    assert np.allclose(err.reshape(-1), oe, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("seed", range(3))
def test_matchers_live(seed):
    rng = np.random.default_rng(740 + seed)
    nq, nt = int(rng.integers(2, 300)), int(rng.integers(2, 300))
    q = rng.integers(0, [256, 4, 2][seed], (nq, 32)).astype(np.uint8); t = rng.integers(0, [256, 4, 2][seed], (nt, 32)).astype(np.uint8)
    knn = cv2.BFMatcher(cv2.NORM_HAMMING, False).knnMatch(q, t, 2)
    oi, od = oracle.match_hamming_knn2(q, t)
    assert [[m.trainIdx for m in r] for r in knn] == oi.tolist() and [[m.distance for m in r] for r in knn] == od.astype(np.float32).tolist()
    cross = cv2.BFMatcher(cv2.NORM_HAMMING, True).match(q, t)
    oq, ot, odd = oracle.match_hamming_cross(q, t)
    assert [(m.queryIdx, m.trainIdx, m.distance) for m in cross] == list(zip(oq.tolist(), ot.tolist(), odd.astype(np.float32).tolist()))
    fq = rng.integers(0, 256, (nq, 128)).astype(np.float32); ft = rng.integers(0, 256, (nt, 128)).astype(np.float32)
    knn = cv2.BFMatcher(cv2.NORM_L2, False).knnMatch(fq, ft, 2)
    oi, od = oracle.match_l2_knn2(fq, ft)
    assert [[m.trainIdx for m in r] for r in knn] == oi.tolist() and [[m.distance for m in r] for r in knn] == od.tolist()


@pytest.mark.parametrize("seed", range(3))
def test_orb_multiscale_detector_live(seed):
    rng = np.random.default_rng(760 + seed)
    w, h, thr = int(rng.integers(200, 500)), int(rng.integers(160, 400)), int(rng.choice([5, 10, 20]))
    img = _img(rng, w, h, seed % 3)
    mask = None
    if seed:
        mask = np.full((h, w), 255, np.uint8)
        for _ in range(100):
            cv2.circle(mask, (int(rng.integers(0, w)), int(rng.integers(0, h))), 8, 0, -1)
    kps = cv2.ORB_create(500, 1.2, 8, 31, 0, 2, cv2.ORB_HARRIS_SCORE, 31, thr).detect(img, mask)
    k2, desc = cv2.ORB_create().compute(img, kps)
    order = [i for _, _, _, i in sorted((k.octave, k.pt[1], k.pt[0], i) for i, k in enumerate(k2))]
    o = oracle.orb_detect(img, mask, fast_threshold=thr)
    assert len(order) == len(o["x"])
    assert np.array_equal(np.array([k2[i].pt for i in order], np.float32).reshape(-1, 2), np.stack([o["x"], o["y"]], 1))
    assert np.array_equal(np.array([k2[i].angle for i in order], np.float32), o["angle"])
    assert np.array_equal(np.array([k2[i].response for i in order], np.float32), o["response"])
    assert np.array_equal(desc[order], o["desc"])


def test_pyramid_and_preprocessing_live():
    rng = np.random.default_rng(780)
    w, h = 333, 251
    img = _img(rng, w, h)
    _, pyr = cv2.buildOpticalFlowPyramid(img, (21, 21), 3, withDerivatives=True)
    P = oracle.Pyramid(img, (21, 21), 3)
    for l in range(P.levels):
        assert np.array_equal(np.ascontiguousarray(pyr[2 * l]), P.image(l))
        assert np.array_equal(np.ascontiguousarray(pyr[2 * l + 1]).reshape(P.image(l).shape + (2,)), P.deriv(l))
    bgr = np.stack([img, np.roll(img, 5, 1), 255 - img], -1).astype(np.uint8)
    gray = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
    assert np.array_equal(gray, oracle.bgr2gray(bgr))
    assert np.array_equal(cv2.createCLAHE(4.0).apply(gray), oracle.clahe(gray, 4.0))
    assert np.array_equal(cv2.resize(img, (277, 209), interpolation=cv2.INTER_LINEAR_EXACT), oracle.resize_linear_exact(img, 277, 209))


@pytest.mark.parametrize("radius", [0.0, 40.0])
def test_assign_landmark_indices_live(radius):
    """oracle/landmarks.py against keypoint_tracker::assign_landmark_indices (keypoint_tracker.cpp:199-291) restated directly
    over cv2.BFMatcher(NORM_HAMMING, crossCheck=True), including the reference's radius_search (first `count` landmarks)"""
    from oracle import landmarks as olm
    rng = np.random.default_rng(800)
    m, n = 700, 300
    cloud = olm.landmark_cloud()
    idx = rng.permutation(np.arange(10, 10 + 2 * m))[:m]
    idx.sort()
    desc = rng.integers(0, 256, (m, 32), dtype=np.uint8)
    xyz = rng.normal(0, 30, (m, 3))
    assert cloud.add(idx, xyz, desc) == m and cloud.add(idx[:9], xyz[:9], desc[:9]) == 0
    q = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    src = rng.choice(m, 200, replace=False)
    q[:200] = desc[src]
    q[:200, :3] ^= rng.integers(0, 256, (200, 3), dtype=np.uint8)            # some within 32 bits, some not
    center = np.array([3.0, -2.0, 1.0])
    got = olm.assign_landmark_indices(q, cloud, center, radius, 32.0)
    cand = m
    if radius > 0:
        d = xyz - center
        cand = int(np.count_nonzero(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2] < radius * radius))
        assert 0 < cand < m
    want = np.full(n, -1, np.int64)
    for mt in cv2.BFMatcher(cv2.NORM_HAMMING, True).match(q, desc[:cand]):
        if mt.distance <= 32.0:
            want[mt.queryIdx] = idx[mt.trainIdx]
    assert np.array_equal(got, want) and (got >= 0).sum() > 30


def test_match_temporal_live():
    """oracle/landmarks.py::match_temporal against utils::match_temporal (matching_utils.cpp:441-563) restated directly over
    cv2.BFMatcher(NORM_HAMMING, crossCheck=True) and cv2.findEssentialMat -- including the reference's e.at<float> reading of
    the CV_64F essential matrix"""
    from oracle import landmarks as olm
    rng = np.random.default_rng(811)
    n0, n1 = 260, 240
    K = np.array([[450.0, 0, 376.0], [0, 450.0, 240.0], [0, 0, 1.0]])
    # a rigid scene seen from two poses, so that findEssentialMat has inliers
    X = np.stack([rng.uniform(-4, 4, n0), rng.uniform(-3, 3, n0), rng.uniform(4, 12, n0)], 1)
    ang = 0.05
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    t = np.array([0.3, 0.02, 0.05])
    p0 = (X / X[:, 2:]) @ K.T
    X1 = X @ R.T + t
    p1_all = (X1 / X1[:, 2:]) @ K.T
    pts_0 = p0[:, :2].astype(np.float32)
    d0 = rng.integers(0, 256, (n0, 32), dtype=np.uint8)
    perm = rng.permutation(n0)[:n1]
    pts_1 = (p1_all[perm, :2] + rng.normal(0, 0.2, (n1, 2))).astype(np.float32)
    d1 = d0[perm].copy()
    d1[::3, 5] ^= 1                                                     # distance 1 .. and some beyond 5:
    d1[1::7, :2] ^= 255
    keys_0 = np.arange(100, 100 + n0); keys_1 = np.arange(5000, 5000 + n1)
    keys_1[:20] = keys_0[perm[:20]]                                     # shared indices drop out of both sides
    order = np.argsort(keys_1, kind="stable"); keys_1, pts_1, d1 = keys_1[order], pts_1[order], d1[order]
    s0, s1 = set(keys_0.tolist()), set(keys_1.tolist())
    u0 = [i for i in range(n0) if int(keys_0[i]) not in s1]; u1 = [i for i in range(n1) if int(keys_1[i]) not in s0]
    Ki = np.linalg.inv(K)
    # threshold 1.0: what the reference passes -- with the float reading of E the epipolar gate then rejects practically
    # everything (the reference's own behaviour); 1e30: every match reaches the mask / distance gates
    for thr in (1.0, 1e30):
        state = {}

        def fem(a, b):
            cv2.setRNGSeed(7)
            E, mask = cv2.findEssentialMat(a, b, K, cv2.RANSAC, 0.99, thr)
            state["E"], state["mask"] = E, mask
            return E, mask

        got = olm.match_temporal(keys_0, pts_0, d0, keys_1, pts_1, d1, K, thr, fem)
        # the restatement over cv2
        matches = cv2.BFMatcher(cv2.NORM_HAMMING, True).match(d0[u0], d1[u1])
        a = np.array([pts_0[u0[m.queryIdx]] for m in matches], np.float32); b = np.array([pts_1[u1[m.trainIdx]] for m in matches], np.float32)
        cv2.setRNGSeed(7)
        E, mask = cv2.findEssentialMat(a, b, K, cv2.RANSAC, 0.99, thr)
        Ef = np.frombuffer(np.ascontiguousarray(E, np.float64).tobytes(), np.float32).reshape(3, 6)[:, :3].astype(np.float64)
        want = []
        for k, (m, msk) in enumerate(zip(matches, mask.ravel())):
            if not msk:
                continue
            pl = np.array([a[k][0], a[k][1], 1.0], np.float64); pr = np.array([b[k][0], b[k][1], 1.0], np.float64)
            err = float((((pr @ Ki.T) @ Ef) @ Ki) @ pl)
            if err > thr or m.distance > 5:
                continue
            want.append((int(keys_0[u0[m.queryIdx]]), int(keys_1[u1[m.trainIdx]]), float(m.distance)))
        assert np.array_equal(state["E"], E) and np.array_equal(state["mask"], mask)
        assert got == want
        if thr > 1.0:
            assert len(got) > 20 and any(m.distance > 5 for m in matches)
    fem = lambda a, b: (np.eye(3), np.ones(len(a), np.uint8))
    assert olm.match_temporal(keys_0[:4], pts_0[:4], d0[:4], keys_1, pts_1, d1, K, 1.0, fem) == []
