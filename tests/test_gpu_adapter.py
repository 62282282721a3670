"""GPU: the C++ adapter in zenslam_cuda/ EXECUTES.  tests/adapter/adapter_harness.cpp is the adapter compiled against the
reference's own headers (pyr_lk.h, keypoint_detector.h, detection_options.h, tracking_options.h, keypoint.h, map.h) plus the
functional OpenCV stand-in of tests/stubs/, linked to libzenslam_cuda.so (tests/adapter/build_harness.py; built in the
container that holds /root/reference, it travels to the GPU box like the .so).  It calls
    zenslam::cuda::create_cuda_pyr_lk()->calc_optical_flow_pyr_lk   through zenslam::pyr_lk            (pyr_lk.h:15-26)
    zenslam::cuda::keypoint_detector_cuda::detect_keypoints         through zenslam::keypoint_detector (keypoint_detector.h:13)
    zenslam::cuda::bf_matcher                                       through cv::DescriptorMatcher::knnMatch / match, the
                                                                    two-image forms zenslam::matcher uses (matcher.cpp:65,79)
    zenslam::cuda::stereo_tracker::track                            (keypoint_tracker.cpp:41-105)
and this test compares what it dumps with the oracle (LK, grid detection, descriptors, matches: bit-exact) and with the
python mirror of the composite flows (PARALLEL_GRID, SIMPLE with its disc mask, the stateful tracker)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle
from zenslam_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "adapter"))
import build_harness  # noqa: E402

DT = {"u1": np.uint8, "i4": np.int32, "f4": np.float32, "f8": np.float64}
W, H, FRAMES, CELL, THR, WIN, LEVEL = 376, 240, 3, 16, 10, 31, 3


def write_arrays(path, arrays):
    with open(path, "wb") as f:
        for name, a in arrays.items():
            a = np.ascontiguousarray(a)
            code = {np.dtype(np.uint8): "u1", np.dtype(np.int32): "i4", np.dtype(np.float32): "f4", np.dtype(np.float64): "f8"}[a.dtype]
            f.write(("%s %s %d %s\n" % (name, code, a.ndim, " ".join(str(d) for d in a.shape))).encode())
            f.write(a.tobytes())
            f.write(b"\n")


def read_arrays(path):
    out = {}
    with open(path, "rb") as f:
        while True:
            line = f.readline()
            if not line:
                break
            if not line.strip():
                continue
            parts = line.decode().split()
            name, code, nd = parts[0], parts[1], int(parts[2])
            shape = tuple(int(p) for p in parts[3:3 + nd])
            dt = np.dtype(DT[code])
            out[name] = np.frombuffer(f.read(int(np.prod(shape)) * dt.itemsize), dt).reshape(shape)
            f.read(1)
    return out


@pytest.fixture(scope="module")
def harness():
    exe = build_harness.EXE
    if build_harness.available():
        exe = build_harness.build()
    if not os.path.exists(exe):
        pytest.fail("tests/adapter/_build/adapter_harness is missing: run __graft_entry__.build() where /root/reference exists")
    return exe


@pytest.fixture(scope="module")
def inputs():
    seq, _ = syn.stereo_sequence(W, H, FRAMES, 5100, subpixel=True)
    rng = np.random.default_rng(51)
    x, y, _ = oracle.grid_detect(seq[0, 0], (CELL, CELL), THR)
    kept, _ = oracle.orb_compute(seq[0, 0], x, y)
    pts = np.stack([x[kept], y[kept]], 1).astype(np.float32)
    pts = np.concatenate([pts + rng.uniform(-0.5, 0.5, pts.shape).astype(np.float32),
                          np.array([[2.0, 3.0], [W - 1.5, H - 2.0], [-40.0, 5.0]], np.float32)])     # borders, one outside
    init = pts + rng.normal(0, 1.5, pts.shape).astype(np.float32)
    # existing keypoints: every third grid corner of the left frame (their cells are occupied / their discs masked)
    ex = np.stack([np.arange(len(kept))[::3] + 7.0, x[kept][::3] + 0.25, y[kept][::3] + 0.5], 1).astype(np.float32)
    # ... and two that were tracked just outside the frame: C++ truncation puts them into column / row 0 (keypoint_detector_grid.cpp:51-52)
    ex = np.concatenate([ex, np.array([[9001.0, -5.5, 40.0], [9002.0, 100.0, -3.2]], np.float32)])
    q = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (260, 32), dtype=np.uint8)
    t[:120] = q[40:160]
    t[:120, 3] ^= 5
    t[200] = t[10]                                                                                  # an exact tie
    qf = rng.integers(0, 120, (150, 128)).astype(np.float32)
    tf = rng.integers(0, 120, (170, 128)).astype(np.float32)
    tf[:60] = qf[30:90]
    # landmarks for the stateful flow: descriptors of corners the tracker will detect in frames 1 and 2 (a few bits flipped),
    # positions some of which lie beyond the 50 m radius
    lm_desc = []
    for f in (1, 2):
        xf, yf, _ = oracle.grid_detect(seq[f, 0], (CELL, CELL), THR)
        _, df = oracle.orb_compute(seq[f, 0], xf, yf)
        lm_desc.append(df[::2])
    lm_desc = np.concatenate(lm_desc).copy()
    lm_desc[:, 7] ^= 3
    lm_index = (700000 + 3 * np.arange(len(lm_desc))).astype(np.int32)
    lm_xyz = rng.normal(0.0, 30.0, (len(lm_desc), 3))
    # processor::process: a BGR frame + rectification maps that leave the frame here and there
    g = seq[0, 0]
    bgr = np.stack([g, np.roll(g, 3, 1), 255 - g // 2], -1).astype(np.uint8)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    map_x = (xx + 3.5 * np.sin(yy / 37.0) - 4.0).astype(np.float32); map_y = (yy * 1.01 + 2.0 * np.cos(xx / 53.0) - 1.5).astype(np.float32)
    # the triangulator's core: the committed cv2 fixture's geometry
    tg = np.load(os.path.join(ROOT, "tests", "golden", "triangulate.npz"))
    return {"dims": np.array([W, H, FRAMES, CELL, THR, WIN, LEVEL], np.int32), "frames": seq, "lk_points": pts, "lk_initial": init,
            "existing": ex, "match_q": q, "match_t": t, "match_qf": qf, "match_tf": tf,
            "lm_index": lm_index, "lm_xyz": lm_xyz, "lm_desc": lm_desc,
            "pp_bgr": bgr, "pp_map_x": map_x, "pp_map_y": map_y,
            "tri_P": np.stack([tg["P0"], tg["P1"]]).astype(np.float64), "tri_F": tg["F"].astype(np.float64), "tri_t": tg["t"].astype(np.float64).reshape(3),
            "tri_pts0": tg["pts0"].astype(np.float32), "tri_pts1": tg["pts1"].astype(np.float32)}


@pytest.fixture(scope="module")
def results(harness, inputs, tmp_path_factory):
    d = tmp_path_factory.mktemp("adapter")
    write_arrays(d / "in.bin", inputs)
    r = subprocess.run([harness, str(d / "in.bin"), str(d / "out.bin")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
    return read_arrays(d / "out.bin")


@pytest.fixture(scope="module")
def ctx():
    from zenslam_b200.runtime import Context
    c = Context()
    yield c
    c.close()


def test_pyr_lk_seam_equals_oracle(results, inputs):
    seq, pts = inputs["frames"], inputs["lk_points"]
    P0, P1 = oracle.Pyramid(seq[0, 0], (WIN, WIN), LEVEL), oracle.Pyramid(seq[1, 0], (WIN, WIN), LEVEL)
    nxt, st, err = oracle.lk_track(P0, P1, pts, None, (WIN, WIN), LEVEL)
    assert st.sum() > 100 and not st.all()
    assert np.array_equal(results["lk.status"].ravel(), st)
    assert np.array_equal(results["lk.next"], nxt) and np.array_equal(results["lk.err"].ravel(), err)
    # a continuous image against an ROI of a padded buffer gives the same answer (the adapter equalises the pitches)
    assert np.array_equal(results["lk_mixed.next"], nxt)
    flags = oracle.LK_GET_MIN_EIGENVALS | oracle.LK_USE_INITIAL_FLOW
    ni, si, _ = oracle.lk_track(P0, P1, pts, inputs["lk_initial"], (WIN, WIN), LEVEL, flags=flags)
    assert np.array_equal(results["lk_init.next"], ni) and np.array_equal(results["lk_init.status"].ravel(), si)
    # forward + backward + gate (keypoint_tracker.cpp:142-186)
    back, sb, _ = oracle.lk_track(P1, P0, nxt, None, (WIN, WIN), LEVEL)
    keep = oracle.fb_check(pts, back, st, sb, 1.0)
    assert np.array_equal(results["lk_fb.keep"].ravel().astype(bool), keep) and keep.sum() > 100
    assert np.array_equal(results["lk_fb.next"][keep], nxt[keep])
    assert results["lk_empty.sizes"].ravel().tolist() == [0, 0, 0]


def _occupancy(ex):
    gw, gh = W // CELL, H // CELL
    occ = np.zeros((gh, gw), np.uint8)
    for _, x, y in ex:
        gx, gy = int(int(x) / CELL), int(int(y) / CELL)                  # C++: both the cast and the division truncate toward zero
        if 0 <= gx < gw and 0 <= gy < gh:
            occ[gy, gx] = 1
    return occ


def test_grid_detector_seam_equals_oracle(results, inputs):
    img = inputs["frames"][0, 0]
    for name, occ in (("grid", None), ("grid_occ", _occupancy(inputs["existing"]))):
        x, y, s = oracle.grid_detect(img, (CELL, CELL), THR, occ)
        kept, desc = oracle.orb_compute(img, x, y)
        f, i, d = results["det.%s.f" % name], results["det.%s.i" % name], results["det.%s.desc" % name]
        assert len(f) == len(kept) > 50
        assert np.array_equal(f[:, 0], x[kept].astype(np.float32)) and np.array_equal(f[:, 1], y[kept].astype(np.float32))
        assert np.array_equal(f[:, 2], s[kept].astype(np.float32))                         # response = FAST score
        assert np.all(f[:, 3] == 7.0) and np.all(f[:, 4] == -1.0)                          # what cv::FAST emits
        assert np.array_equal(i[:, 0], 1000 + np.arange(len(kept))) and np.all(i[:, 1] == 0) and np.all(i[:, 2] == -1)
        assert np.array_equal(d, desc)
        assert int(results["det.%s.index_next" % name].ravel()[0]) == 1000 + len(kept)
    assert len(results["det.grid_occ.f"]) < len(results["det.grid.f"])                     # occupied cells were skipped
    assert int(results["det.rejects_sift"].ravel()[0]) == 1


def test_parallel_and_simple_detectors_equal_python_mirror(results, inputs, ctx):
    from zenslam_b200 import detection_options, keypoint
    from zenslam_b200.detection import keypoint_detector_parallel, keypoint_detector_simple
    img = inputs["frames"][0, 0]
    existing = {int(i): keypoint(pt=(float(x), float(y)), index=int(i)) for i, x, y in inputs["existing"]}
    cases = (("parallel", keypoint_detector_parallel, dict(algorithm="PARALLEL_GRID")),
             ("simple", keypoint_detector_simple, dict(algorithm="SIMPLE")),
             ("simple_orb", keypoint_detector_simple, dict(algorithm="SIMPLE", feature_detector="ORB")))
    for name, cls, kw in cases:
        opts = detection_options(cell_size=(CELL, CELL), fast_threshold=THR, **kw)
        keypoint.index_next = 1000
        want = cls(opts, ctx).detect_keypoints(img, existing)
        f, i, d = results["det.%s.f" % name], results["det.%s.i" % name], results["det.%s.desc" % name]
        assert len(f) == len(want) > 30, (name, len(f), len(want))
        assert np.array_equal(f[:, :2], np.array([k.pt for k in want], np.float32)), name
        assert np.array_equal(f[:, 2], np.array([k.response for k in want], np.float32)), name
        assert np.array_equal(f[:, 3], np.array([k.size for k in want], np.float32)), name
        assert np.array_equal(f[:, 4], np.array([k.angle for k in want], np.float32)), name
        assert np.array_equal(i[:, 0], [k.index for k in want]) and np.array_equal(i[:, 1], [k.octave for k in want]), name
        assert np.array_equal(d, np.stack([k.descriptor for k in want])), name
    # the disc mask really removed corners next to the existing keypoints
    keypoint.index_next = 0
    free = keypoint_detector_simple(detection_options(cell_size=(CELL, CELL), fast_threshold=THR, algorithm="SIMPLE"), ctx).detect_keypoints(img, {})
    assert len(results["det.simple.f"]) < len(free)


def _rows(results, prefix):
    ln, idx, dist = results[prefix + ".len"].ravel(), results[prefix + ".idx"].reshape(-1, 3), results[prefix + ".dist"].ravel()
    out, o = [], 0
    for n in ln:
        out.append([(int(idx[o + j, 0]), int(idx[o + j, 1]), int(idx[o + j, 2]), float(dist[o + j])) for j in range(n)])
        o += n
    return out


def test_descriptor_matcher_forwarding_equals_oracle(results, inputs):
    """cv::DescriptorMatcher::knnMatch(query, train, out, 2) and ::match(query, train, out) reach knnMatchImpl with
    std::vector<cv::Mat>(1, cv::Mat()) as masks: the call must go through (it used to assert) and return
    cv::BFMatcher's answer."""
    q, t = inputs["match_q"], inputs["match_t"]
    idx, dist = oracle.match_hamming_knn2(q, t)
    rows = _rows(results, "match.knn")
    assert len(rows) == len(q)
    for i, row in enumerate(rows):
        assert [(m[0], m[1], m[2]) for m in row] == [(i, int(idx[i, 0]), 0), (i, int(idx[i, 1]), 0)]
        assert [m[3] for m in row] == [float(dist[i, 0]), float(dist[i, 1])]
    oq, ot, od = oracle.match_hamming_cross(q, t)
    (cross,) = _rows(results, "match.cross")
    assert [(m[0], m[1]) for m in cross] == list(zip(oq.tolist(), ot.tolist())) and [m[3] for m in cross] == od.astype(np.float32).tolist()
    assert len(cross) > 100
    one = _rows(results, "match.knn_one")
    assert len(one) == len(q) and all(len(r) == 1 and r[0][1] == 0 for r in one)
    idx, dist = oracle.match_l2_knn2(inputs["match_qf"], inputs["match_tf"])
    rows = _rows(results, "match.l2")
    assert [[m[1] for m in r] for r in rows] == idx.tolist()
    assert np.array_equal(np.array([[m[3] for m in r] for r in rows], np.float32), dist)
    assert int(results["match.refuses_mask"].ravel()[0]) == 1


@pytest.mark.parametrize("with_landmarks", [False, True])
def test_stereo_tracker_class_equals_python_device_tracker(results, inputs, ctx, with_landmarks):
    """zenslam::cuda::stereo_tracker::track (+ add_landmarks / set_camera_center: assign_landmark_indices inside the step)
    against the python device tracker, which tests/test_gpu_landmarks.py ties to the oracle's assign_landmark_indices"""
    from zenslam_b200 import detection_options, keypoint, slam_options, tracking_options
    from zenslam_b200.keypoint_tracker import device_keypoint_tracker
    seq = inputs["frames"]
    opts = slam_options(detection=detection_options(cell_size=(CELL, CELL), fast_threshold=THR),
                        tracking=tracking_options(klt_window_size=(WIN, WIN), klt_max_level=LEVEL, filter_epipolar=False))
    keypoint.index_next = 0
    trk = device_keypoint_tracker(opts, ctx, W, H, landmark_capacity=4096 if with_landmarks else 0)
    pre = "trk_lm" if with_landmarks else "trk"
    if with_landmarks:
        n_lm = len(inputs["lm_index"])
        assert trk.add_landmarks(inputs["lm_index"], inputs["lm_xyz"], inputs["lm_desc"]) == n_lm
        assert results["trk_lm.added"].ravel().tolist() == [n_lm, 0]
        trk.set_camera_center([1.0, -2.0, 0.5])
    for t in range(FRAMES):
        maps = trk.track(seq[t, 0], seq[t, 1])
        for cam in range(2):
            f, i, d = (results["%s.%d.%d.%s" % (pre, t, cam, s)] for s in ("f", "i", "desc"))
            want = maps[cam]
            assert i[:, 0].tolist() == list(want), (t, cam)
            assert np.array_equal(f[:, :2], np.array([want[k].pt for k in want], np.float32)), (t, cam)
            assert np.array_equal(f[:, 2], np.array([want[k].response for k in want], np.float32)), (t, cam)
            assert np.array_equal(d, np.stack([want[k].descriptor for k in want])), (t, cam)
        assert int(results["%s.%d.index_next" % (pre, t)].ravel()[0]) == keypoint.index_next
    last = results["%s.%d.0.i" % (pre, FRAMES - 1)][:, 0]
    assert len(last) > 200
    assert bool((last >= 700000).any()) == with_landmarks           # landmark indices live in the maps exactly when a store exists
    trk.close()


def test_process_image_and_triangulation_equal_oracle(results, inputs):
    """zenslam::cuda::process_image (processor.cpp:25-55: BGR2GRAY, CLAHE 4.0, remap) bit-exact against the oracle, and
    zenslam::cuda::triangulate_points (triangulator.cpp:39-132) against the oracle within the stated float tolerance"""
    bgr = inputs["pp_bgr"]
    gray = oracle.bgr2gray(bgr)
    assert np.array_equal(results["pp.plain"], gray)
    want = oracle.remap_linear(oracle.clahe(gray, 4.0), inputs["pp_map_x"], inputs["pp_map_y"])
    assert np.array_equal(results["pp.full"], want)
    assert (want == 0).any() and not np.array_equal(want, gray)           # the maps really leave the frame / move pixels
    P = inputs["tri_P"]
    oxyz, okeep, odiag = oracle.triangulate_keypoints(P[0], P[1], inputs["tri_F"], inputs["tri_t"], inputs["tri_pts0"], inputs["tri_pts1"])
    xyz = results["tri.xyz"]
    scale = np.maximum(np.abs(oxyz).max(1, keepdims=True), 1e-6)
    assert (np.abs(xyz - oxyz) / scale).max() < 1e-5
    n0 = np.linalg.norm(oxyz, axis=1)
    margin = np.min(np.stack([np.abs(np.abs(odiag[:, 0]) - 0.01), np.abs(n0 - 1.0), np.abs(n0 - 50.0), np.abs(odiag[:, 1] - 1.0),
                              np.abs(odiag[:, 2] - 1.0), np.abs(odiag[:, 3] - 0.25), np.abs(oxyz[:, 2])]), 0)
    safe = margin > 1e-6
    keep = results["tri.keep"].ravel().astype(bool)
    assert safe.mean() > 0.99 and np.array_equal(keep[safe], okeep[safe]) and keep.sum() > 20
    wide = results["tri.keep_no_epipolar"].ravel().astype(bool)
    assert (wide | ~keep).all() and wide.sum() >= keep.sum()              # without the epipolar filter: a superset
