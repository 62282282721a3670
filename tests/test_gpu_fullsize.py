"""GPU: the CUDA path (through the C ABI) against cv2 itself at the BASELINE frame sizes and on the reference's shipped
tumvi.yaml settings -- the committed fixtures tests/golden/fullsize_*.npz (written from cv2 by
tests/golden/make_golden_fullsize.py) and, where cv2 imports on the box, the whole C2 stereo frame live.

Bars (north_star): keypoints, responses, descriptors, match indices BIT-EXACT; LK status flags and forward-backward keep
decisions IDENTICAL (0 mismatches); LK positions within 0.01 px of cv2 (and bit-identical to the oracle)."""
import zlib

import numpy as np
import pytest

import oracle
from fullsize_cases import CASES, frames
from zenslam_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

LK_TOL_PX = 0.01


@pytest.fixture(scope="module")
def ctx():
    from zenslam_b200.runtime import Context
    c = Context()
    yield c
    c.close()


def dev(ctx, a, dtype=None):
    return ctx.to_device(np.ascontiguousarray(a), dtype)


def load(golden, name):
    g = golden("fullsize_" + name)
    imgs = frames(name)
    assert [zlib.crc32(np.ascontiguousarray(i).tobytes()) for i in imgs] == g["crc"].tolist(), "synthetic generator drifted"
    return g, imgs


@pytest.mark.parametrize("name", list(CASES))
def test_detect_describe_match_fullsize_vs_cv2(ctx, golden, name):
    from zenslam_b200.runtime import (Pyramid, corner_subpix, fast_grid_detect, match_hamming_cross, match_hamming_knn2,
                                      orb_compute)
    g, (L0, R0, _) = load(golden, name)
    (w, h), cell, thr, subpix = CASES[name][0], CASES[name][2], CASES[name][3], CASES[name][4]
    win, ml = CASES[name][5][0]
    p = Pyramid(ctx, w, h, 2, win, ml)
    p.upload(np.stack([L0, R0]), 0); p.build(0, 2)
    xy, resp, n = fast_grid_detect(p, 0, 2, cell, thr)
    nn = n.cpu().numpy()
    for i, cam in enumerate("lr"):
        a = xy[i, :nn[i]].cpu().numpy()
        assert np.array_equal(a[:, 0], g[f"cv_grid_{cam}_x"]) and np.array_equal(a[:, 1], g[f"cv_grid_{cam}_y"])
        assert np.array_equal(resp[i, :nn[i]].cpu().numpy(), g[f"cv_grid_{cam}_r"])
    if subpix:
        corner_subpix(p, 0, 2, xy, n)
        for i, cam in enumerate("lr"):
            a = xy[i, :nn[i]].cpu().numpy()
            assert np.array_equal(a[:, 0], g[f"cv_subpix_{cam}_x"]) and np.array_equal(a[:, 1], g[f"cv_subpix_{cam}_y"])
    oxy, oresp, src, on, desc = orb_compute(p, 0, 2, xy, resp, n)
    on_h = on.cpu().numpy()
    for i, cam in enumerate("lr"):
        a = oxy[i, :on_h[i]].cpu().numpy()
        assert np.array_equal(a[:, 0], g[f"cv_orb_{cam}_x"]) and np.array_equal(a[:, 1], g[f"cv_orb_{cam}_y"])
        assert np.array_equal(desc[i, :on_h[i]].cpu().numpy(), g[f"cv_orb_{cam}_desc"])
    idx, dist, ps = match_hamming_knn2(ctx, desc[0:1], on[0:1], desc[1:2], on[1:2], 0.8)
    nl = on_h[0]
    assert np.array_equal(idx[0, :nl].cpu().numpy(), g["cv_knn_idx"])
    assert np.array_equal(dist[0, :nl].cpu().numpy(), g["cv_knn_dist"])
    assert np.array_equal(np.nonzero(ps[0, :nl].cpu().numpy())[0], g["cv_ratio_q"])
    cidx, _ = match_hamming_cross(ctx, desc[0:1], on[0:1], desc[1:2], on[1:2])
    cidx = cidx[0, :nl].cpu().numpy()
    keepq = np.nonzero(cidx >= 0)[0]
    assert np.array_equal(keepq, g["cv_cross_q"]) and np.array_equal(cidx[keepq], g["cv_cross_t"])
    p.close()


def lk_cases():
    for name, c in CASES.items():
        for win, ml in c[5]:
            yield name, win, ml


@pytest.mark.parametrize("name,win,ml", list(lk_cases()))
def test_lk_forward_backward_fullsize_vs_cv2(ctx, golden, name, win, ml):
    """forward + backward + gate in one launch (zs_klt_track_fb) for the three jobs of the fixture"""
    from zenslam_b200.runtime import LK, Pyramid, klt_track
    g, (L0, R0, L1) = load(golden, name)
    (w, h) = CASES[name][0]
    p = Pyramid(ctx, w, h, 3, win, ml)
    p.upload(np.stack([L0, R0, L1]), 0); p.build(0, 3)
    jobs = [("temporal", 0, 2), ("stereo", 0, 1), ("stereo_rev", 1, 0)]
    cap = max(len(g[f"pts_{j}"]) for j, _, _ in jobs)
    pts = np.zeros((3, cap, 2), np.float32)
    cnt = np.zeros(3, np.int32)
    for i, (j, _, _) in enumerate(jobs):
        q = g[f"pts_{j}"]
        pts[i, :len(q)] = q; cnt[i] = len(q)
    lk = LK(win, ml, 99, 0.001)
    O = {0: oracle.Pyramid(L0, win, ml), 1: oracle.Pyramid(R0, win, ml), 2: oracle.Pyramid(L1, win, ml)}
    total = 0
    for thr, kk in ((1.0, "cv_keep1_"), (2.0, "cv_keep2_")):
        p1, st, err, keep = klt_track(p, dev(ctx, np.array([a for _, a, _ in jobs], np.int32)),
                                      dev(ctx, np.array([b for _, _, b in jobs], np.int32)), dev(ctx, pts), dev(ctx, cnt), lk,
                                      None, thr)
        p1, st, err, keep = [a.cpu().numpy() for a in (p1, st, err, keep)]
        for i, (j, a, b) in enumerate(jobs):
            k = f"{j}_w{win[0]}_l{ml}"
            n = cnt[i]
            assert np.array_equal(st[i, :n], g[f"cv_fst_{k}"]), "forward status flags differ from cv2"
            ok = st[i, :n] > 0
            assert np.abs(p1[i, :n] - g[f"cv_fwd_{k}"])[ok].max() < LK_TOL_PX
            assert np.allclose(err[i, :n], g[f"cv_err_{k}"], rtol=1e-4, atol=1e-6)
            assert np.array_equal(keep[i, :n], g[kk + k]), "forward-backward keep decisions differ from cv2"
            if thr == 1.0:
                # and bit-identical to the oracle (same exact integer sums)
                o1, os_, oe = oracle.lk_track(O[a], O[b], pts[i, :n], None, win, ml)
                assert np.array_equal(p1[i, :n], o1) and np.array_equal(st[i, :n], os_) and np.array_equal(err[i, :n], oe)
                total += n
    assert total >= 4000
    p.close()


def test_c2_stereo_frames_live_vs_cv2(ctx):
    """The whole BASELINE configs[1] stereo frame -- oracle/cv2_ref.stereo_frame, i.e. the reference's call pattern over
    real cv2 calls -- against zs_frontend_process_host for 3 consecutive frames (skipped where cv2 is absent)."""
    pytest.importorskip("cv2")
    from oracle import cv2_ref
    from zenslam_b200 import detection_options, slam_options, tracking_options
    from zenslam_b200.frontend import KINDS, StereoFrontend
    w, h, B = 752, 480, 4
    opts = slam_options(matcher="KNN", detection=detection_options(), tracking=tracking_options())
    seq, _ = syn.stereo_sequence(w, h, B, 4242, subpixel=True)
    fe = StereoFrontend(ctx, w, h, B, opts)
    res = fe.process(np.ascontiguousarray(seq[:, 0]), np.ascontiguousarray(seq[:, 1]))
    fo = oracle.FrontendOptions()
    prev = None
    checked = tracks = 0
    for k in range(B):
        L, R = seq[k, 0], seq[k, 1]
        if prev is None:
            kl = np.zeros((0, 2), np.float32); kr = kl
            ref = cv2_ref.stereo_frame(L, R, L, R, kl, kr, fo)
        else:
            ref = cv2_ref.stereo_frame(prev[0], prev[1], L, R, prev[2], prev[3], fo)
        nl, nr = int(res["n_left"][k]), int(res["n_right"][k])
        assert nl == len(ref["kp_l"]) and nr == len(ref["kp_r"])
        assert np.array_equal(res["kp_left"][k, :nl], ref["kp_l"]) and np.array_equal(res["kp_right"][k, :nr], ref["kp_r"])
        assert np.array_equal(res["resp_left"][k, :nl], ref["resp_l"]) and np.array_equal(res["resp_right"][k, :nr], ref["resp_r"])
        assert np.array_equal(res["desc_left"][k, :nl], ref["desc_l"]) and np.array_equal(res["desc_right"][k, :nr], ref["desc_r"])
        got = [(i, int(res["match_idx"][k, i, 0])) for i in np.nonzero(res["match_pass"][k, :nl])[0]]
        assert got == [(q, t) for q, t, _ in ref["matches"]]
        for kind, key in zip(KINDS, ("temporal_l", "temporal_r", "stereo_lr", "stereo_rl")):
            if prev is None and kind.startswith("temporal"):
                continue
            ki = KINDS.index(kind)
            p1, keep = ref[key]
            n = len(p1)
            assert int(res["track_n"][ki, k]) == n
            assert np.array_equal(res["track_keep"][ki, k, :n].astype(bool), keep), (k, kind)
            if keep.any():
                assert np.abs(res["track_pts"][ki, k, :n] - p1)[keep].max() < LK_TOL_PX
            tracks += n
        prev = (L, R, ref["kp_l"], ref["kp_r"])
        checked += 1
    assert checked >= 3 and tracks > 10000
    fe.close()
