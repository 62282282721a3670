"""Pin the C oracle (oracle/zs_oracle.c) against the committed cv2 golden vectors (CPU only)."""
import numpy as np
import pytest

import oracle


def test_pyramid_levels_and_planes(golden):
    g = golden("pyramid")
    L = g["L"]
    h, w = L.shape
    assert oracle.pyramid_num_levels(w, h, (15, 15), 3) == int(g["cv_levels_w15"])
    assert oracle.pyramid_num_levels(w, h, (31, 31), 3) == int(g["cv_levels_w31"])
    assert oracle.pyramid_num_levels(w, h, (63, 63), 4) == int(g["cv_levels_w63"])
    P = oracle.Pyramid(L, (15, 15), 3)
    assert P.levels == int(g["cv_levels_w15"])
    for l in range(P.levels):
        assert np.array_equal(P.image(l), g[f"cv_pyr_img{l}"])
        assert np.array_equal(P.deriv(l), g[f"cv_pyr_der{l}"])


@pytest.mark.parametrize("thr", [1, 10, 40])
def test_fast_full_frame(golden, thr):
    g = golden("detect")
    x, y, s = oracle.fast_detect(g["L"], thr)
    assert np.array_equal(x, g[f"cv_fast_t{thr}_x"])
    assert np.array_equal(y, g[f"cv_fast_t{thr}_y"])
    assert np.array_equal(s, g[f"cv_fast_t{thr}_r"])


@pytest.mark.parametrize("name", ["L", "Lq"])
@pytest.mark.parametrize("cell,thr", [((16, 16), 10), ((32, 32), 10), ((64, 64), 1), ((24, 16), 5)])
def test_grid_detect(golden, name, cell, thr):
    g = golden("detect")
    x, y, s = oracle.grid_detect(g[name], cell, thr)
    k = f"cv_grid_{name}_c{cell[0]}x{cell[1]}_t{thr}"
    assert np.array_equal(x, g[k + "_x"]) and np.array_equal(y, g[k + "_y"]) and np.array_equal(s, g[k + "_r"])


def test_grid_detect_occupancy(golden):
    g = golden("detect")
    x, y, s = oracle.grid_detect(g["L"], (16, 16), 10, g["occ"])
    assert np.array_equal(x, g["cv_grid_occ_x"]) and np.array_equal(y, g["cv_grid_occ_y"])
    assert np.array_equal(s, g["cv_grid_occ_r"])


def test_orb_blur_and_descriptors(golden):
    g = golden("detect")
    L = g["L"]
    assert np.array_equal(oracle.orb_blur(L), g["cv_orb_blur"])
    x, y, _ = oracle.grid_detect(L, (16, 16), 10)
    kept, desc = oracle.orb_compute(L, x, y)
    assert np.array_equal(x[kept].astype(np.float32), g["cv_orb_kx"])
    assert np.array_equal(y[kept].astype(np.float32), g["cv_orb_ky"])
    assert np.array_equal(desc, g["cv_orb_desc"])


def test_orb_rotated_subpixel(golden):
    g = golden("detect")
    kept, desc = oracle.orb_compute(g["L"], g["orb_in_x"], g["orb_in_y"], g["orb_in_a"])
    assert np.array_equal(g["orb_in_x"][kept], g["cv_orb_rot_kx"])
    assert np.array_equal(g["orb_in_y"][kept], g["cv_orb_rot_ky"])
    assert np.array_equal(desc, g["cv_orb_rot_desc"])


@pytest.mark.parametrize("nm,qk,tk", [("orb", "dl", "dr"), ("b16", "q16", "t16"), ("one", "dl", "dr")])
def test_hamming_matching(golden, nm, qk, tk):
    g = golden("match")
    q, t = g[qk], g[tk]
    if nm == "one":
        t = t[:1]
    idx, dist = oracle.match_hamming_knn2(q, t)
    assert np.array_equal(idx, g[f"cv_knn_{nm}_idx"])
    valid = idx >= 0
    assert np.array_equal(dist.astype(np.float32)[valid], g[f"cv_knn_{nm}_dist"][valid])
    oq, ot, od = oracle.match_hamming_cross(q, t)
    assert np.array_equal(oq, g[f"cv_cross_{nm}_q"]) and np.array_equal(ot, g[f"cv_cross_{nm}_t"])
    assert np.array_equal(od.astype(np.float32), g[f"cv_cross_{nm}_d"])
    rq, rt, rd = oracle.ratio_test(idx, dist.astype(np.float32), 0.8)
    assert np.array_equal(rq, g[f"cv_ratio_{nm}_q"]) and np.array_equal(rt, g[f"cv_ratio_{nm}_t"])
    assert np.array_equal(rd, g[f"cv_ratio_{nm}_d"])


def test_l2_matching_sift(golden):
    g = golden("match")
    q, t = g["sift0"].astype(np.float32), g["sift1"].astype(np.float32)
    idx, dist = oracle.match_l2_knn2(q, t)
    assert np.array_equal(idx, g["cv_knn_sift_idx"])
    assert np.array_equal(dist, g["cv_knn_sift_dist"])          # bit-exact distances
    oq, ot, od = oracle.match_l2_cross(q, t)
    assert np.array_equal(oq, g["cv_cross_sift_q"]) and np.array_equal(ot, g["cv_cross_sift_t"])
    assert np.array_equal(od, g["cv_cross_sift_d"])
    rq, rt, rd = oracle.ratio_test(idx, dist, 0.8)
    assert np.array_equal(rq, g["cv_ratio_sift_q"]) and np.array_equal(rt, g["cv_ratio_sift_t"])


LK_CASES = [((15, 15), 3), ((21, 21), 2), ((31, 31), 3), ((31, 31), 0), ((63, 63), 3), ((31, 21), 3)]
LK_TOL = 0.01   # px, the north-star tolerance; status flags must be identical


@pytest.mark.parametrize("win,ml", LK_CASES)
@pytest.mark.parametrize("init", [False, True])
def test_lk(golden, win, ml, init):
    g = golden("klt")
    PA, PB = oracle.Pyramid(g["A"], win, ml), oracle.Pyramid(g["B"], win, ml)
    k = f"{'cv_lki' if init else 'cv_lk'}_w{win[0]}x{win[1]}_l{ml}"
    flags = oracle.LK_GET_MIN_EIGENVALS | (oracle.LK_USE_INITIAL_FLOW if init else 0)
    p1, st, err = oracle.lk_track(PA, PB, g["pts"], g["init"] if init else None, win, ml, flags=flags)
    assert np.array_equal(st, g[k + "_st"])
    ok = st > 0
    assert np.abs(p1 - g[k + "_p1"])[ok].max() < LK_TOL
    assert np.allclose(err, g[k + "_err"], rtol=1e-4, atol=1e-6)


def test_fb_gate(golden):
    g = golden("klt")
    win, ml = (31, 31), 3
    PA, PB = oracle.Pyramid(g["A"], win, ml), oracle.Pyramid(g["B"], win, ml)
    p1, st, _ = oracle.lk_track(PA, PB, g["pts"], None, win, ml)
    pb, sb, _ = oracle.lk_track(PB, PA, p1, None, win, ml)
    keep = oracle.fb_check(g["pts"], pb, st, sb, 1.0)
    assert np.array_equal(keep, g["cv_fb_keep"])
    assert np.abs(p1 - g["cv_fb_p1"])[keep].max() < LK_TOL


def test_empty_inputs():
    idx, dist = oracle.match_hamming_knn2(np.zeros((0, 32), np.uint8), np.zeros((5, 32), np.uint8))
    assert idx.shape == (0, 2)
    oq, _, _ = oracle.match_hamming_cross(np.zeros((3, 32), np.uint8), np.zeros((0, 32), np.uint8))
    assert len(oq) == 0
    x, y, s = oracle.grid_detect(np.full((64, 64), 7, np.uint8), (16, 16), 10)
    assert len(x) == 0
    P = oracle.Pyramid(np.zeros((64, 64), np.uint8), (15, 15), 3)
    p1, st, err = oracle.lk_track(P, P, np.zeros((0, 2), np.float32))
    assert p1.shape == (0, 2)


def test_corner_subpix_golden(golden):
    """cv::cornerSubPix (PARALLEL_GRID detector): the C restatement is bit-exact against cv2 with IPP off on every
    fixture point, image-border points included, and within 5e-3 px of cv2's IPP path."""
    g = golden("subpix")
    got = oracle.corner_subpix(g["L"], g["pts"])
    assert np.array_equal(got, g["cv_subpix"])
    assert np.abs(got - g["cv_subpix_ipp"]).max() < 5e-3
    # degenerate inputs
    assert oracle.corner_subpix(g["L"], np.zeros((0, 2), np.float32)).shape == (0, 2)
    flat = np.full((64, 64), 77, np.uint8)
    p = np.array([[20.5, 30.25]], np.float32)
    assert np.array_equal(oracle.corner_subpix(flat, p), p)            # singular normal matrix: point unchanged


def test_preprocessing_golden(golden):
    """processor::process image path: BGR->gray, CLAHE(4.0), remap(INTER_LINEAR) -- bit-exact against cv2"""
    g = golden("preproc")
    gray = oracle.bgr2gray(g["bgr"])
    assert np.array_equal(gray, g["cv_gray"])
    cl = oracle.clahe(gray, 4.0)
    assert np.array_equal(cl, g["cv_clahe"])
    assert np.array_equal(oracle.clahe(gray, 2.0, (4, 6)), g["cv_clahe_2_4x6"])
    assert np.array_equal(oracle.remap_linear(gray, g["map_x"], g["map_y"]), g["cv_remap_gray"])
    assert np.array_equal(oracle.remap_linear(cl, g["map_x"], g["map_y"]), g["cv_remap_clahe"])


def _points3d_from_cv(X4):
    """utils::triangulate_points' conversion (triangulation_utils.cpp:150-157)"""
    X4 = X4.astype(np.float64)
    ok = np.abs(X4[3]) > 1e-9
    return np.where(ok, X4[:3] / np.where(ok, X4[3], 1.0), 0.0).T


def test_triangulate_golden(golden):
    """cv::triangulatePoints restated (DLT + OpenCV's one-sided Jacobi SVD, float output): identical to cv2 on the fixture;
    the gates behave as triangulator.cpp:118-127 says"""
    g = golden("triangulate")
    xyz, keep, diag = oracle.triangulate_keypoints(g["P0"], g["P1"], g["F"], g["t"], g["pts0"], g["pts1"])
    want = _points3d_from_cv(g["cv_points4d"])
    assert np.array_equal(xyz, want)
    n0 = np.linalg.norm(xyz, axis=1)
    ref_keep = ((np.abs(diag[:, 0]) < 0.01) & (xyz[:, 2] > 0) & (n0 > 1.0) & (n0 < 50.0) & (diag[:, 1] < 1.0) & (diag[:, 2] < 1.0) &
                (diag[:, 3] > 0.25) & (diag[:, 3] < 179.75))
    assert np.array_equal(keep, ref_keep) and 50 < keep.sum() < len(keep)
    # without the epipolar filter more pairs survive, and none is lost
    _, keep2, _ = oracle.triangulate_keypoints(g["P0"], g["P1"], None, g["t"], g["pts0"], g["pts1"])
    assert keep2.sum() >= keep.sum() and np.all(keep2[keep])


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_orb_multiscale_detector_golden(golden, case):
    """`feature: ORB` (keypoint_detector_simple.cpp:17,49,54): the C restatement of cv::ORB::detect (8-level
    INTER_LINEAR_EXACT pyramid, FAST + mask + border, retainBest by FAST score, Harris, retainBest, IC angle) and of
    cv::ORB::compute on those multi-scale keypoints reproduces cv2 bit for bit -- positions, sizes, angles, Harris
    responses, octaves, descriptors -- as a set in canonical (octave, y, x) order"""
    g = golden("orb_detect")
    mask = g[case + "_mask"] if (case + "_mask") in g.files else None
    o = oracle.orb_detect(g[case + "_img"], mask, fast_threshold=int(g[case + "_thr"]))
    for k in ("x", "y", "size", "angle", "response", "octave", "desc"):
        assert np.array_equal(o[k], g[case + "_" + k]), k


def test_resize_linear_exact_and_fast_atan2_golden(golden):
    g = golden("orb_detect")
    l1 = oracle.resize_linear_exact(g["a_img"], 313, 200)
    assert np.array_equal(l1, g["resize_313x200"])
    assert np.array_equal(oracle.resize_linear_exact(l1, 261, 167), g["resize_261x167"])
    got = np.array([oracle.fast_atan2(y, x) for y, x in g["atan_yx"]], np.float32)
    assert np.array_equal(got, g["atan_deg"])
