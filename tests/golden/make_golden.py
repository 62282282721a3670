"""Generate the committed golden fixtures from the REAL dependency (cv2).

Run here (cv2 4.13.0 importable):  python tests/golden/make_golden.py
Writes tests/golden/*.npz.  Inputs are stored with the outputs so that the tests need neither
cv2 nor the generator.  Every array that OpenCV produced is prefixed ``cv_``.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import cv2  # noqa: E402

from oracle import cv2_ref  # noqa: E402
from zenslam_b200 import synthetic as syn  # noqa: E402


def kps_arrays(kps):
    return (np.array([k.pt[0] for k in kps], np.float32), np.array([k.pt[1] for k in kps], np.float32),
            np.array([k.response for k in kps], np.float32))


def main():
    meta = dict(cv2_version=cv2.__version__)
    W, H = 256, 192
    base = syn.base_texture(W, H, 4242)
    L = syn.crop(base, W, H, 0, 0)
    R = syn.crop(base, W, H, 6, 0)
    N = syn.crop(base, W, H, 3.5, -2.25)      # sub-pixel moved "next" frame
    Lq = (L // 4 * 4).astype(np.uint8)        # quantised: many tied per-cell maxima

    # ---- pyramid ---------------------------------------------------------------------------
    out = dict(L=L, R=R, N=N)
    for win, ml in (((15, 15), 3), ((31, 31), 3), ((63, 63), 4)):
        mlv, pyr = cv2.buildOpticalFlowPyramid(L, win, ml)
        out[f"cv_levels_w{win[0]}"] = np.int32(mlv + 1)
        if win[0] == 15:
            for l in range(mlv + 1):
                out[f"cv_pyr_img{l}"] = np.ascontiguousarray(pyr[2 * l])
                out[f"cv_pyr_der{l}"] = np.ascontiguousarray(pyr[2 * l + 1])
    np.savez_compressed(os.path.join(HERE, "pyramid.npz"), **out)

    # ---- FAST / grid / ORB -----------------------------------------------------------------
    out = dict(L=L, Lq=Lq)
    for thr in (1, 10, 40):
        x, y, r = kps_arrays(cv2.FastFeatureDetector_create(thr).detect(L, None))
        out[f"cv_fast_t{thr}_x"], out[f"cv_fast_t{thr}_y"], out[f"cv_fast_t{thr}_r"] = x, y, r
    for name, img in (("L", L), ("Lq", Lq)):
        for cell, thr in (((16, 16), 10), ((32, 32), 10), ((64, 64), 1), ((24, 16), 5)):
            x, y, r = cv2_ref.grid_detect(img, cell, thr)
            k = f"cv_grid_{name}_c{cell[0]}x{cell[1]}_t{thr}"
            out[k + "_x"], out[k + "_y"], out[k + "_r"] = x, y, r
    occ = np.zeros((H // 16, W // 16), np.uint8)
    occ[::2, 1::3] = 1
    x, y, r = cv2_ref.grid_detect(L, (16, 16), 10, occ)
    out["occ"] = occ
    out["cv_grid_occ_x"], out["cv_grid_occ_y"], out["cv_grid_occ_r"] = x, y, r
    gx, gy, _ = cv2_ref.grid_detect(L, (16, 16), 10)
    kx, ky, desc = cv2_ref.orb_compute(L, gx, gy)
    out["cv_orb_kx"], out["cv_orb_ky"], out["cv_orb_desc"] = kx, ky, desc
    g = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
    out["cv_orb_blur"] = cv2.sepFilter2D(L, cv2.CV_8U, g, g, borderType=cv2.BORDER_REFLECT_101)
    rng = np.random.default_rng(77)
    fx = rng.uniform(20, W - 20, 300).astype(np.float32)
    fy = rng.uniform(20, H - 20, 300).astype(np.float32)
    fx[:8] = [30.5, 31.5, 31.49, 224.5, 225.5, 100, 100, 100]
    fy[:8] = [100, 100, 100, 100, 100, 30.5, 160.5, 161.5]
    fa = rng.uniform(0, 360, 300).astype(np.float32)
    kx, ky, desc = cv2_ref.orb_compute(L, fx, fy, fa)
    out["orb_in_x"], out["orb_in_y"], out["orb_in_a"] = fx, fy, fa
    out["cv_orb_rot_kx"], out["cv_orb_rot_ky"], out["cv_orb_rot_desc"] = kx, ky, desc
    np.savez_compressed(os.path.join(HERE, "detect.npz"), **out)

    # ---- matching --------------------------------------------------------------------------
    gxr, gyr, _ = cv2_ref.grid_detect(R, (16, 16), 10)
    _, _, dl = cv2_ref.orb_compute(L, gx, gy)
    _, _, dr = cv2_ref.orb_compute(R, gxr, gyr)
    rng = np.random.default_rng(78)
    q16 = np.zeros((150, 32), np.uint8); q16[:, :2] = rng.integers(0, 256, (150, 2))
    t16 = np.zeros((170, 32), np.uint8); t16[:, :2] = rng.integers(0, 256, (170, 2))
    t16[11] = q16[5]; t16[90] = q16[5]; t16[40] = t16[41]
    out = dict(dl=dl, dr=dr, q16=q16, t16=t16)

    def knn_arr(q, t, norm):
        bf = cv2.BFMatcher(norm, False)
        idx = np.full((len(q), 2), -1, np.int32); dist = np.zeros((len(q), 2), np.float32)
        for i, k in enumerate(bf.knnMatch(q, t, 2)):
            for j, m in enumerate(k):
                idx[i, j] = m.trainIdx; dist[i, j] = m.distance
        return idx, dist

    def tri(ms):
        return (np.array([m[0] for m in ms], np.int32), np.array([m[1] for m in ms], np.int32),
                np.array([m[2] for m in ms], np.float32))

    for nm, q, t in (("orb", dl, dr), ("b16", q16, t16), ("one", dl, dr[:1])):
        out[f"cv_knn_{nm}_idx"], out[f"cv_knn_{nm}_dist"] = knn_arr(q, t, cv2.NORM_HAMMING)
        out[f"cv_cross_{nm}_q"], out[f"cv_cross_{nm}_t"], out[f"cv_cross_{nm}_d"] = tri(cv2_ref.match_cross(q, t))
        out[f"cv_ratio_{nm}_q"], out[f"cv_ratio_{nm}_t"], out[f"cv_ratio_{nm}_d"] = tri(
            cv2_ref.match_knn_ratio(q, t, 0.8))
    sift = cv2.SIFT_create(nfeatures=300)
    _, s0 = sift.detectAndCompute(L, None)
    _, s1 = sift.detectAndCompute(R, None)
    assert np.all(s0 == np.rint(s0)) and s0.max() <= 255
    out["sift0"], out["sift1"] = s0.astype(np.uint8), s1.astype(np.uint8)
    out["cv_knn_sift_idx"], out["cv_knn_sift_dist"] = knn_arr(s0, s1, cv2.NORM_L2)
    out["cv_cross_sift_q"], out["cv_cross_sift_t"], out["cv_cross_sift_d"] = tri(cv2_ref.match_cross(s0, s1, "l2"))
    out["cv_ratio_sift_q"], out["cv_ratio_sift_t"], out["cv_ratio_sift_d"] = tri(
        cv2_ref.match_knn_ratio(s0, s1, 0.8, "l2"))
    np.savez_compressed(os.path.join(HERE, "match.npz"), **out)

    # ---- KLT -------------------------------------------------------------------------------
    rng = np.random.default_rng(79)
    n = 400
    pts = np.stack([rng.uniform(-12, W + 12, n), rng.uniform(-12, H + 12, n)], 1).astype(np.float32)
    pts[:40] = np.rint(pts[:40])
    Lb = L.copy(); Nb = N.copy()
    Lb[60:110, 60:130] = 90; Nb[60:110, 60:130] = 90        # textureless block -> min-eig rejections
    init = pts - np.array([3.5, -2.25], np.float32) + rng.normal(0, 1.5, (n, 2)).astype(np.float32)
    init[::9] += 30
    out = dict(A=Lb, B=Nb, pts=pts, init=init)
    for win, ml in (((15, 15), 3), ((21, 21), 2), ((31, 31), 3), ((31, 31), 0), ((63, 63), 3), ((31, 21), 3)):
        k = f"w{win[0]}x{win[1]}_l{ml}"
        p1, st, err = cv2_ref.lk(Lb, Nb, pts, None, win, ml)
        out[f"cv_lk_{k}_p1"], out[f"cv_lk_{k}_st"], out[f"cv_lk_{k}_err"] = p1, st, err
        p1, st, err = cv2_ref.lk(Lb, Nb, pts, init, win, ml)
        out[f"cv_lki_{k}_p1"], out[f"cv_lki_{k}_st"], out[f"cv_lki_{k}_err"] = p1, st, err
    p1, keep = cv2_ref.track_fb(Lb, Nb, pts, None, (31, 31), 3, 1.0)
    out["cv_fb_p1"], out["cv_fb_keep"] = p1, keep
    np.savez_compressed(os.path.join(HERE, "klt.npz"), **out)

    with open(os.path.join(HERE, "README.md"), "w") as f:
        f.write("Golden fixtures written by make_golden.py from cv2 %s (opencv-python-headless).\n"
                "Arrays prefixed `cv_` are OpenCV outputs; the rest are the inputs they were computed from.\n"
                % meta["cv2_version"])
    for fn in sorted(os.listdir(HERE)):
        if fn.endswith(".npz"):
            print(fn, os.path.getsize(os.path.join(HERE, fn)))


if __name__ == "__main__":
    main()
