"""Golden fixtures at the BASELINE frame sizes and on the reference's shipped configuration, written from the
REAL dependency (cv2).

    python tests/golden/make_golden_fullsize.py          (cv2 4.13.0 importable here)

Writes tests/golden/fullsize_<name>.npz for
  * c2     752x480   GRID 16x16 thr 10 (BASELINE configs[1])            LK 31x31 / max_level 3
  * tumvi  752x480   GRID 64x64 thr 1 + cornerSubPix (tumvi.yaml:38-47)  LK 63x63 / max_level 4
  * c4     1280x1024 GRID 16x16 thr 10 (BASELINE configs[3])            LK 31x31 / 3 and 63x63 / 4
  * c5     3840x2160 GRID 32x32 thr 10 (BASELINE configs[4])            LK 31x31 / 3
Per case: grid corners of both cameras, ORB::compute rows, Hamming kNN-2 / ratio / cross-check of left vs right, and for
each LK setting the forward result, the backward result and the forward-backward keep flag of every track of
  temporal (left_0 -> left_1), stereo (left_0 -> right_0) and reverse stereo (right_0 -> left_0)
(keypoint_tracker.cpp:142-186, 379-423).  The frames are seeded synthetic images (zenslam_b200.synthetic, pure NumPy);
only the seed and a CRC of every frame are stored, the tests regenerate them.
Every array that OpenCV produced is prefixed ``cv_``.
"""
from __future__ import annotations

import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import cv2  # noqa: E402

from oracle import cv2_ref  # noqa: E402
from zenslam_b200 import synthetic as syn  # noqa: E402,F401

sys.path.insert(0, os.path.join(HERE, ".."))
from fullsize_cases import CASES, frames  # noqa: E402

def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def extra_tracks(w, h, n, seed):
    """points the detector never yields: near / outside the borders, integer and sub-pixel"""
    rng = np.random.default_rng(seed)
    p = np.stack([rng.uniform(-20, w + 20, n), rng.uniform(-20, h + 20, n)], 1).astype(np.float32)
    p[: n // 4] = np.rint(p[: n // 4])
    return p


def main(only=None):
    for name, ((w, h), seed, cell, thr, subpix, lks, max_tracks, n_extra) in CASES.items():
        if only and name not in only:
            continue
        L0, R0, L1 = frames(name)
        out = dict(crc=np.array([crc(L0), crc(R0), crc(L1)], np.uint32), cv2_version=np.bytes_(cv2.__version__))
        kp = {}
        for cam, img in (("l", L0), ("r", R0)):
            x, y, r = cv2_ref.grid_detect(img, cell, thr)
            out[f"cv_grid_{cam}_x"], out[f"cv_grid_{cam}_y"], out[f"cv_grid_{cam}_r"] = x, y, r
            if subpix:
                # keypoint_detector_parallel.cpp:160-170 -- pinned with IPP off (a vcpkg OpenCV has none)
                cv2.ipp.setUseIPP(False)
                c = np.stack([x, y], 1).astype(np.float32).reshape(-1, 1, 2).copy()
                cv2.cornerSubPix(img, c, (5, 5), (-1, -1), (cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 30, 0.01))
                cv2.ipp.setUseIPP(True)
                x, y = np.ascontiguousarray(c[:, 0, 0]), np.ascontiguousarray(c[:, 0, 1])
                out[f"cv_subpix_{cam}_x"], out[f"cv_subpix_{cam}_y"] = x, y
            kx, ky, desc = cv2_ref.orb_compute(img, x, y)
            out[f"cv_orb_{cam}_x"], out[f"cv_orb_{cam}_y"], out[f"cv_orb_{cam}_desc"] = kx, ky, desc
            kp[cam] = np.stack([kx, ky], 1).astype(np.float32)
        dl, dr = out["cv_orb_l_desc"], out["cv_orb_r_desc"]
        bf = cv2.BFMatcher(cv2.NORM_HAMMING, False)
        idx = np.full((len(dl), 2), -1, np.int32); dist = np.zeros((len(dl), 2), np.float32)
        for i, k in enumerate(bf.knnMatch(dl, dr, 2)):
            for j, m in enumerate(k):
                idx[i, j] = m.trainIdx; dist[i, j] = m.distance
        out["cv_knn_idx"], out["cv_knn_dist"] = idx, dist
        rt = cv2_ref.match_knn_ratio(dl, dr, 0.8)
        out["cv_ratio_q"], out["cv_ratio_t"] = (np.array([m[0] for m in rt], np.int32), np.array([m[1] for m in rt], np.int32))
        cr = cv2_ref.match_cross(dl, dr)
        out["cv_cross_q"], out["cv_cross_t"] = (np.array([m[0] for m in cr], np.int32), np.array([m[1] for m in cr], np.int32))

        # tracks: the detector's keypoints (what the reference tracks) + border / outside points
        ex = extra_tracks(w, h, n_extra, seed + 1)
        jobs = {"temporal": (L0, L1, kp["l"]), "stereo": (L0, R0, kp["l"]), "stereo_rev": (R0, L0, kp["r"])}
        for jn, (A, B, p) in jobs.items():
            if max_tracks and len(p) > max_tracks:
                sel = np.sort(np.random.default_rng(seed + 2).choice(len(p), max_tracks, replace=False))
                p = p[sel]
            p = np.concatenate([p, ex]).astype(np.float32)
            out[f"pts_{jn}"] = p
            for win, ml in lks:
                k = f"{jn}_w{win[0]}_l{ml}"
                p1, st, err = cv2_ref.lk(A, B, p, None, win, ml)
                pb, sb, _ = cv2_ref.lk(B, A, p1, None, win, ml)
                d = pb - p
                nrm = np.sqrt(d[:, 0].astype(np.float64) ** 2 + d[:, 1].astype(np.float64) ** 2)
                out[f"cv_fwd_{k}"], out[f"cv_fst_{k}"], out[f"cv_err_{k}"] = p1, st, err
                out[f"cv_bwd_{k}"], out[f"cv_bst_{k}"] = pb, sb
                # klt_threshold: 1.0 is the option default, 2 is what tumvi.yaml:47 ships
                out[f"cv_keep1_{k}"] = ((st != 0) & (sb != 0) & (nrm < 1.0)).astype(np.uint8)
                out[f"cv_keep2_{k}"] = ((st != 0) & (sb != 0) & (nrm < 2.0)).astype(np.uint8)
        fn = os.path.join(HERE, f"fullsize_{name}.npz")
        np.savez_compressed(fn, **out)
        ntr = sum(len(out[f"pts_{j}"]) for j in jobs) * len(lks)
        print(f"{name}: {len(kp['l'])}/{len(kp['r'])} keypoints, {ntr} tracks, {os.path.getsize(fn)} bytes")


if __name__ == "__main__":
    main(sys.argv[1:])
