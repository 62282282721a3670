"""Golden fixture for cv::triangulatePoints as utils::triangulate_points calls it (mapping/triangulation_utils.cpp:135-160).
Run here:  python tests/golden/make_golden_triangulate.py  -> tests/golden/triangulate.npz"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import cv2  # noqa: E402


def rig():
    K = np.array([[458.0, 0, 367.2], [0, 457.3, 248.4], [0, 0, 1]])
    R = cv2.Rodrigues(np.array([0.002, np.deg2rad(0.4), -0.003]))[0]
    tc = np.array([0.11, 0.001, -0.002])                       # camera 1 centre in camera 0
    P0 = K @ np.hstack([np.eye(3), np.zeros((3, 1))])
    P1 = K @ np.hstack([R.T, (-R.T @ tc)[:, None]])
    tx = np.array([[0, -tc[2], tc[1]], [tc[2], 0, -tc[0]], [-tc[1], tc[0], 0]])
    E = R.T @ tx.T                                             # x1^T E x0 = 0 for X1 = R^T (X0 - tc)
    Ki = np.linalg.inv(K)
    F = Ki.T @ E @ Ki
    return K, P0, P1, F / np.abs(F).max() * 0.05, tc      # scale: good pairs land around 1e-3, the reference threshold is 0.01


def main():
    rng = np.random.default_rng(3)
    W, H = 752, 480
    K, P0, P1, F, tc = rig()
    n = 2000
    Z = rng.uniform(0.5, 80, n)
    X = np.stack([(rng.uniform(0, W, n) - K[0, 2]) / K[0, 0] * Z, (rng.uniform(0, H, n) - K[1, 2]) / K[1, 1] * Z, Z], 1)

    def proj(P, X):
        q = (P @ np.hstack([X, np.ones((len(X), 1))]).T).T
        return q[:, :2] / q[:, 2:]
    noise = np.where(rng.random((n, 1)) < 0.3, 1.5, 0.05)       # 30 % of the pairs are poor matches
    p0 = (proj(P0, X) + rng.normal(0, 1, (n, 2)) * noise).astype(np.float32)
    p1 = (proj(P1, X) + rng.normal(0, 1, (n, 2)) * noise).astype(np.float32)
    p1[:5] = p0[:5]                                             # zero disparity: points at infinity
    X4 = cv2.triangulatePoints(P0, P1, p0.T.copy(), p1.T.copy())
    np.savez_compressed(os.path.join(HERE, "triangulate.npz"), P0=P0, P1=P1, F=F, t=tc, pts0=p0, pts1=p1, cv_points4d=X4,
                        cv2_version=cv2.__version__)
    print("wrote triangulate.npz", X4.dtype, X4.shape)


if __name__ == "__main__":
    main()
