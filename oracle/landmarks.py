"""Landmark association of the hot path's callers, restated on the CPU (NumPy) -- TEST INFRASTRUCTURE.

  * `assign_landmark_indices`  keypoint_tracker::assign_landmark_indices (zenslam_core/source/tracking/keypoint_tracker.cpp:199-291)
  * `match_keypoints3d`        utils::match_keypoints3d, both overloads (zenslam_core/source/matching/matching_utils.cpp:132-343)
  * `landmark_cloud`           the parts of point3d_cloud / map<point3d> they read (types/point3d_cloud.cpp:52-67, types/map.h:186-232)

One reference quirk is reproduced on purpose (the CUDA path does the same, INTEGRATION.md lists it): `point3d_cloud::radius_search`
asks nanoflann how many points lie within the radius and then returns the FIRST `count` points of the cloud in insertion order
(`this->operator()(i)` for i < count, point3d_cloud.cpp:61-64) -- not the points the search found.  The descriptor arithmetic is
the oracle's Hamming matcher (oracle.match_hamming_cross), pinned to cv2 elsewhere.
"""
from __future__ import annotations

import numpy as np


class landmark_cloud:
    """point3d_cloud as the path sees it: landmarks in insertion order; `+=` of a map keeps existing indices and appends the
    new ones in key order (map::operator+=(const map&), types/map.h:222-236; slam_thread.cpp:210)."""

    def __init__(self):
        self.index = np.zeros(0, np.int64)
        self.xyz = np.zeros((0, 3), np.float64)
        self.desc = np.zeros((0, 32), np.uint8)

    def __len__(self):
        return len(self.index)

    def add(self, index, xyz, desc):
        index = np.asarray(index, np.int64); xyz = np.asarray(xyz, np.float64).reshape(-1, 3)
        desc = np.asarray(desc, np.uint8).reshape(-1, 32)
        order = np.argsort(index, kind="stable")                 # `other` is a std::map: iterated in key order
        have = set(self.index.tolist())
        keep = []
        for i in order:
            if int(index[i]) not in have:
                have.add(int(index[i])); keep.append(i)
        keep = np.array(keep, np.int64)
        if len(keep):
            self.index = np.concatenate([self.index, index[keep]])
            self.xyz = np.concatenate([self.xyz, xyz[keep]])
            self.desc = np.concatenate([self.desc, desc[keep]])
        return len(keep)

    def radius_count(self, center, radius):
        """nanoflann radiusSearch(query, radius^2) with L2_Simple_Adaptor: squared distance accumulated x, y, z in double,
        accepted when strictly below the squared radius"""
        c = np.asarray(center, np.float64).reshape(3)
        d = self.xyz - c
        d2 = d[:, 0] * d[:, 0]
        d2 = d2 + d[:, 1] * d[:, 1]
        d2 = d2 + d[:, 2] * d[:, 2]
        return int(np.count_nonzero(d2 < float(radius) * float(radius)))

    def radius_search(self, center, radius):
        """-> row range [0, count): the reference's result (see the module docstring)"""
        return self.radius_count(center, radius)


def assign_landmark_indices(kp_desc, cloud: landmark_cloud, camera_center, match_radius, max_descriptor_distance):
    """-> int64 array, one entry per keypoint: the landmark index the keypoint takes, or -1 (keypoint_tracker.cpp:199-291)"""
    import oracle
    kp_desc = np.asarray(kp_desc, np.uint8).reshape(-1, 32)
    out = np.full(len(kp_desc), -1, np.int64)
    if len(kp_desc) == 0 or len(cloud) == 0:
        return out
    m = len(cloud)
    if match_radius > 0.0:
        m = cloud.radius_search(camera_center, match_radius)
        if m == 0:
            return out
    q, t, d = oracle.match_hamming_cross(kp_desc, cloud.desc[:m])          # BFMatcher(NORM_HAMMING, true).match(2d, 3d)
    ok = d.astype(np.float32).astype(np.float64) <= float(max_descriptor_distance)
    out[q[ok]] = cloud.index[t[ok]]
    return out


def _project(P, pts):
    """utils::project (utils/utils_opencv.cpp:443-482): x = P [X 1] in double; (x0 / x2, x1 / x2), or (0, 0) when |x2| <= 1e-9.
    -> (uv, w)"""
    P = np.asarray(P, np.float64).reshape(3, 4)
    h = pts @ P[:, :3].T + P[:, 3]
    w = h[:, 2]
    ok = np.abs(w) > 1e-9
    uv = np.zeros((len(pts), 2), np.float64)
    uv[ok] = h[ok, :2] / w[ok, None]
    return uv, w


def match_keypoints3d(cloud: landmark_cloud, kp_index, kp_xy, kp_desc, R, t, projection, radius, threshold,
                      image_size=None, frustum_margin=None):
    """utils::match_keypoints3d (matching_utils.cpp:132-216; with image_size / frustum_margin the overload at :218-343 with
    enable_frustum_culling): -> (landmark index, keypoint index, reprojection error) arrays in cv::BFMatcher's output order.
    kp_* are the keypoints in key order; R, t = pose_of_camera0_in_world."""
    import oracle
    empty = (np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.float32))
    kp_index = np.asarray(kp_index, np.int64)
    if len(cloud) == 0 or len(kp_index) == 0:
        return empty
    un = ~np.isin(kp_index, cloud.index)                                   # keypoints.values_unmatched(points3d_world)
    if not un.any():
        return empty
    R = np.asarray(R, np.float64).reshape(3, 3); t = np.asarray(t, np.float64).reshape(3)
    m = cloud.radius_search(t, radius)
    # pose.inv() * point: cv::Affine3d::inv() = (R^T, -R^T t); the product is R^T p + (-R^T t)
    Rt = R.T
    ti = -(Rt @ t)
    cam = cloud.xyz[:m] @ Rt.T + ti
    keep = cam[:, 2] > 0.0
    if image_size is not None:
        # is_in_frustum (matching_utils.cpp:105-130): in front of the camera, |w| >= 1e-9, projection inside the image + margin
        uv, w = _project(projection, cam)
        keep &= (np.abs(w) >= 1e-9) & (uv[:, 0] >= -frustum_margin) & (uv[:, 0] < image_size[0] + frustum_margin) & \
                (uv[:, 1] >= -frustum_margin) & (uv[:, 1] < image_size[1] + frustum_margin)
    rows = np.nonzero(keep)[0]
    if len(rows) == 0:
        return empty
    q, tr, _ = oracle.match_hamming_cross(cloud.desc[rows], np.asarray(kp_desc, np.uint8).reshape(-1, 32)[un])
    uv, _ = _project(projection, cam[rows])
    kxy = np.asarray(kp_xy, np.float32).reshape(-1, 2)[un].astype(np.float64)
    e = uv[q] - kxy[tr]
    err = np.sqrt(e[:, 0] * e[:, 0] + e[:, 1] * e[:, 1])
    ok = err < float(threshold)
    return cloud.index[rows][q[ok]], kp_index[un][tr[ok]], err[ok].astype(np.float32)


def match_temporal(keys_0, pts_0, desc_0, keys_1, pts_1, desc_1, camera_matrix, threshold, find_essential_mat):
    """utils::match_temporal (zenslam_core/source/matching/matching_utils.cpp:441-563) restated: unmatched keypoints of both
    maps in key order (:453-483), fewer than five on either side -> nothing (:485-488), BFMatcher(NORM_HAMMING, crossCheck)
    (:490-497), the caller's findEssentialMat (:512-519), then mask / epipolar error / distance <= 5 (:531-559).  The essential
    matrix is read the way the reference reads it: e.at<float> on the CV_64F result (:519-528).
    keys_* ascending int arrays, pts_* (n, 2) f32, desc_* (n, 32) u8 -> [(index_0, index_1, distance)]"""
    import oracle
    keys_0 = np.asarray(keys_0, np.int64); keys_1 = np.asarray(keys_1, np.int64)
    s0, s1 = set(keys_0.tolist()), set(keys_1.tolist())
    u0 = [i for i in np.argsort(keys_0, kind="stable") if int(keys_0[i]) not in s1]
    u1 = [i for i in np.argsort(keys_1, kind="stable") if int(keys_1[i]) not in s0]
    if len(u0) < 5 or len(u1) < 5:
        return []
    oq, ot, od = oracle.match_hamming_cross(np.asarray(desc_0, np.uint8)[u0], np.asarray(desc_1, np.uint8)[u1])
    if len(oq) == 0:
        return []
    p0 = np.asarray(pts_0, np.float32)[[u0[a] for a in oq]]; p1 = np.asarray(pts_1, np.float32)[[u1[b] for b in ot]]
    essential, mask = find_essential_mat(p0, p1)
    e64 = np.ascontiguousarray(essential, np.float64).reshape(3, 3)
    E = np.frombuffer(e64.tobytes(), np.float32).reshape(3, 6)[:, :3].astype(np.float64)
    k_inv = np.linalg.inv(np.asarray(camera_matrix, np.float64).reshape(3, 3))
    out = []
    for j, m in enumerate(np.asarray(mask).ravel()[:len(oq)]):
        if not m:
            continue
        a, b = np.append(p0[j].astype(np.float64), 1.0), np.append(p1[j].astype(np.float64), 1.0)
        err = float((((b @ k_inv.T) @ E) @ k_inv) @ a)
        if err > threshold or float(od[j]) > 5:
            continue
        out.append((int(keys_0[u0[oq[j]]]), int(keys_1[u1[ot[j]]]), float(od[j])))
    return out
