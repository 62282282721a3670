"""zenslam::matcher mirror (zenslam_core/include/zenslam/matching/matcher.h:14-38,
zenslam_core/source/matching/matcher.cpp:11-217, utils::create_matcher matching_utils.cpp:63-95).

BRUTE = 1-NN with cross-check, KNN = 2-NN + Lowe ratio; Hamming for binary descriptors, L2 otherwise.
The FLANN mode (approximate LSH / KD-tree) and the findFundamentalMat RANSAC gate (matcher.cpp:83-103)
are CPU algorithms outside the hot path: FLANN raises, the gate is an optional callable.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, lib
from .options import slam_options
from .runtime import Context
from .types import DMatch


class matcher:
    def __init__(self, opts: slam_options, is_binary: bool, ctx: Context, epipolar_gate=None):
        if opts.matcher == "FLANN":
            raise NotImplementedError("matcher: FLANN (approximate) mode is not part of the CUDA backend")
        self._options, self._is_binary, self._ctx, self._gate = opts, is_binary, ctx, epipolar_gate

    # -- descriptor stage ------------------------------------------------------------------------
    def _match_rows(self, d0: np.ndarray, d1: np.ndarray):
        nq, nt = len(d0), len(d1)
        if nq == 0 or nt == 0:
            return []
        if self._is_binary:
            q = np.ascontiguousarray(d0, np.uint8); t = np.ascontiguousarray(d1, np.uint8)
            dim, norm = 32, 0
        else:
            q = np.ascontiguousarray(d0, np.float32); t = np.ascontiguousarray(d1, np.float32)
            dim, norm = q.shape[1], 1
        mode = 0 if self._options.matcher == "KNN" else 1
        qi = np.empty(nq, np.int32); ti = np.empty(nq, np.int32); dist = np.empty(nq, np.float32)
        n = C.c_int(0)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        check(lib().zs_match_host(self._ctx._h, p(q), nq, p(t), nt, dim, norm, mode, float(self._options.matcher_ratio),
                                  p(qi), p(ti), p(dist), C.byref(n)))
        return [(int(qi[i]), int(ti[i]), float(dist[i])) for i in range(n.value)]

    def _finish(self, rows, un_l, un_r):
        if self._gate is not None and len(rows) >= 8:        # matcher.cpp:83-103 (caller-supplied, CPU)
            rows = self._gate(rows, un_l, un_r)
        # matcher.cpp:105-111: re-key to keypoint indices
        return [DMatch(un_l[q].index, un_r[t].index, d) for q, t, d in rows]

    # -- the two reference overloads -----------------------------------------------------------------
    def match_keypoints(self, keypoints_0, keypoints_1) -> list:
        if isinstance(keypoints_0, dict):
            # map overload (matcher.cpp:13-114): ascending index order, skip indices present in the other set
            un_l = [kp for idx, kp in sorted(keypoints_0.items()) if kp.index not in keypoints_1]
            un_r = [kp for idx, kp in sorted(keypoints_1.items()) if kp.index not in keypoints_0]
        else:
            # vector overload (matcher.cpp:116-217): all keypoints with non-empty descriptors
            if not keypoints_0 or not keypoints_1:
                return []
            un_l = [kp for kp in keypoints_0 if kp.descriptor is not None and len(kp.descriptor)]
            un_r = [kp for kp in keypoints_1 if kp.descriptor is not None and len(kp.descriptor)]
        if not un_l or not un_r:
            return []
        d0 = np.stack([kp.descriptor for kp in un_l]); d1 = np.stack([kp.descriptor for kp in un_r])
        return self._finish(self._match_rows(d0, d1), un_l, un_r)


def assign_landmark_indices(ctx: Context, keypoints: list, landmark_descriptors: np.ndarray, landmark_indices,
                            max_descriptor_distance: float) -> int:
    """keypoint_tracker::assign_landmark_indices, after the radius search has picked the candidate landmarks
    (zenslam_core/source/tracking/keypoint_tracker.cpp:199-291): cross-checked Hamming 1-NN of the keypoints that carry
    a descriptor against the landmark descriptors; a keypoint whose match has distance <= max_descriptor_distance takes
    the landmark's index.  Mutates `keypoints` like the reference; returns the number of re-indexed keypoints."""
    if not keypoints or landmark_descriptors is None or len(landmark_descriptors) == 0:
        return 0
    rows = [i for i, kp in enumerate(keypoints) if kp.descriptor is not None and len(kp.descriptor)]
    if not rows:
        return 0
    q = np.ascontiguousarray(np.stack([keypoints[i].descriptor for i in rows]), np.uint8)
    t = np.ascontiguousarray(landmark_descriptors, np.uint8)
    lm = np.empty(len(rows), np.int32); dist = np.empty(len(rows), np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    check(lib().zs_assign_landmarks_host(ctx._h, p(q), len(rows), p(t), len(t), float(max_descriptor_distance), p(lm), p(dist)))
    n = 0
    for j, i in enumerate(rows):
        if lm[j] >= 0:
            keypoints[i].index = int(landmark_indices[int(lm[j])])
            n += 1
    return n


def match_keypoints3d(ctx: Context, landmark_index, landmark_xyz, landmark_desc, keypoints, R, t, projection, radius: float,
                      threshold: float, image_size=None, frustum_margin: float = 50.0) -> list:
    """utils::match_keypoints3d (zenslam_core/source/matching/matching_utils.cpp:132-216; with image_size the overload with
    frustum culling, :218-343).  Landmarks in the cloud's insertion order (index (M,), xyz (M, 3) world, desc (M, 32));
    `keypoints`: a map / list of keypoints -- the ones whose index is a landmark index are dropped here, like
    keypoints.values_unmatched(points3d_world), the rest go in key order.  R, t = pose_of_camera0_in_world; projection 3x4.
    -> [DMatch(queryIdx = landmark index, trainIdx = keypoint index, distance = reprojection error)]"""
    lm_i = np.ascontiguousarray(landmark_index, np.int32); lm_x = np.ascontiguousarray(landmark_xyz, np.float64).reshape(-1, 3)
    lm_d = np.ascontiguousarray(landmark_desc, np.uint8).reshape(-1, 32)
    kps = [keypoints[k] for k in sorted(keypoints)] if isinstance(keypoints, dict) else sorted(keypoints, key=lambda k: k.index)
    have = set(lm_i.tolist())
    kps = [k for k in kps if k.index not in have]
    if len(lm_i) == 0 or not kps:
        return []
    k_i = np.array([k.index for k in kps], np.int32); k_xy = np.array([k.pt for k in kps], np.float32).reshape(-1, 2)
    k_d = np.ascontiguousarray(np.stack([k.descriptor for k in kps]), np.uint8)
    Rm = np.ascontiguousarray(R, np.float64).reshape(3, 3); tv = np.ascontiguousarray(t, np.float64).reshape(3)
    P = np.ascontiguousarray(projection, np.float64).reshape(3, 4)
    cap = min(len(lm_i), len(kps))
    o_l = np.empty(cap, np.int32); o_k = np.empty(cap, np.int32); o_e = np.empty(cap, np.float32)
    n = C.c_int(0)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    w, h = (int(image_size[0]), int(image_size[1])) if image_size is not None else (0, 0)
    check(lib().zs_match_keypoints3d_host(ctx._h, p(lm_i), p(lm_x), p(lm_d), len(lm_i), p(k_i), p(k_xy), p(k_d), len(kps), p(Rm), p(tv), p(P),
                                          float(radius), float(threshold), w, h, float(frustum_margin), p(o_l), p(o_k), p(o_e), C.byref(n)))
    return [DMatch(int(o_l[i]), int(o_k[i]), float(o_e[i])) for i in range(n.value)]


def _essential_as_the_reference_reads_it(essential) -> np.ndarray:
    """utils::match_temporal reads the CV_64F matrix cv::findEssentialMat returns with e.at<float>(i, j)
    (matching_utils.cpp:519-528): element (i, j) is the 32-bit float found at byte offset 24 i + 4 j of the double buffer.
    Reproduced as written -- a drop-in must gate on the same numbers."""
    e64 = np.ascontiguousarray(essential, np.float64).reshape(3, 3)
    return np.frombuffer(e64.tobytes(), np.float32).reshape(3, 6)[:, :3].astype(np.float64)


def match_temporal(ctx: Context, keypoints_map_0: dict, keypoints_map_1: dict, camera_matrix, threshold: float,
                   find_essential_mat) -> list:
    """utils::match_temporal (zenslam_core/source/matching/matching_utils.cpp:441-563): the keypoints of either
    map whose index the other map lacks, in key order; nothing when either side has fewer than five; cross-checked 1-NN
    (cv::BFMatcher(NORM_HAMMING, true).match) on the device; then the caller's cv::findEssentialMat(points_0, points_1,
    camera_matrix, RANSAC, 0.99, threshold, mask) -- a randomised CPU algorithm, out of scope like the other RANSAC gates --
    passed in as `find_essential_mat(points_0, points_1) -> (E 3x3 float64, mask)`; a match survives when its mask is set,
    pt_1^T K^-T E K^-1 pt_0 <= threshold and its descriptor distance <= 5.
    -> [DMatch(keypoint index in map 0, keypoint index in map 1, descriptor distance)]"""
    un_0 = [kp for _, kp in sorted(keypoints_map_0.items()) if kp.index not in keypoints_map_1]
    un_1 = [kp for _, kp in sorted(keypoints_map_1.items()) if kp.index not in keypoints_map_0]
    if len(un_0) < 5 or len(un_1) < 5:
        return []
    q = np.ascontiguousarray(np.stack([kp.descriptor for kp in un_0]), np.uint8)
    t = np.ascontiguousarray(np.stack([kp.descriptor for kp in un_1]), np.uint8)
    idx = np.empty(len(q), np.int32); dist = np.empty(len(q), np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    check(lib().zs_knn_match_host(ctx._h, p(q), len(q), p(t), len(t), 32, 0, 1, 1, p(idx), p(dist)))
    rows = [(i, int(idx[i]), float(dist[i])) for i in range(len(q)) if idx[i] >= 0]
    if not rows:
        return []
    points_0 = np.array([un_0[a].pt for a, _, _ in rows], np.float32)
    points_1 = np.array([un_1[b].pt for _, b, _ in rows], np.float32)
    essential, mask = find_essential_mat(points_0, points_1)
    E = _essential_as_the_reference_reads_it(essential)
    k_inv = np.linalg.inv(np.asarray(camera_matrix, np.float64).reshape(3, 3))
    out = []
    for (a, b, d), m in zip(rows, np.asarray(mask).ravel()):          # std::views::zip stops at the shorter range
        if not m:
            continue
        pt_0 = np.array([np.float64(np.float32(un_0[a].pt[0])), np.float64(np.float32(un_0[a].pt[1])), 1.0])
        pt_1 = np.array([np.float64(np.float32(un_1[b].pt[0])), np.float64(np.float32(un_1[b].pt[1])), 1.0])
        error = float((((pt_1 @ k_inv.T) @ E) @ k_inv) @ pt_0)
        if error > threshold or d > 5:
            continue
        out.append(DMatch(un_0[a].index, un_1[b].index, d))
    return out


def _epilines(points, which_image: int, fundamental) -> np.ndarray:
    """cv::computeCorrespondEpilines for CV_32F points (calib3d fundam.cpp): l = F p (which_image 1) or F^T p (2), accumulated
    in double, scaled by 1 / sqrt(a^2 + b^2) (1 when that is 0), stored as float32"""
    F = np.asarray(fundamental, np.float64).reshape(3, 3)
    if which_image == 2:
        F = F.T
    p = np.asarray(points, np.float32).reshape(-1, 2).astype(np.float64)
    a = F[0, 0] * p[:, 0] + F[0, 1] * p[:, 1] + F[0, 2]
    b = F[1, 0] * p[:, 0] + F[1, 1] * p[:, 1] + F[1, 2]
    c = F[2, 0] * p[:, 0] + F[2, 1] * p[:, 1] + F[2, 2]
    nu = a * a + b * b
    nu = np.where(nu != 0, 1.0 / np.sqrt(np.where(nu != 0, nu, 1.0)), 1.0)
    return np.stack([a * nu, b * nu, c * nu], 1).astype(np.float32)


def match_keylines(ctx: Context, keylines_map_0: dict, keylines_map_1: dict, fundamental, epipolar_threshold: float) -> list:
    """utils::match_keylines (zenslam_core/source/matching/matching_utils.cpp:345-439): every keyline of either
    map in key order, cross-checked Hamming 1-NN of their 32-byte LBD descriptors on the device
    (cv::BFMatcher(NORM_HAMMING, true).match), then the epipolar gate on both endpoints and the midpoint: point-to-epiline
    distance in BOTH images <= epipolar_threshold, in float like the reference's expression.  Keyline detection and
    description (LSD / LBD) stay on the host (out of scope).  -> [DMatch(keyline index 0, keyline index 1, distance)]"""
    kl0 = [kl for _, kl in sorted(keylines_map_0.items())]
    kl1 = [kl for _, kl in sorted(keylines_map_1.items())]
    if not kl0 or not kl1:
        return []
    q = np.ascontiguousarray(np.stack([k.descriptor for k in kl0]), np.uint8).reshape(len(kl0), -1)
    t = np.ascontiguousarray(np.stack([k.descriptor for k in kl1]), np.uint8).reshape(len(kl1), -1)
    idx = np.empty(len(q), np.int32); dist = np.empty(len(q), np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    check(lib().zs_knn_match_host(ctx._h, p(q), len(q), p(t), len(t), 32, 0, 1, 1, p(idx), p(dist)))
    f32 = np.float32
    out = []
    for i in range(len(q)):
        if idx[i] < 0:
            continue
        a, b = kl0[i], kl1[int(idx[i])]
        pts0 = np.array([(a.startPointX, a.startPointY), (a.endPointX, a.endPointY), a.pt], f32)
        pts1 = np.array([(b.startPointX, b.startPointY), (b.endPointX, b.endPointY), b.pt], f32)
        lines1 = _epilines(pts0, 1, fundamental)             # epilines of image-0 points, in image 1
        lines0 = _epilines(pts1, 2, fundamental)             # epilines of image-1 points, in image 0
        good = True
        for k in range(3):
            e0 = abs(f32(f32(f32(lines0[k, 0] * pts0[k, 0]) + f32(lines0[k, 1] * pts0[k, 1])) + lines0[k, 2])) / \
                np.sqrt(f32(f32(lines0[k, 0] * lines0[k, 0]) + f32(lines0[k, 1] * lines0[k, 1])))
            e1 = abs(f32(f32(f32(lines1[k, 0] * pts1[k, 0]) + f32(lines1[k, 1] * pts1[k, 1])) + lines1[k, 2])) / \
                np.sqrt(f32(f32(lines1[k, 0] * lines1[k, 0]) + f32(lines1[k, 1] * lines1[k, 1])))
            if float(e0) > epipolar_threshold or float(e1) > epipolar_threshold:
                good = False
                break
        if good:
            out.append(DMatch(a.index, b.index, float(dist[i])))
    return out


def assign_keyline_landmark_indices(ctx: Context, keylines: list, landmark_descriptors, landmark_indices,
                                    max_descriptor_distance: float) -> int:
    """keyline_tracker's landmark association (zenslam_core/source/tracking/keyline_tracker.cpp:135-163): 1-NN WITHOUT cross
    check (cv::BFMatcher(NORM_HAMMING, false).match) of the keylines that carry a descriptor against the landmark
    descriptors; distance <= max_descriptor_distance renames the keyline to the landmark's index.  Mutates `keylines`;
    returns the number of renamed keylines."""
    rows = [i for i, kl in enumerate(keylines) if kl.descriptor is not None and len(kl.descriptor)]
    if not rows or landmark_descriptors is None or len(landmark_descriptors) == 0:
        return 0
    q = np.ascontiguousarray(np.stack([keylines[i].descriptor for i in rows]), np.uint8).reshape(len(rows), -1)
    t = np.ascontiguousarray(landmark_descriptors, np.uint8).reshape(-1, 32)
    idx = np.empty(len(q), np.int32); dist = np.empty(len(q), np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    check(lib().zs_knn_match_host(ctx._h, p(q), len(q), p(t), len(t), 32, 0, 1, 0, p(idx), p(dist)))
    n = 0
    for j, i in enumerate(rows):
        if idx[j] >= 0 and float(dist[j]) <= max_descriptor_distance:
            keylines[i].index = int(landmark_indices[int(idx[j])])
            n += 1
    return n
