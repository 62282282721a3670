"""ctypes binding of libzenslam_cuda.so (the C ABI declared in include/zenslam_cuda.h).

The library is the product: there is no Python or CPU fallback.  Loading fails loudly when the
shared object has not been built (``python -m zenslam_b200.build``), and creating a context fails
loudly (``ZenslamCudaError``) when no sm_100 device is present.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libzenslam_cuda.so")

ZS_OK = 0
LK_USE_INITIAL_FLOW = 4
LK_GET_MIN_EIGENVALS = 8


class ZenslamCudaError(RuntimeError):
    pass


class LkParams(C.Structure):
    _fields_ = [("win_w", C.c_int), ("win_h", C.c_int), ("max_level", C.c_int), ("max_iters", C.c_int),
                ("epsilon", C.c_double), ("flags", C.c_int), ("min_eig_threshold", C.c_double)]


class TriangulationParams(C.Structure):
    _fields_ = [("filter_epipolar", C.c_int), ("epipolar_threshold", C.c_double), ("reprojection_threshold", C.c_double),
                ("min_depth", C.c_double), ("max_depth", C.c_double)]


class TrackerOptions(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("cell_w", C.c_int), ("cell_h", C.c_int), ("fast_threshold", C.c_int),
                ("klt_win_w", C.c_int), ("klt_win_h", C.c_int), ("klt_max_level", C.c_int), ("klt_threshold", C.c_double),
                ("capacity", C.c_int), ("first_index", C.c_int), ("sequences", C.c_int), ("parallel_grid", C.c_int),
                ("landmark_capacity", C.c_int), ("landmark_match_radius", C.c_double), ("landmark_match_distance", C.c_double)]


class TrackerResults(C.Structure):
    _fields_ = [("cap", C.c_int), ("n", C.c_void_p), ("index", C.c_void_p * 2), ("xy", C.c_void_p * 2),
                ("response", C.c_void_p * 2), ("desc", C.c_void_p * 2), ("next_index", C.c_void_p)]


class FrontendOptions(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("batch", C.c_int),
                ("cell_w", C.c_int), ("cell_h", C.c_int), ("fast_threshold", C.c_int),
                ("klt_win_w", C.c_int), ("klt_win_h", C.c_int), ("klt_max_level", C.c_int),
                ("klt_threshold", C.c_double), ("matcher_ratio", C.c_double),
                ("max_iters", C.c_int), ("epsilon", C.c_double), ("min_eig_threshold", C.c_double),
                ("parallel_grid", C.c_int)]


class FrontendResults(C.Structure):
    _fields_ = [("cap", C.c_int),
                ("n_left", C.c_void_p), ("n_right", C.c_void_p),
                ("kp_left", C.c_void_p), ("kp_right", C.c_void_p),
                ("resp_left", C.c_void_p), ("resp_right", C.c_void_p),
                ("desc_left", C.c_void_p), ("desc_right", C.c_void_p),
                ("match_idx", C.c_void_p), ("match_dist", C.c_void_p), ("match_pass", C.c_void_p),
                ("track_pts", C.c_void_p), ("track_keep", C.c_void_p), ("track_n", C.c_void_p)]


P = C.c_void_p
I = C.c_int
Z = C.c_size_t
D = C.c_double

# name -> (restype, argtypes); every symbol include/zenslam_cuda.h declares
SIGNATURES = {
    "zs_is_available": (I, []),
    "zs_version": (C.c_char_p, []),
    "zs_status_string": (C.c_char_p, [I]),
    "zs_last_error_string": (C.c_char_p, []),
    "zs_context_create": (I, [I, P, C.POINTER(P)]),
    "zs_context_destroy": (None, [P]),
    "zs_context_synchronize": (I, [P]),
    "zs_context_async_error": (I, [P]),
    "zs_context_stream": (P, [P]),
    "zs_context_launch_count": (C.c_uint64, [P]),
    "zs_context_reload_switches": (I, [P]),
    "zs_cvt_bgr2gray": (I, [P, P, Z, Z, I, I, I, P, Z, Z]),
    "zs_clahe": (I, [P, P, Z, Z, I, I, I, D, I, I, P, Z, Z]),
    "zs_remap_linear": (I, [P, P, Z, Z, I, I, I, P, P, Z, Z, I, I, P, Z, Z]),
    "zs_process_image_host": (I, [P, P, I, I, I, Z, I, D, P, P, P]),
    "zs_pyramid_level0": (I, [P, I, C.POINTER(P), C.POINTER(Z), C.POINTER(Z)]),
    "zs_pyramid_create": (I, [P, I, I, I, I, I, I, C.POINTER(P)]),
    "zs_pyramid_destroy": (None, [P]),
    "zs_pyramid_levels": (I, [P]),
    "zs_pyramid_level_size": (I, [P, I, C.POINTER(I), C.POINTER(I)]),
    "zs_pyramid_upload": (I, [P, P, P, Z, Z, I, I, I]),
    "zs_pyramid_build": (I, [P, P, I, I]),
    "zs_pyramid_download_image": (I, [P, P, I, I, P]),
    "zs_pyramid_download_deriv": (I, [P, P, I, I, P]),
    "zs_fast_grid_detect": (I, [P, P, I, I, I, I, I, P, P, P, P, I]),
    "zs_fast_detect": (I, [P, P, I, I, I, P, P, P, P, I]),
    "zs_corner_subpix": (I, [P, P, I, I, P, P, I, I, I, I, D]),
    "zs_orb_compute": (I, [P, P, I, I, P, P, P, P, I, P, P, P, P, P]),
    "zs_orb_download_blur": (I, [P, P, I, P]),
    "zs_orb_detector_create": (I, [P, I, I, I, I, C.c_float, I, I, I, I, C.POINTER(P)]),
    "zs_orb_detector_destroy": (None, [P]),
    "zs_orb_detector_capacity": (I, [P]),
    "zs_orb_detector_level": (I, [P, I, C.POINTER(I), C.POINTER(I), C.POINTER(C.c_float), C.POINTER(I)]),
    "zs_orb_detect_and_compute": (I, [P, P, P, Z, Z, P, Z, Z, I, P, P, P, P, P, P, P]),
    "zs_orb_detector_download_level": (I, [P, P, I, I, I, P]),
    "zs_detect_keypoints_orb_host": (I, [P, P, I, I, Z, P, Z, I, C.c_float, I, I, I, I, P, P, P, P, P, P, P, I,
                                         C.POINTER(I)]),
    "zs_match_hamming_knn2": (I, [P, P, P, Z, P, P, Z, I, I, I, D, P, P, P]),
    "zs_match_hamming_cross": (I, [P, P, P, Z, P, P, Z, I, I, I, P, P]),
    "zs_match_l2_knn2": (I, [P, P, P, Z, P, P, Z, I, I, I, I, D, P, P, P]),
    "zs_match_l2_cross": (I, [P, P, P, Z, P, P, Z, I, I, I, I, P, P]),
    "zs_match_l2_knn2_u8": (I, [P, P, P, P, P, I, I, I, I, D, P, P, P]),
    "zs_match_l2_cross_u8": (I, [P, P, P, P, P, I, I, I, I, P, P]),
    "zs_klt_track": (I, [P, P, P, P, P, P, P, I, I, C.POINTER(LkParams), P, P]),
    "zs_klt_track_fb": (I, [P, P, P, P, P, P, P, I, I, C.POINTER(LkParams), D, P, P, P]),
    "zs_calc_optical_flow_pyr_lk_host": (I, [P, P, P, I, I, Z, P, P, I, P, P, C.POINTER(LkParams)]),
    "zs_lk_cache_stats": (I, [P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "zs_track_keypoints_host": (I, [P, P, P, I, I, Z, P, P, I, C.POINTER(LkParams), D, P, P, P, P]),
    "zs_detect_keypoints_grid_host": (I, [P, P, I, I, Z, I, I, I, P, P, P, P, P, C.POINTER(I)]),
    "zs_detect_keypoints_parallel_host": (I, [P, P, I, I, Z, I, I, I, P, P, P, P, P, C.POINTER(I)]),
    "zs_detect_keypoints_simple_host": (I, [P, P, I, I, Z, P, Z, I, P, P, P, P, I, C.POINTER(I)]),
    "zs_match_host": (I, [P, P, I, P, I, I, I, I, D, P, P, P, C.POINTER(I)]),
    "zs_knn_match_host": (I, [P, P, I, P, I, I, I, I, I, P, P]),
    "zs_assign_landmarks_host": (I, [P, P, I, P, I, D, P, P]),
    "zs_triangulate_keypoints": (I, [P, P, P, P, P, P, P, I, C.POINTER(TriangulationParams), P, P, P]),
    "zs_triangulate_keypoints_host": (I, [P, P, P, P, P, P, P, I, C.POINTER(TriangulationParams), P, P, P]),
    "zs_tracker_create": (I, [P, C.POINTER(TrackerOptions), C.POINTER(P)]),
    "zs_tracker_destroy": (None, [P]),
    "zs_tracker_capacity": (I, [P]),
    "zs_tracker_sequences": (I, [P]),
    "zs_tracker_set_predictions": (I, [P, I, I, P, P, I]),
    "zs_tracker_landmarks_add_host": (I, [P, I, P, P, P, I, C.POINTER(I)]),
    "zs_tracker_landmarks_size": (I, [P, I]),
    "zs_tracker_set_camera_center": (I, [P, I, P]),
    "zs_match_keypoints3d_host": (I, [P, P, P, P, I, P, P, P, I, P, P, P, D, D, I, I, D, P, P, P, C.POINTER(I)]),
    "zs_tracker_track_host": (I, [P, P, P, Z, Z, C.POINTER(TrackerResults)]),
    "zs_tracker_track": (I, [P, P, P, Z, Z]),
    "zs_tracker_download": (I, [P, C.POINTER(TrackerResults)]),
    "zs_tracker_filter_epipolar": (I, [P, I, P, D]),
    "zs_tracker_submit_host": (I, [P, P, P, Z, Z, C.POINTER(TrackerResults)]),
    "zs_tracker_wait": (I, [P]),
    "zs_tracker_in_flight": (I, [P]),
    "zs_frontend_create": (I, [P, C.POINTER(FrontendOptions), C.POINTER(P)]),
    "zs_frontend_destroy": (None, [P]),
    "zs_frontend_capacity": (I, [P]),
    "zs_frontend_h2d_bytes": (Z, [P]),
    "zs_frontend_d2h_bytes": (Z, [P]),
    "zs_frontend_upload": (I, [P, P, P, Z, Z, I]),
    "zs_frontend_run": (I, [P]),
    "zs_frontend_timing_enable": (I, [P, I]),
    "zs_frontend_timing_collect": (I, [P, C.POINTER(C.c_float), C.POINTER(I)]),
    "zs_frontend_download": (I, [P, C.POINTER(FrontendResults)]),
    "zs_frontend_process_host": (I, [P, P, P, Z, Z, C.POINTER(FrontendResults)]),
    "zs_frontend_submit_host": (I, [P, P, P, Z, Z, C.POINTER(FrontendResults)]),
    "zs_frontend_wait": (I, [P]),
    "zs_frontend_set_preprocess": (I, [P, I, I, D, P, P, P, P]),
    "zs_frontend_in_flight": (I, [P]),
}

_lib = None


def lib():
    """Load libzenslam_cuda.so; raise if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ZenslamCudaError(
                "libzenslam_cuda.so is missing (%s). Build it with `python -m zenslam_b200.build`; "
                "this package has no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)        # AttributeError here means the header and the library diverged
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status: int):
    if status != ZS_OK:
        L = lib()
        raise ZenslamCudaError("%s: %s" % (L.zs_status_string(status).decode(), L.zs_last_error_string().decode()))
