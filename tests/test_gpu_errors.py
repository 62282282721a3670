"""GPU: error behaviour of the C ABI -- bad arguments come back as status codes with a message (never a crash, never a
silent fallback), empty inputs are empty outputs like the reference (matcher.cpp:55-56, keypoint_tracker.cpp:122)."""
import ctypes as C

import numpy as np
import pytest

from zenslam_b200 import _lib

pytestmark = pytest.mark.gpu

ZS_ERR_INVALID, ZS_ERR_UNSUPPORTED = -2, -5


@pytest.fixture(scope="module")
def ctx():
    from zenslam_b200.runtime import Context
    c = Context()
    yield c
    c.close()


def test_bad_arguments_return_status(ctx):
    L = _lib.lib()
    p = C.c_void_p()
    assert L.zs_pyramid_create(ctx._h, 0, 480, 1, 31, 31, 3, C.byref(p)) == ZS_ERR_INVALID
    assert L.zs_pyramid_create(ctx._h, 752, 480, 1, 200, 31, 3, C.byref(p)) == ZS_ERR_INVALID        # window > 127
    assert b"window" in L.zs_last_error_string()
    assert L.zs_pyramid_create(ctx._h, 64, 64, 2, 15, 15, 2, C.byref(p)) == 0
    # LK window must match the pyramid's; GET_MIN_EIGENVALS is the only mode
    prm = _lib.LkParams(31, 31, 3, 30, 0.01, _lib.LK_GET_MIN_EIGENVALS, 1e-4)
    z = C.c_void_p(0)
    one = (C.c_int * 1)(0)
    assert L.zs_klt_track(ctx._h, p, one, one, one, one, one, 1, 4, C.byref(prm), one, one) == ZS_ERR_INVALID
    prm2 = _lib.LkParams(15, 15, 2, 30, 0.01, 0, 1e-4)
    assert L.zs_klt_track(ctx._h, p, one, one, one, one, one, 1, 4, C.byref(prm2), one, one) == ZS_ERR_INVALID
    assert L.zs_klt_track(ctx._h, p, z, z, z, z, z, 1, 4, C.byref(prm2), z, z) == ZS_ERR_INVALID      # null pointers
    assert L.zs_fast_grid_detect(ctx._h, p, 0, 1, 3, 3, 10, None, one, one, one, 1000) == ZS_ERR_INVALID   # cell < 7
    assert L.zs_corner_subpix(ctx._h, p, 0, 1, one, one, 4, 9, 9, 30, 0.01) == ZS_ERR_INVALID          # half window > 7
    L.zs_pyramid_destroy(p)
    fe = C.c_void_p()
    o = _lib.FrontendOptions(752, 480, 0, 16, 16, 10, 31, 31, 3, 1.0, 0.8, 99, 0.001, 1e-4)
    assert L.zs_frontend_create(ctx._h, C.byref(o), C.byref(fe)) == ZS_ERR_INVALID                     # batch 0
    assert L.zs_status_string(ZS_ERR_INVALID) and L.zs_status_string(0)
    # GRID with cells that can reach the reference's ORB::detect fallback for empty cells (keypoint_detector_grid.cpp:92-95)
    # is refused, not approximated; PARALLEL_GRID (no fallback in the reference) and provably safe sizes are accepted
    for cell, thr, parallel, want in ((80, 10, 0, ZS_ERR_UNSUPPORTED), (64, 25, 0, ZS_ERR_UNSUPPORTED), (64, 10, 0, 0), (80, 10, 1, 0)):
        o = _lib.FrontendOptions(752, 480, 1, cell, cell, thr, 31, 31, 3, 1.0, 0.8, 99, 0.001, 1e-4, parallel)
        assert L.zs_frontend_create(ctx._h, C.byref(o), C.byref(fe)) == want, (cell, thr, parallel)
        if want == 0:
            L.zs_frontend_destroy(fe)
        else:
            assert b"fallback" in L.zs_last_error_string()


def test_empty_inputs_are_empty_outputs(ctx):
    from zenslam_b200 import slam_options
    from zenslam_b200.matching import assign_landmark_indices, matcher
    from zenslam_b200.tracking import create_cuda_pyr_lk
    m = matcher(slam_options(matcher="KNN"), True, ctx)
    assert m.match_keypoints([], []) == [] and m.match_keypoints({}, {}) == []
    assert assign_landmark_indices(ctx, [], np.zeros((0, 32), np.uint8), [], 32.0) == 0
    lk = create_cuda_pyr_lk(ctx)
    img = np.zeros((64, 64), np.uint8)
    pts, st, err = lk.calc_optical_flow_pyr_lk([img], [img], np.zeros((0, 2), np.float32), None, (15, 15), 2)
    assert pts.shape == (0, 2) and len(st) == 0 and len(err) == 0
    # a textureless image tracks nothing but does not fail: status 0, the reference's min-eigenvalue rule
    q = np.array([[32.0, 32.0], [10.5, 50.25]], np.float32)
    pts, st, err = lk.calc_optical_flow_pyr_lk([img], [img], q, None, (15, 15), 2)
    assert list(st) == [0, 0]


def test_two_contexts_and_frontends_coexist(ctx):
    """independent contexts (own streams) give the same answers; nothing is process-global except the library"""
    from zenslam_b200 import slam_options, synthetic as syn
    from zenslam_b200.frontend import StereoFrontend
    from zenslam_b200.runtime import Context
    seq, _ = syn.stereo_sequence(320, 240, 2, 5)
    L, R = np.ascontiguousarray(seq[:, 0]), np.ascontiguousarray(seq[:, 1])
    c2 = Context(stream=None)
    a = StereoFrontend(ctx, 320, 240, 2, slam_options()); b = StereoFrontend(c2, 320, 240, 2, slam_options())
    a.submit(L, R); b.submit(L, R)
    ra, rb = a.wait(), b.wait()
    for k in ra:
        assert np.array_equal(ra[k], rb[k]), k
    a.close(); b.close(); c2.close()
