"""CPU oracle for the ZenSLAM stereo front-end hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product (``zenslam_b200`` and
``libzenslam_cuda.so``) never does; it fails loudly when the CUDA library is missing.

Two layers live here:

* ``zs_oracle.c`` -- a plain-C restatement of the OpenCV algorithms the reference calls
  (FAST-9-16 + NMS, grid bucketing, ORB blur + rBRIEF, BFMatcher Hamming / L2, optical-flow
  pyramid, pyramidal LK), every function citing the reference call site it stands for.
  Built on demand by :func:`build` into ``oracle/_build/libzs_oracle.so``.
* ``cv2_ref.py`` -- the reference's per-frame call pattern driven through the real
  dependency (Python ``cv2``): used to pin the C restatement (fixtures in ``tests/golden``)
  and as the timed CPU baseline.

Parity status: the reference's own tests hold no vectors for this path, so
the oracle is pinned against cv2 4.13.0 outputs -- committed fixtures in ``tests/golden/``
(``make_golden.py`` is the generating script) plus live checks where cv2 imports.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_SO = os.path.join(_BUILD, "libzs_oracle.so")
_SRC = os.path.join(_HERE, "zs_oracle.c")

LK_USE_INITIAL_FLOW = 4
LK_GET_MIN_EIGENVALS = 8


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc only; no OpenCV needed)."""
    os.makedirs(_BUILD, exist_ok=True)
    hdr = os.path.join(_HERE, "..", "include", "zs_orb_pattern.h")
    newest = max(os.path.getmtime(_SRC), os.path.getmtime(hdr))
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < newest:
        tmp = _SO + ".tmp.%d" % os.getpid()
        subprocess.check_call(
            ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-mfma", "-ffp-contract=off",
             "-fvisibility=hidden", _SRC, "-lm", "-o", tmp])
        os.replace(tmp, _SO)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.zso_pyramid_build.restype = C.c_void_p
        _lib.zso_pyramid_build.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        _lib.zso_pyramid_free.argtypes = [C.c_void_p]
        _lib.zso_pyramid_levels.argtypes = [C.c_void_p]
        _lib.zso_pyramid_level_size.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _lib.zso_pyramid_get_image.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib.zso_pyramid_get_deriv.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib.zso_lk_track.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                      C.c_double]
        _lib.zso_ratio_test.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p,
                                        C.c_void_p]
        _lib.zso_fb_check.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double,
                                      C.c_void_p]
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _u8img(img) -> np.ndarray:
    img = np.ascontiguousarray(img)
    assert img.dtype == np.uint8 and img.ndim == 2
    return img


# ---------------------------------------------------------------------------------------------
# pyramid
# ---------------------------------------------------------------------------------------------
def pyr_down(img):
    img = _u8img(img)
    h, w = img.shape
    out = np.empty(((h + 1) // 2, (w + 1) // 2), np.uint8)
    lib().zso_pyr_down(_p(img), w, h, w, _p(out), out.shape[1])
    return out


def scharr(img):
    """(h, w, 2) int16: (dx, dy)."""
    img = _u8img(img)
    h, w = img.shape
    out = np.empty((h, w, 2), np.int16)
    lib().zso_scharr(_p(img), w, h, w, _p(out), 2 * w)
    return out


def pyramid_num_levels(w, h, win, max_level):
    return lib().zso_pyramid_num_levels(w, h, win[0], win[1], max_level)


class Pyramid:
    """cv::buildOpticalFlowPyramid(img, pyr, win, max_level, withDerivatives=true) equivalent."""

    def __init__(self, img, win=(31, 31), max_level=3):
        img = _u8img(img)
        h, w = img.shape
        self.win = tuple(win)
        self._h = lib().zso_pyramid_build(_p(img), w, h, w, win[0], win[1], max_level)
        self.levels = lib().zso_pyramid_levels(self._h)

    def __del__(self):
        try:                                   # module globals may already be gone at interpreter shutdown
            if getattr(self, "_h", None):
                lib().zso_pyramid_free(self._h)
                self._h = None
        except Exception:
            pass

    def level_size(self, l):
        w, h = C.c_int(), C.c_int()
        lib().zso_pyramid_level_size(self._h, l, C.byref(w), C.byref(h))
        return w.value, h.value

    def image(self, l):
        w, h = self.level_size(l)
        out = np.empty((h, w), np.uint8)
        lib().zso_pyramid_get_image(self._h, l, _p(out))
        return out

    def deriv(self, l):
        w, h = self.level_size(l)
        out = np.empty((h, w, 2), np.int16)
        lib().zso_pyramid_get_deriv(self._h, l, _p(out))
        return out


# ---------------------------------------------------------------------------------------------
# detection / description
# ---------------------------------------------------------------------------------------------
def fast_detect(img, threshold):
    """cv::FAST(img, threshold, true, TYPE_9_16) -> (x int32[n], y int32[n], score int32[n]), raster order."""
    img = _u8img(img)
    h, w = img.shape
    cap = max(1, w * h)
    xs = np.empty(cap, np.int32); ys = np.empty(cap, np.int32); sc = np.empty(cap, np.int32)
    n = lib().zso_fast_detect(_p(img), w, h, w, int(threshold), _p(xs), _p(ys), _p(sc), cap)
    return xs[:n].copy(), ys[:n].copy(), sc[:n].copy()


def grid_detect(img, cell=(16, 16), threshold=10, occupied=None):
    """keypoint_detector_grid cells -> per-cell strongest FAST corner, cell row-major order."""
    img = _u8img(img)
    h, w = img.shape
    gw, gh = w // cell[0], h // cell[1]
    cap = max(1, gw * gh)
    xs = np.empty(cap, np.int32); ys = np.empty(cap, np.int32); sc = np.empty(cap, np.int32)
    occ = None
    if occupied is not None:
        occ = np.ascontiguousarray(occupied, np.uint8).reshape(gh, gw)
    n = lib().zso_grid_detect(_p(img), w, h, w, cell[0], cell[1], int(threshold),
                              _p(occ) if occ is not None else None, _p(xs), _p(ys), _p(sc))
    return xs[:n].copy(), ys[:n].copy(), sc[:n].copy()


def orb_blur(img):
    img = _u8img(img)
    h, w = img.shape
    out = np.empty_like(img)
    lib().zso_orb_blur(_p(img), w, h, w, _p(out), w)
    return out


def orb_filter(xs, ys, w, h):
    xs = np.ascontiguousarray(xs, np.float32); ys = np.ascontiguousarray(ys, np.float32)
    kept = np.empty(max(1, len(xs)), np.int32)
    n = lib().zso_orb_filter(_p(xs), _p(ys), len(xs), w, h, _p(kept))
    return kept[:n].copy()


def orb_compute(img, xs, ys, angles=None):
    """cv::ORB::create().compute(img, kps): returns (kept indices, descriptors n' x 32)."""
    img = _u8img(img)
    h, w = img.shape
    xs = np.ascontiguousarray(xs, np.float32); ys = np.ascontiguousarray(ys, np.float32)
    kept = orb_filter(xs, ys, w, h)
    blurred = orb_blur(img)
    kx = np.ascontiguousarray(xs[kept]); ky = np.ascontiguousarray(ys[kept])
    ka = None
    if angles is not None:
        ka = np.ascontiguousarray(np.asarray(angles, np.float32)[kept])
    desc = np.zeros((len(kept), 32), np.uint8)
    if len(kept):
        lib().zso_orb_describe(_p(blurred), w, h, w, _p(kx), _p(ky), _p(ka) if ka is not None else None,
                               len(kept), _p(desc))
    return kept, desc


def resize_linear_exact(img, dw, dh):
    """cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR_EXACT) for u8 single channel."""
    img = _u8img(img)
    h, w = img.shape
    out = np.empty((dh, dw), np.uint8)
    lib().zso_resize_linear_exact(_p(img), w, h, w, _p(out), dw, dh, dw)
    return out


def fast_atan2(y, x):
    f = lib().zso_fast_atan2_deg
    f.restype = C.c_float; f.argtypes = [C.c_float, C.c_float]
    return float(f(float(y), float(x)))


def orb_level_sizes(w, h, scale_factor=1.2, nlevels=8):
    out = []
    s = C.c_float(); lw = C.c_int(); lh = C.c_int()
    f = lib().zso_orb_level_size
    f.argtypes = [C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    for l in range(nlevels):
        f(w, h, scale_factor, l, C.byref(s), C.byref(lw), C.byref(lh))
        out.append((float(s.value), int(lw.value), int(lh.value)))
    return out


def orb_detect(img, mask=None, nfeatures=500, scale_factor=1.2, nlevels=8, edge=31, patch=31, fast_threshold=20,
               describe=True):
    """cv::ORB::create(nfeatures, scale_factor, nlevels, edge, 0, 2, HARRIS_SCORE, patch, fast_threshold)->detect(img,
    mask) [+ cv::ORB::create()->compute] in canonical order (octave, y, x).
    Returns dict(x, y, size, angle, response, octave[, desc])."""
    img = _u8img(img)
    h, w = img.shape
    m = None
    if mask is not None:
        m = np.ascontiguousarray(mask, np.uint8)
        assert m.shape == img.shape
    cap = 4 * nfeatures + 64
    f = lib().zso_orb_detect
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int,
                  C.c_int, C.c_int] + [C.c_void_p] * 7 + [C.c_int]
    while True:
        x = np.empty(cap, np.float32); y = np.empty(cap, np.float32); size = np.empty(cap, np.float32)
        ang = np.empty(cap, np.float32); resp = np.empty(cap, np.float32); octv = np.empty(cap, np.int32)
        desc = np.zeros((cap, 32), np.uint8) if describe else None
        n = f(_p(img), w, h, w, _p(m) if m is not None else None, w, nfeatures, scale_factor, nlevels, edge, patch,
              fast_threshold, _p(x), _p(y), _p(size), _p(ang), _p(resp), _p(octv), _p(desc) if describe else None, cap)
        if n <= cap:
            break
        cap = n
    out = dict(x=x[:n].copy(), y=y[:n].copy(), size=size[:n].copy(), angle=ang[:n].copy(), response=resp[:n].copy(),
               octave=octv[:n].copy())
    if describe:
        out["desc"] = desc[:n].copy()
    return out


# ---------------------------------------------------------------------------------------------
# matching
# ---------------------------------------------------------------------------------------------
def match_hamming_knn2(q, t):
    q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
    nq, nt = len(q), len(t)
    nb = q.shape[1] if nq else (t.shape[1] if nt else 32)
    idx = np.full((nq, 2), -1, np.int32); dist = np.zeros((nq, 2), np.int32)
    if nq:
        lib().zso_match_hamming_knn2(_p(q), nq, _p(t), nt, nb, _p(idx), _p(dist))
    return idx, dist


def match_l2_knn2(q, t):
    q = np.ascontiguousarray(q, np.float32); t = np.ascontiguousarray(t, np.float32)
    nq, nt = len(q), len(t)
    dim = q.shape[1] if nq else 128
    idx = np.full((nq, 2), -1, np.int32); dist = np.zeros((nq, 2), np.float32)
    if nq:
        lib().zso_match_l2_knn2(_p(q), nq, _p(t), nt, dim, _p(idx), _p(dist))
    return idx, dist


def ratio_test(idx, dist, ratio):
    idx = np.ascontiguousarray(idx, np.int32); dist = np.ascontiguousarray(dist, np.float32)
    n = len(idx)
    oq = np.empty(max(1, n), np.int32); ot = np.empty(max(1, n), np.int32); od = np.empty(max(1, n), np.float32)
    m = lib().zso_ratio_test(_p(idx), _p(dist), n, float(ratio), _p(oq), _p(ot), _p(od))
    return oq[:m].copy(), ot[:m].copy(), od[:m].copy()


def match_hamming_cross(q, t):
    q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
    nq, nt = len(q), len(t)
    oq = np.empty(max(1, nq), np.int32); ot = np.empty(max(1, nq), np.int32); od = np.empty(max(1, nq), np.int32)
    m = 0
    if nq and nt:
        m = lib().zso_match_hamming_cross(_p(q), nq, _p(t), nt, q.shape[1], _p(oq), _p(ot), _p(od))
    return oq[:m].copy(), ot[:m].copy(), od[:m].copy()


def match_l2_cross(q, t):
    q = np.ascontiguousarray(q, np.float32); t = np.ascontiguousarray(t, np.float32)
    nq, nt = len(q), len(t)
    oq = np.empty(max(1, nq), np.int32); ot = np.empty(max(1, nq), np.int32); od = np.empty(max(1, nq), np.float32)
    m = 0
    if nq and nt:
        m = lib().zso_match_l2_cross(_p(q), nq, _p(t), nt, q.shape[1], _p(oq), _p(ot), _p(od))
    return oq[:m].copy(), ot[:m].copy(), od[:m].copy()


# ---------------------------------------------------------------------------------------------
# KLT
# ---------------------------------------------------------------------------------------------
def lk_track(prev: Pyramid, nxt: Pyramid, prev_pts, next_pts=None, win=(31, 31), max_level=3,
             max_iters=99, eps=0.001, flags=LK_GET_MIN_EIGENVALS, min_eig=1e-4):
    """cv::calcOpticalFlowPyrLK -> (next_pts (n,2) f32, status u8, err f32)."""
    pp = np.ascontiguousarray(prev_pts, np.float32).reshape(-1, 2)
    n = len(pp)
    if next_pts is not None and (flags & LK_USE_INITIAL_FLOW):
        npts = np.ascontiguousarray(next_pts, np.float32).reshape(-1, 2).copy()
    else:
        npts = np.zeros((n, 2), np.float32)
    st = np.zeros(n, np.uint8); err = np.zeros(n, np.float32)
    if n:
        lib().zso_lk_track(prev._h, nxt._h, _p(pp), _p(npts), n, _p(st), _p(err), win[0], win[1], max_level,
                           max_iters, float(eps), flags, float(min_eig))
    return npts, st, err


def fb_check(p0, p0_back, status, status_back, klt_threshold):
    p0 = np.ascontiguousarray(p0, np.float32).reshape(-1, 2)
    pb = np.ascontiguousarray(p0_back, np.float32).reshape(-1, 2)
    st = np.ascontiguousarray(status, np.uint8); sb = np.ascontiguousarray(status_back, np.uint8)
    keep = np.zeros(len(p0), np.uint8)
    if len(p0):
        lib().zso_fb_check(_p(p0), _p(pb), _p(st), _p(sb), len(p0), float(klt_threshold), _p(keep))
    return keep.astype(bool)


@dataclass
class FrontendOptions:
    """The options.yaml keys the hot path reads, reference defaults."""
    cell_size: tuple = (16, 16)
    fast_threshold: int = 10
    klt_window_size: tuple = (31, 31)
    klt_max_level: int = 3
    klt_threshold: float = 1.0
    matcher_ratio: float = 0.8
    parallel_grid: bool = False          # detection.algorithm == PARALLEL_GRID (tumvi.yaml:43)


# ---------------------------------------------------------------------------------------------
# cornerSubPix (PARALLEL_GRID detector)
# ---------------------------------------------------------------------------------------------
def corner_subpix(img, xy, win=(5, 5), max_iters=30, eps=0.01):
    """cv::cornerSubPix(img, pts, win, (-1,-1), (EPS+COUNT, max_iters, eps)) as keypoint_detector_parallel calls it
    (keypoint_detector_parallel.cpp:160-170).  xy: (n, 2) float32 -> refined copy."""
    img = _u8img(img)
    h, w = img.shape
    out = np.ascontiguousarray(xy, np.float32).copy()
    L = lib()
    L.zso_corner_subpix.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_double]
    L.zso_corner_subpix(_p(img), w, h, w, _p(out), len(out), win[0], win[1], max_iters, float(eps))
    return out


# ---------------------------------------------------------------------------------------------
# pre-processing (processor::process): BGR->gray, CLAHE, remap
# ---------------------------------------------------------------------------------------------
def bgr2gray(bgr):
    bgr = np.ascontiguousarray(bgr, np.uint8)
    h, w, c = bgr.shape
    assert c == 3
    out = np.empty((h, w), np.uint8)
    L = lib()
    L.zso_bgr2gray.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    L.zso_bgr2gray(_p(bgr), w, h, 3 * w, _p(out), w)
    return out


def clahe(img, clip_limit=4.0, tiles=(8, 8)):
    img = _u8img(img)
    h, w = img.shape
    out = np.empty_like(img)
    L = lib()
    L.zso_clahe.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_int]
    L.zso_clahe(_p(img), w, h, w, float(clip_limit), tiles[0], tiles[1], _p(out), w)
    return out


def remap_linear(img, map_x, map_y):
    img = _u8img(img)
    h, w = img.shape
    map_x = np.ascontiguousarray(map_x, np.float32); map_y = np.ascontiguousarray(map_y, np.float32)
    dh, dw = map_x.shape
    out = np.empty((dh, dw), np.uint8)
    L = lib()
    L.zso_remap_linear.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                   C.c_void_p, C.c_int]
    L.zso_remap_linear(_p(img), w, h, w, _p(map_x), _p(map_y), dw, dw, dh, _p(out), dw)
    return out


# ---------------------------------------------------------------------------------------------
# stereo triangulation with gates (triangulator::triangulate_keypoints)
# ---------------------------------------------------------------------------------------------
def triangulate_keypoints(P0, P1, F, t, pts0, pts1, epipolar_threshold=0.01, reprojection_threshold=1.0, min_depth=1.0,
                          max_depth=50.0):
    """-> xyz (n,3) f64, keep (n,) bool, diag (n,4) f64 [epipolar error, reprojection error cam0, cam1, angle deg]"""
    P0 = np.ascontiguousarray(P0, np.float64); P1 = np.ascontiguousarray(P1, np.float64)
    t = np.ascontiguousarray(t, np.float64)
    pts0 = np.ascontiguousarray(pts0, np.float32); pts1 = np.ascontiguousarray(pts1, np.float32)
    n = len(pts0)
    xyz = np.zeros((n, 3), np.float64); keep = np.zeros(n, np.uint8); diag = np.zeros((n, 4), np.float64)
    Fp = None
    if F is not None:
        F = np.ascontiguousarray(F, np.float64); Fp = _p(F)
    L = lib()
    L.zso_triangulate_keypoints.argtypes = [C.c_void_p] * 6 + [C.c_int] + [C.c_double] * 4 + [C.c_void_p] * 3
    L.zso_triangulate_keypoints(_p(P0), _p(P1), Fp, _p(t), _p(pts0), _p(pts1), n, float(epipolar_threshold),
                                float(reprojection_threshold), float(min_depth), float(max_depth), _p(xyz), _p(keep), _p(diag))
    return xyz, keep.astype(bool), diag
