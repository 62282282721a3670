"""Thin Python layer over the C ABI: context, pyramids and the batched device-pointer calls.

PyTorch is used for plumbing only -- device buffers (``torch.empty(..., device='cuda')``) and the
current CUDA stream; every computation happens inside libzenslam_cuda.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from ._lib import (LK_GET_MIN_EIGENVALS, LK_USE_INITIAL_FLOW, LkParams, ZenslamCudaError, check, lib)


def is_available() -> bool:
    """cf. zenslam::metal::is_available (zenslam_metal/include/zenslam_metal/pyr_lk.h:9)."""
    return bool(lib().zs_is_available())


def _torch():
    import torch
    return torch


def _ptr(t):
    """device pointer of a torch tensor, or NULL"""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous()
    return C.c_void_p(t.data_ptr())


@dataclass
class LK:
    """cv::calcOpticalFlowPyrLK arguments as the reference passes them (keypoint_tracker.cpp:142-170)."""
    win: tuple = (31, 31)
    max_level: int = 3
    max_iters: int = 99
    epsilon: float = 0.001
    flags: int = LK_GET_MIN_EIGENVALS
    min_eig_threshold: float = 1e-4

    def c(self) -> LkParams:
        return LkParams(self.win[0], self.win[1], self.max_level, self.max_iters, self.epsilon, self.flags,
                        self.min_eig_threshold)


class Context:
    """zs_context: one per GPU.  By default work is enqueued on torch's current stream so that
    torch.cuda.Event timing and torch tensors see it in order."""

    def __init__(self, device: int | None = None, stream="torch"):
        L = lib()
        torch = _torch()
        if not torch.cuda.is_available():
            raise ZenslamCudaError("no CUDA device: zenslam_b200 has no CPU fallback")
        if device is None:
            device = torch.cuda.current_device()
        self.device = device
        torch.cuda.set_device(device)
        sp = None
        if stream == "torch":
            sp = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
            if not sp.value:
                # the legacy default stream has handle 0, which the C API reads as "create one"; use a real one
                self._tstream = torch.cuda.Stream(device)
                torch.cuda.set_stream(self._tstream)
                sp = C.c_void_p(self._tstream.cuda_stream)
        h = C.c_void_p()
        check(L.zs_context_create(device, sp, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib().zs_context_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        check(lib().zs_context_synchronize(self._h))

    def async_error(self):
        """waits for the stream and raises if a kernel of a stream-asynchronous entry rejected its DATA (zs_match_l2_* on
        descriptors that are not integers in 0..255)"""
        check(lib().zs_context_async_error(self._h))

    @property
    def launches(self) -> int:
        return int(lib().zs_context_launch_count(self._h))

    def reload_switches(self):
        """re-read the ZS_* A/B switches from the environment (they are read once, at context creation)"""
        check(lib().zs_context_reload_switches(self._h))

    def empty(self, shape, dtype):
        torch = _torch()
        return torch.empty(shape, dtype=dtype, device="cuda:%d" % self.device)

    def to_device(self, a, dtype=None):
        torch = _torch()
        if isinstance(a, np.ndarray):
            a = torch.from_numpy(np.ascontiguousarray(a))
        if dtype is not None:
            a = a.to(dtype)
        return a.to("cuda:%d" % self.device, non_blocking=False).contiguous()


class Pyramid:
    """zs_pyramid: `slots` images with their optical-flow pyramids resident in HBM
    (cv::buildOpticalFlowPyramid layout; utils_opencv.cpp:525-530)."""

    def __init__(self, ctx: Context, width: int, height: int, slots: int, win=(31, 31), max_level=3):
        self.ctx, self.width, self.height, self.slots = ctx, width, height, slots
        self.win, self.max_level = tuple(win), max_level
        h = C.c_void_p()
        check(lib().zs_pyramid_create(ctx._h, width, height, slots, win[0], win[1], max_level, C.byref(h)))
        self._h = h
        self.levels = lib().zs_pyramid_levels(h)

    def close(self):
        if getattr(self, "_h", None) and getattr(self.ctx, "_h", None):
            lib().zs_pyramid_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def level_size(self, level):
        w, h = C.c_int(), C.c_int()
        check(lib().zs_pyramid_level_size(self._h, level, C.byref(w), C.byref(h)))
        return w.value, h.value

    def upload(self, images, first=0):
        """images: (n, H, W) uint8, numpy (host) or torch cuda tensor (device)."""
        if isinstance(images, np.ndarray):
            images = np.ascontiguousarray(images, np.uint8)
            if images.ndim == 2:
                images = images[None]
            n, h, w = images.shape
            assert (h, w) == (self.height, self.width)
            check(lib().zs_pyramid_upload(self.ctx._h, self._h, images.ctypes.data_as(C.c_void_p), w, w * h, first, n, 1))
            self.ctx.synchronize()     # the numpy buffer may go away
        else:
            if images.dim() == 2:
                images = images[None]
            n, h, w = images.shape
            assert (h, w) == (self.height, self.width) and images.is_contiguous()
            check(lib().zs_pyramid_upload(self.ctx._h, self._h, _ptr(images), w, w * h, first, n, 0))
        return n

    def build(self, first=0, count=None):
        check(lib().zs_pyramid_build(self.ctx._h, self._h, first, self.slots if count is None else count))

    def image(self, slot, level):
        w, h = self.level_size(level)
        out = np.empty((h, w), np.uint8)
        check(lib().zs_pyramid_download_image(self.ctx._h, self._h, slot, level, out.ctypes.data_as(C.c_void_p)))
        return out

    def deriv(self, slot, level):
        w, h = self.level_size(level)
        out = np.empty((h, w, 2), np.int16)
        check(lib().zs_pyramid_download_deriv(self.ctx._h, self._h, slot, level, out.ctypes.data_as(C.c_void_p)))
        return out

    def blur(self, slot):
        out = np.empty((self.height, self.width), np.uint8)
        check(lib().zs_orb_download_blur(self.ctx._h, self._h, slot, out.ctypes.data_as(C.c_void_p)))
        return out


# ---------------------------------------------------------------------------------------------
# batched device calls (torch cuda tensors in / out)
# ---------------------------------------------------------------------------------------------
def fast_grid_detect(pyr: Pyramid, first, count, cell=(16, 16), threshold=10, occupied=None):
    """-> xy (count, cap, 2) f32, response (count, cap) f32, n (count,) i32; cap = grid_w*grid_h."""
    torch = _torch()
    ctx = pyr.ctx
    cap = max(1, (pyr.width // cell[0]) * (pyr.height // cell[1]))
    xy = ctx.empty((count, cap, 2), torch.float32)
    resp = ctx.empty((count, cap), torch.float32)
    n = ctx.empty((count,), torch.int32)
    occ = None
    if occupied is not None:
        occ = ctx.to_device(occupied, torch.uint8).reshape(count, -1)
    check(lib().zs_fast_grid_detect(ctx._h, pyr._h, first, count, cell[0], cell[1], int(threshold), _ptr(occ),
                                    _ptr(xy), _ptr(resp), _ptr(n), cap))
    return xy, resp, n


def fast_detect(pyr: Pyramid, first, count, threshold=10, mask=None, cap=65536):
    torch = _torch()
    ctx = pyr.ctx
    xy = ctx.empty((count, cap, 2), torch.float32)
    resp = ctx.empty((count, cap), torch.float32)
    n = ctx.empty((count,), torch.int32)
    m = None
    if mask is not None:
        m = ctx.to_device(mask, torch.uint8).reshape(count, pyr.height, pyr.width)
    check(lib().zs_fast_detect(ctx._h, pyr._h, first, count, int(threshold), _ptr(m), _ptr(xy), _ptr(resp), _ptr(n), cap))
    return xy, resp, n


def corner_subpix(pyr: Pyramid, first, count, xy, n, win=(5, 5), max_iters=30, epsilon=0.01):
    """cv::cornerSubPix on level 0 of slots first..; xy (count, cap, 2) f32 cuda, refined IN PLACE; n (count,) i32"""
    cap = xy.shape[1]
    check(lib().zs_corner_subpix(pyr.ctx._h, pyr._h, first, count, _ptr(xy), _ptr(n), cap, win[0], win[1], max_iters,
                                 float(epsilon)))
    return xy


def orb_compute(pyr: Pyramid, first, count, xy, resp, n, angle=None):
    """-> xy', resp', src_index, n', desc (count, cap, 32) u8"""
    torch = _torch()
    ctx = pyr.ctx
    cap = xy.shape[1]
    oxy = ctx.empty((count, cap, 2), torch.float32)
    oresp = ctx.empty((count, cap), torch.float32)
    src = ctx.empty((count, cap), torch.int32)
    on = ctx.empty((count,), torch.int32)
    desc = ctx.empty((count, cap, 32), torch.uint8)
    check(lib().zs_orb_compute(ctx._h, pyr._h, first, count, _ptr(xy), _ptr(resp), _ptr(angle), _ptr(n), cap,
                               _ptr(oxy), _ptr(oresp), _ptr(src), _ptr(on), _ptr(desc)))
    return oxy, oresp, src, on, desc


class OrbDetector:
    """zs_orb_detector: cv::ORB::create(nfeatures, scale_factor, nlevels, edge, 0, 2, HARRIS_SCORE, patch, fast_threshold)
    ->detect(image, mask) [+ cv::ORB::create()->compute] for batches of up to max_images frames of one size
    (keypoint_detector_simple.cpp:17,27,49,54).  Keypoints come back in canonical order (octave, y, x)."""

    def __init__(self, ctx: Context, width: int, height: int, max_images: int = 1, nfeatures: int = 500,
                 scale_factor: float = 1.2, nlevels: int = 8, edge_threshold: int = 31, patch_size: int = 31,
                 fast_threshold: int = 20):
        self.ctx, self.width, self.height, self.max_images, self.nlevels = ctx, width, height, max_images, nlevels
        h = C.c_void_p()
        check(lib().zs_orb_detector_create(ctx._h, width, height, max_images, nfeatures, float(scale_factor), nlevels,
                                           edge_threshold, patch_size, fast_threshold, C.byref(h)))
        self._h = h
        self.cap = lib().zs_orb_detector_capacity(h)

    def close(self):
        if getattr(self, "_h", None):
            lib().zs_orb_detector_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def level(self, level):
        w, h, s, n = C.c_int(), C.c_int(), C.c_float(), C.c_int()
        check(lib().zs_orb_detector_level(self._h, level, C.byref(w), C.byref(h), C.byref(s), C.byref(n)))
        return w.value, h.value, s.value, n.value

    def download_level(self, image, level, which=0):
        import numpy as np
        w, h, _, _ = self.level(level)
        out = np.empty((h, w), np.uint8)
        check(lib().zs_orb_detector_download_level(self.ctx._h, self._h, image, level, which, out.ctypes.data_as(C.c_void_p)))
        return out

    def detect_and_compute(self, images, masks=None, describe=True):
        """images / masks: (count, H, W) u8 (numpy or cuda tensors).  -> dict of cuda tensors: xy (count, cap, 2),
        size, angle, response (count, cap) f32, octave (count, cap) i32, n (count,) i32[, desc (count, cap, 32) u8]"""
        torch = _torch()
        ctx = self.ctx
        img = ctx.to_device(images, torch.uint8).reshape(-1, self.height, self.width).contiguous()
        count = img.shape[0]
        m = None
        if masks is not None:
            m = ctx.to_device(masks, torch.uint8).reshape(count, self.height, self.width).contiguous()
        cap = self.cap
        out = dict(xy=ctx.empty((count, cap, 2), torch.float32), size=ctx.empty((count, cap), torch.float32),
                   angle=ctx.empty((count, cap), torch.float32), response=ctx.empty((count, cap), torch.float32),
                   octave=ctx.empty((count, cap), torch.int32), n=ctx.empty((count,), torch.int32))
        if describe:
            out["desc"] = ctx.empty((count, cap, 32), torch.uint8)
        plane = self.width * self.height
        check(lib().zs_orb_detect_and_compute(ctx._h, self._h, _ptr(img), self.width, plane, _ptr(m), self.width, plane, count,
                                              _ptr(out["xy"]), _ptr(out["size"]), _ptr(out["angle"]), _ptr(out["response"]),
                                              _ptr(out["octave"]), _ptr(out["n"]), _ptr(out.get("desc"))))
        return out


def match_hamming_knn2(ctx: Context, q, nq, t, nt, ratio=0.8):
    """q (pairs, cap_q, 32) u8, t (pairs, cap_t, 32) u8 -> idx (pairs, cap_q, 2), dist, pass"""
    torch = _torch()
    pairs, cap_q, cap_t = q.shape[0], q.shape[1], t.shape[1]
    idx = ctx.empty((pairs, cap_q, 2), torch.int32)
    dist = ctx.empty((pairs, cap_q, 2), torch.float32)
    ps = ctx.empty((pairs, cap_q), torch.uint8)
    check(lib().zs_match_hamming_knn2(ctx._h, _ptr(q), _ptr(nq), cap_q * 32, _ptr(t), _ptr(nt), cap_t * 32, pairs,
                                      cap_q, cap_t, float(ratio), _ptr(idx), _ptr(dist), _ptr(ps)))
    return idx, dist, ps


def match_hamming_cross(ctx: Context, q, nq, t, nt):
    torch = _torch()
    pairs, cap_q, cap_t = q.shape[0], q.shape[1], t.shape[1]
    idx = ctx.empty((pairs, cap_q), torch.int32)
    dist = ctx.empty((pairs, cap_q), torch.float32)
    check(lib().zs_match_hamming_cross(ctx._h, _ptr(q), _ptr(nq), cap_q * 32, _ptr(t), _ptr(nt), cap_t * 32, pairs,
                                       cap_q, cap_t, _ptr(idx), _ptr(dist)))
    return idx, dist


def match_l2_knn2(ctx: Context, q, nq, t, nt, ratio=0.8):
    """q (pairs, cap_q, dim) f32 integer-valued, or u8 (cv::SIFT with descriptorType CV_8U).  Stream-asynchronous: float rows
    that are not integers in 0..255 come back as all -1 and make ctx.async_error() raise."""
    torch = _torch()
    pairs, cap_q, dim = q.shape
    cap_t = t.shape[1]
    idx = ctx.empty((pairs, cap_q, 2), torch.int32)
    dist = ctx.empty((pairs, cap_q, 2), torch.float32)
    ps = ctx.empty((pairs, cap_q), torch.uint8)
    if q.dtype == torch.uint8:
        check(lib().zs_match_l2_knn2_u8(ctx._h, _ptr(q), _ptr(nq), _ptr(t), _ptr(nt), pairs, cap_q, cap_t, dim, float(ratio),
                                        _ptr(idx), _ptr(dist), _ptr(ps)))
        return idx, dist, ps
    check(lib().zs_match_l2_knn2(ctx._h, _ptr(q), _ptr(nq), cap_q * dim, _ptr(t), _ptr(nt), cap_t * dim, pairs,
                                 cap_q, cap_t, dim, float(ratio), _ptr(idx), _ptr(dist), _ptr(ps)))
    return idx, dist, ps


def match_l2_cross(ctx: Context, q, nq, t, nt):
    torch = _torch()
    pairs, cap_q, dim = q.shape
    cap_t = t.shape[1]
    idx = ctx.empty((pairs, cap_q), torch.int32)
    dist = ctx.empty((pairs, cap_q), torch.float32)
    if q.dtype == torch.uint8:
        check(lib().zs_match_l2_cross_u8(ctx._h, _ptr(q), _ptr(nq), _ptr(t), _ptr(nt), pairs, cap_q, cap_t, dim, _ptr(idx), _ptr(dist)))
        return idx, dist
    check(lib().zs_match_l2_cross(ctx._h, _ptr(q), _ptr(nq), cap_q * dim, _ptr(t), _ptr(nt), cap_t * dim, pairs,
                                  cap_q, cap_t, dim, _ptr(idx), _ptr(dist)))
    return idx, dist


def klt_track(pyr: Pyramid, prev_slot, next_slot, prev_pts, count, lk: LK, next_pts=None, fb_threshold=None):
    """prev_pts (jobs, cap, 2) f32; count (jobs,) i32; prev_slot/next_slot (jobs,) i32 (all cuda).
    -> next_pts, status, err[, keep]"""
    torch = _torch()
    ctx = pyr.ctx
    jobs, cap = prev_pts.shape[0], prev_pts.shape[1]
    if next_pts is None:
        assert not (lk.flags & LK_USE_INITIAL_FLOW)
        next_pts = ctx.empty((jobs, cap, 2), torch.float32)
        next_pts.zero_()
    status = ctx.empty((jobs, cap), torch.uint8); status.zero_()
    err = ctx.empty((jobs, cap), torch.float32); err.zero_()
    prm = lk.c()
    if fb_threshold is None:
        check(lib().zs_klt_track(ctx._h, pyr._h, _ptr(prev_slot), _ptr(next_slot), _ptr(prev_pts), _ptr(next_pts),
                                 _ptr(count), jobs, cap, C.byref(prm), _ptr(status), _ptr(err)))
        return next_pts, status, err
    keep = ctx.empty((jobs, cap), torch.uint8); keep.zero_()
    check(lib().zs_klt_track_fb(ctx._h, pyr._h, _ptr(prev_slot), _ptr(next_slot), _ptr(prev_pts), _ptr(next_pts),
                                _ptr(count), jobs, cap, C.byref(prm), float(fb_threshold), _ptr(status), _ptr(err),
                                _ptr(keep)))
    return next_pts, status, err, keep
