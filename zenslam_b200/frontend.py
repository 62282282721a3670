"""Batched stereo front-end (zs_frontend): B consecutive stereo frames per call through the C ABI.

Per frame: 2 pyramids, 2 grid detections + ORB, 1 stereo kNN-ratio match and 4 forward+backward KLT
pairs -- the call pattern of keypoint_tracker::track (keypoint_tracker.cpp:41-105) without the map-dependent
steps.  `process(left, right)` is the end-to-end call with host buffers; `upload` / `run` / `download`
split it for HBM-resident timing.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import FrontendOptions, FrontendResults, check, lib
from .options import slam_options
from .runtime import Context, _ptr

KINDS = ("temporal_left", "temporal_right", "stereo_lr", "stereo_rl")


class StereoFrontend:
    def __init__(self, ctx: Context, width: int, height: int, batch: int, options: slam_options | None = None,
                 max_iters: int = 99, epsilon: float = 0.001, min_eig_threshold: float = 1e-4):
        o = options or slam_options()
        self.ctx, self.width, self.height, self.batch, self.options = ctx, width, height, batch, o
        self._opt = FrontendOptions(width, height, batch, o.detection.cell_size[0], o.detection.cell_size[1],
                                    o.detection.fast_threshold, o.tracking.klt_window_size[0], o.tracking.klt_window_size[1],
                                    o.tracking.klt_max_level, o.tracking.klt_threshold, o.matcher_ratio, max_iters, epsilon,
                                    min_eig_threshold, 1 if o.detection.algorithm == "PARALLEL_GRID" else 0)
        h = C.c_void_p()
        check(lib().zs_frontend_create(ctx._h, C.byref(self._opt), C.byref(h)))
        self._h = h
        self.cap = lib().zs_frontend_capacity(h)
        self.h2d_bytes = int(lib().zs_frontend_h2d_bytes(h))
        self.d2h_bytes = int(lib().zs_frontend_d2h_bytes(h))
        B, cap = batch, self.cap
        self._host = None
        self._submitted = self._waited = 0
        self._channels = 1
        self._shapes = dict(
            n_left=((B,), np.int32), n_right=((B,), np.int32),
            kp_left=((B, cap, 2), np.float32), kp_right=((B, cap, 2), np.float32),
            resp_left=((B, cap), np.float32), resp_right=((B, cap), np.float32),
            desc_left=((B, cap, 32), np.uint8), desc_right=((B, cap, 32), np.uint8),
            match_idx=((B, cap, 2), np.int32), match_dist=((B, cap, 2), np.float32), match_pass=((B, cap), np.uint8),
            track_pts=((4, B, cap, 2), np.float32), track_keep=((4, B, cap), np.uint8), track_n=((4, B), np.int32))

    def close(self):
        if getattr(self, "_h", None) and getattr(self.ctx, "_h", None):
            lib().zs_frontend_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _results(self, which: int = 0):
        """pinned host result buffers (two sets, allocated once) + the C structs that point at them"""
        if self._host is None:
            import torch
            self._host, self._res = [], []
            for _ in range(2):
                host = {k: torch.empty(s, dtype=getattr(torch, np.dtype(d).name), pin_memory=True)
                        for k, (s, d) in self._shapes.items()}
                r = FrontendResults()
                r.cap = self.cap
                for k, t in host.items():
                    setattr(r, k, t.data_ptr())
                self._host.append(host); self._res.append(r)
        return self._host[which], self._res[which]

    def set_preprocess(self, channels: int = 1, clahe_enabled: bool = False, clahe_clip_limit: float = 4.0, maps=None):
        """raw camera frames in (processor::process, processor.cpp:25-55): channels 3 = BGR; maps = ((map_x_l, map_y_l),
        (map_x_r, map_y_r)) float32 (H, W) arrays or None.  submit()/process() then take (B, H, W[, 3]) raw frames."""
        ptrs = [None] * 4
        keep = []
        if maps is not None:
            for i, m in enumerate((maps[0][0], maps[0][1], maps[1][0], maps[1][1])):
                a = np.ascontiguousarray(m, np.float32)
                assert a.shape == (self.height, self.width)
                keep.append(a); ptrs[i] = a.ctypes.data_as(C.c_void_p)
        check(lib().zs_frontend_set_preprocess(self._h, channels, 1 if clahe_enabled else 0, float(clahe_clip_limit), *ptrs))
        self._channels = channels
        self.h2d_bytes = int(lib().zs_frontend_h2d_bytes(self._h))

    def upload(self, left, right):
        """left/right: (B, H, W) uint8 -- numpy / pinned torch CPU tensor (host) or torch cuda tensor."""
        host = isinstance(left, np.ndarray) or not left.is_cuda
        lp = left.ctypes.data if isinstance(left, np.ndarray) else left.data_ptr()
        rp = right.ctypes.data if isinstance(right, np.ndarray) else right.data_ptr()
        assert tuple(left.shape) == (self.batch, self.height, self.width) == tuple(right.shape)
        check(lib().zs_frontend_upload(self._h, C.c_void_p(lp), C.c_void_p(rp), self.width, self.width * self.height,
                                       1 if host else 0))

    def run(self):
        check(lib().zs_frontend_run(self._h))

    STAGES = ("pyramid", "fast_grid", "orb", "match", "klt", "carry")

    def timing_enable(self, on=True):
        check(lib().zs_frontend_timing_enable(self._h, 1 if on else 0))

    def timing_collect(self):
        """-> ({stage: mean ms per run}, runs) measured with CUDA events on the context's stream"""
        ms = (C.c_float * 6)()
        runs = C.c_int(0)
        check(lib().zs_frontend_timing_collect(self._h, ms, C.byref(runs)))
        n = max(1, runs.value)
        return {k: ms[i] / n for i, k in enumerate(self.STAGES)}, runs.value

    def download(self) -> dict:
        host, res = self._results()
        check(lib().zs_frontend_download(self._h, C.byref(res)))
        return {k: t.numpy() for k, t in host.items()}

    def process(self, left, right) -> dict:
        """End-to-end: H2D of the batch, the whole hot path, D2H of every result (synchronous)."""
        assert self._submitted == self._waited, "process() while submissions are outstanding: call wait() first"
        host, res = self._results()
        lp = left.ctypes.data if isinstance(left, np.ndarray) else left.data_ptr()
        rp = right.ctypes.data if isinstance(right, np.ndarray) else right.data_ptr()
        check(lib().zs_frontend_process_host(self._h, C.c_void_p(lp), C.c_void_p(rp), self.width * self._channels,
                                             self.width * self.height * self._channels, C.byref(res)))
        return {k: t.numpy() for k, t in host.items()}

    # ---- pipelined end-to-end path: up to two batches in flight (H2D | kernels | D2H overlap) ----
    def submit(self, left, right):
        """enqueue one batch (host buffers, ideally pinned); returns at once.  Call wait() for the results,
        in submission order; at most two submissions may be outstanding."""
        host, res = self._results(self._submitted & 1)
        lp = left.ctypes.data if isinstance(left, np.ndarray) else left.data_ptr()
        rp = right.ctypes.data if isinstance(right, np.ndarray) else right.data_ptr()
        check(lib().zs_frontend_submit_host(self._h, C.c_void_p(lp), C.c_void_p(rp), self.width * self._channels,
                                            self.width * self.height * self._channels, C.byref(res)))
        self._submitted += 1

    def wait(self) -> dict:
        """block until the oldest outstanding submission has landed in host memory; the returned arrays are
        views of a pinned buffer that is reused two submissions later"""
        host, _ = self._results(self._waited & 1)
        check(lib().zs_frontend_wait(self._h))
        self._waited += 1
        return {k: t.numpy() for k, t in host.items()}

    @property
    def in_flight(self) -> int:
        return int(lib().zs_frontend_in_flight(self._h))
