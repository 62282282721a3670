"""CPU: the oracle's restatement of the reference's detector GLUE against the reference's OWN code.

oracle/_ref/libzs_ref_glue.so is zenslam_core/source/detection/keypoint_detector_{grid,parallel,simple}.cpp compiled unmodified
from /root/reference (oracle/build_ref.py) on an OpenCV stand-in whose cv::FAST / cv::ORB / cv::cornerSubPix are the C oracle's
(pinned to real cv2 elsewhere).  What is compared here is therefore everything the reference does AROUND those primitives --
cell geometry and the cv::Size division, the occupancy grid from keypoints_existing, the ROI per cell, the strongest keypoint
per cell (first on ties), cell row-major order, cv::cornerSubPix on the selected corners, the disc mask of the SIMPLE detector,
ORB's border filter, sequential keypoint::index_next indices -- as restated in oracle/zs_oracle.c (zso_grid_detect) and
oracle/__init__.py, which in turn is what the CUDA path is held to bit for bit in the GPU tests.
Skipped where oracle/_ref has not been built (no /root/reference)."""
import numpy as np
import pytest

import oracle
from oracle import reference_glue as ref
from zenslam_b200 import synthetic as syn

if not ref.available():
    try:
        from oracle import build_ref
        if build_ref.available():
            build_ref.build()
    except Exception:
        pass

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libzs_ref_glue.so not built (needs /root/reference)")


def _frame(rng, w, h, kind):
    if kind == 0:
        L, _ = syn.stereo_pair(w, h, int(rng.integers(1, 1 << 30)))
        return L
    if kind == 1:                                     # coarse quantisation: many equal responses -> tie rules
        L, _ = syn.stereo_pair(w, h, int(rng.integers(1, 1 << 30)))
        return (L // 32 * 32).astype(np.uint8)
    img = rng.integers(0, 256, (h, w)).astype(np.uint8)      # noise: nearly every pixel a corner
    img[h // 3: h // 2, w // 4: w // 2] = 90          # and a flat block: empty cells
    return img


def _existing(rng, w, h, n):
    ex = np.stack([rng.permutation(5000)[:n] + 10, rng.uniform(-20, w + 20, n), rng.uniform(-20, h + 20, n)], 1).astype(np.float32)
    ex[: n // 3, 1:] = np.floor(ex[: n // 3, 1:])     # on cell boundaries now and then
    return ex


def _occupancy(ex, w, h, cell):
    gw, gh = w // cell[0], h // cell[1]
    occ = np.zeros((gh, gw), np.uint8)
    for _, x, y in ex:
        gx, gy = int(int(x) / cell[0]), int(int(y) / cell[1])    # gsl::narrow_cast<int> and the integer division both truncate toward
                                                                  # zero: a keypoint at x in (-cell, 0) occupies column 0
        if 0 <= gx < gw and 0 <= gy < gh:
            occ[gy, gx] = 1
    return occ


CASES = [(320, 240, (16, 16), 10, 0), (333, 247, (16, 16), 10, 1), (752, 480, (32, 32), 7, 0), (200, 150, (24, 17), 5, 2),
         (640, 400, (64, 64), 1, 0), (129, 131, (16, 16), 20, 2), (310, 204, (40, 40), 10, 1)]


@pytest.mark.parametrize("w,h,cell,thr,kind", CASES)
def test_grid_detector_glue(w, h, cell, thr, kind):
    rng = np.random.default_rng(w * 7 + h)
    img = _frame(rng, w, h, kind)
    for ex in (None, _existing(rng, w, h, 60)):
        got = ref.detect_keypoints(ref.GRID, img, cell, thr, ex, index_next=41)
        occ = None if ex is None else _occupancy(ex, w, h, cell)
        x, y, s = oracle.grid_detect(img, cell, thr, occ)
        kept, desc = oracle.orb_compute(img, x, y)
        assert len(got["xy"]) == len(kept)
        assert np.array_equal(got["xy"], np.stack([x[kept], y[kept]], 1).astype(np.float32))
        assert np.array_equal(got["response"], s[kept].astype(np.float32))
        assert np.array_equal(got["desc"], desc)
        assert np.array_equal(got["index"], 41 + np.arange(len(kept))) and got["index_next"] == 41 + len(kept)
        assert np.all(got["size"] == 7) and np.all(got["angle"] == -1) and np.all(got["octave"] == 0)
    assert len(kept) > 0 or kind == 2


@pytest.mark.parametrize("w,h,cell,thr,kind", CASES[:5])
def test_parallel_grid_detector_glue(w, h, cell, thr, kind):
    rng = np.random.default_rng(w * 11 + h)
    img = _frame(rng, w, h, kind)
    ex = _existing(rng, w, h, 40)
    got = ref.detect_keypoints(ref.PARALLEL_GRID, img, cell, thr, ex, index_next=5)
    x, y, s = oracle.grid_detect(img, cell, thr, _occupancy(ex, w, h, cell))
    refined = oracle.corner_subpix(img, np.stack([x, y], 1).astype(np.float32))
    kept, desc = oracle.orb_compute(img, refined[:, 0].copy(), refined[:, 1].copy())
    assert np.array_equal(got["xy"], refined[kept]) and np.array_equal(got["response"], s[kept].astype(np.float32))
    assert np.array_equal(got["desc"], desc) and np.array_equal(got["index"], 5 + np.arange(len(kept)))
    assert len(kept) > 0


@pytest.mark.parametrize("w,h,thr,kind", [(320, 240, 20, 0), (257, 199, 10, 1), (160, 120, 40, 2)])
def test_simple_detector_glue(w, h, thr, kind):
    rng = np.random.default_rng(w * 13 + h)
    img = _frame(rng, w, h, kind)
    x, y, s = oracle.fast_detect(img, thr)
    for ex in (None, _existing(rng, w, h, 50)):
        got = ref.detect_keypoints(ref.SIMPLE, img, (16, 16), thr, ex, index_next=0)
        keep = np.ones(len(x), bool)
        if ex is not None:
            for _, ex_x, ex_y in ex:                  # cv::circle(mask, Point(pt) [cvRound], min(cell) / 2, 0, filled)
                cx, cy = int(np.rint(np.float32(ex_x))), int(np.rint(np.float32(ex_y)))
                keep &= ~((x - cx) ** 2 + (y - cy) ** 2 <= 8 * 8)
        kept, desc = oracle.orb_compute(img, x[keep], y[keep])
        assert np.array_equal(got["xy"], np.stack([x[keep][kept], y[keep][kept]], 1).astype(np.float32))
        assert np.array_equal(got["response"], s[keep][kept].astype(np.float32)) and np.array_equal(got["desc"], desc)
        assert np.array_equal(got["index"], np.arange(len(kept))) and got["index_next"] == len(kept)
