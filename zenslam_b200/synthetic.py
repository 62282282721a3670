"""Seeded synthetic stereo sequences of the BASELINE.json shapes.

A textured base image (Gaussian-blurred uniform noise, sigma 1.6, min-max normalised to u8) is
cropped at integer offsets, so the ground-truth flow between views is exact: the right view is
the left view shifted by ``disparity`` px in x; frame t is displaced by a bounded random walk.
A sub-pixel variant resamples the crop bilinearly at a fractional offset.
Pure NumPy (no cv2) so that the product-side bench can use it.
"""
from __future__ import annotations

import numpy as np

MARGIN = 64


def _gauss_kernel(sigma: float) -> np.ndarray:
    r = int(np.ceil(4 * sigma))
    x = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-0.5 * (x / sigma) ** 2)
    return k / k.sum()


def _blur(img: np.ndarray, sigma: float) -> np.ndarray:
    k = _gauss_kernel(sigma)
    r = len(k) // 2
    p = np.pad(img, ((0, 0), (r, r)), mode="reflect")
    out = np.zeros_like(img)
    for i, kv in enumerate(k):
        out += kv * p[:, i:i + img.shape[1]]
    p = np.pad(out, ((r, r), (0, 0)), mode="reflect")
    out2 = np.zeros_like(img)
    for i, kv in enumerate(k):
        out2 += kv * p[i:i + img.shape[0], :]
    return out2


def base_texture(width: int, height: int, seed: int, margin: int = MARGIN) -> np.ndarray:
    rng = np.random.default_rng(seed)
    base = np.floor(rng.random((height + 2 * margin, width + 2 * margin)) * 255.0)
    b = _blur(base, 1.6)
    b = (b - b.min()) / (b.max() - b.min()) * 255.0
    return np.clip(np.rint(b), 0, 255).astype(np.uint8)


def crop(base: np.ndarray, width: int, height: int, ox: float, oy: float, margin: int = MARGIN) -> np.ndarray:
    """View of the base at offset (ox, oy) from the centred crop; bilinear if fractional."""
    fx, fy = float(ox) + margin, float(oy) + margin
    ix, iy = int(np.floor(fx)), int(np.floor(fy))
    ax, ay = fx - ix, fy - iy
    if ax == 0.0 and ay == 0.0:
        return np.ascontiguousarray(base[iy:iy + height, ix:ix + width])
    t = base[iy:iy + height + 1, ix:ix + width + 1].astype(np.float64)
    v = ((1 - ax) * (1 - ay) * t[:-1, :-1] + ax * (1 - ay) * t[:-1, 1:]
         + (1 - ax) * ay * t[1:, :-1] + ax * ay * t[1:, 1:])
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def stereo_pair(width: int, height: int, seed: int, disparity: int = 6):
    """C1: left = centred crop, right = crop shifted ``disparity`` px in x."""
    base = base_texture(width, height, seed)
    return crop(base, width, height, 0, 0), crop(base, width, height, disparity, 0)


def stereo_sequence(width: int, height: int, frames: int, seed: int, disparity: int = 6,
                    max_step: int = 9, subpixel: bool = False):
    """C2: (frames, 2, H, W) u8; per-frame motion is a random walk with steps <= max_step px,
    clamped to the margin.  ``subpixel`` adds a fractional part to every offset."""
    base = base_texture(width, height, seed)
    rng = np.random.default_rng(seed + 7919)
    out = np.empty((frames, 2, height, width), np.uint8)
    offs = np.zeros((frames, 2), np.float64)
    ox = oy = 0.0
    lim = MARGIN - disparity - 2
    for t in range(frames):
        if t:
            ox = float(np.clip(ox + rng.integers(-max_step, max_step + 1), -lim, lim))
            oy = float(np.clip(oy + rng.integers(-max_step, max_step + 1), -lim, lim))
        fx = fy = 0.0
        if subpixel:
            fx, fy = rng.random(2) * 0.9
        offs[t] = (ox + fx, oy + fy)
        out[t, 0] = crop(base, width, height, ox + fx, oy + fy)
        out[t, 1] = crop(base, width, height, ox + fx + disparity, oy + fy)
    return out, offs
