"""Seeded random sweep over shapes and options (odd frame sizes, every window the kernels specialise on and some they do
not, thresholds, cell sizes, ragged counts, points on and beyond the borders, noisy / flat / quantised images): every
stage through the C ABI against the oracle, bit for bit.  Sized to finish in well under a minute on a B200."""
import numpy as np
import pytest

import oracle
from zenslam_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from zenslam_b200.runtime import Context
    c = Context()
    yield c
    c.close()


def dev(ctx, a):
    return ctx.to_device(np.ascontiguousarray(a))


def _image(rng, w, h, kind):
    base = syn.crop(syn.base_texture(w, h, int(rng.integers(1, 1 << 30))), w, h, 0, 0)
    if kind == 1:                                   # heavy quantisation: ties everywhere (FAST maxima, equal distances)
        base = (base // 32 * 32).astype(np.uint8)
    elif kind == 2:                                 # a flat block: zero gradients, LK min-eigenvalue rejection
        base = base.copy(); base[h // 4:h // 2, w // 4:w // 2] = 77
    elif kind == 3:                                 # salt-and-pepper extremes
        base = base.copy(); m = rng.random((h, w)) < 0.02; base[m] = rng.choice([0, 255], int(m.sum()))
    return np.ascontiguousarray(base)


@pytest.mark.parametrize("seed", range(18))
def test_detect_describe_match_random(ctx, seed):
    from zenslam_b200.runtime import Pyramid, fast_grid_detect, match_hamming_cross, match_hamming_knn2, orb_compute
    rng = np.random.default_rng(100 + seed)
    w, h = int(rng.integers(97, 420)), int(rng.integers(80, 300))
    cell = [(16, 16), (32, 32), (16, 16), (24, 20), (13, 17), (32, 32)][seed % 6]
    thr = int(rng.choice([0, 5, 10, 20, 40]))
    imgs = np.stack([_image(rng, w, h, (seed + k) % 4) for k in range(3)])
    pyr = Pyramid(ctx, w, h, 3, (15, 15), 1)
    pyr.upload(imgs, 0); pyr.build(0, 3)
    gw, gh = w // cell[0], h // cell[1]
    occ = (rng.random((3, gh, gw)) < 0.2).astype(np.uint8)
    xy, resp, n = fast_grid_detect(pyr, 0, 3, cell, thr, occ)
    oxy, oresp, src, on, desc = orb_compute(pyr, 0, 3, xy, resp, n)
    descs, counts = [], []
    for k in range(3):
        x, y, s = oracle.grid_detect(imgs[k], cell, thr, occ[k])
        nk = int(n[k])
        assert nk == len(x), (k, nk, len(x))
        assert np.array_equal(xy[k, :nk].cpu().numpy(), np.stack([x, y], 1).astype(np.float32))
        assert np.array_equal(resp[k, :nk].cpu().numpy(), s.astype(np.float32))
        kept, d = oracle.orb_compute(imgs[k], x, y)
        m = int(on[k])
        assert m == len(kept) and np.array_equal(src[k, :m].cpu().numpy(), kept)
        assert np.array_equal(desc[k, :m].cpu().numpy(), d)
        descs.append(d); counts.append(m)
    if min(counts) >= 2:
        cap = desc.shape[1]
        q = desc[:2].contiguous(); t = desc[1:3].contiguous()
        nq, nt = on[:2].contiguous(), on[1:3].contiguous()
        idx, dist, ps = match_hamming_knn2(ctx, q, nq, t, nt, 0.8)
        cidx, cdist = match_hamming_cross(ctx, q, nq, t, nt)
        for k in range(2):
            oi, od = oracle.match_hamming_knn2(descs[k], descs[k + 1])
            a = counts[k]
            assert np.array_equal(idx[k, :a].cpu().numpy(), oi) and np.array_equal(dist[k, :a].cpu().numpy(), od.astype(np.float32))
            assert np.array_equal(np.nonzero(ps[k, :a].cpu().numpy())[0], oracle.ratio_test(oi, od, 0.8)[0])
            oq, ot, odd = oracle.match_hamming_cross(descs[k], descs[k + 1])
            got = cidx[k, :a].cpu().numpy()
            keep = np.nonzero(got >= 0)[0]
            assert np.array_equal(keep, oq) and np.array_equal(got[keep], ot)
        assert cap >= max(counts)
    pyr.close()


@pytest.mark.parametrize("seed", range(24))
def test_klt_random(ctx, seed):
    from zenslam_b200 import LK_GET_MIN_EIGENVALS, LK_USE_INITIAL_FLOW
    from zenslam_b200.runtime import LK, Pyramid, klt_track
    rng = np.random.default_rng(200 + seed)
    win = [(31, 31), (21, 21), (15, 15), (9, 9), (31, 31), (25, 13), (31, 31), (41, 41)][seed % 8]
    ml = int(rng.integers(0, 5))
    w, h = int(rng.integers(120, 500)), int(rng.integers(100, 360))
    base = syn.base_texture(w, h, int(rng.integers(1, 1 << 30)))
    dx, dy = rng.uniform(-7, 7, 2)
    A = syn.crop(base, w, h, 0, 0); B = syn.crop(base, w, h, float(dx), float(dy))
    if seed % 3 == 2:
        A = A.copy(); A[h // 3:h // 2, w // 3:w // 2] = 128      # textureless block
    n = int(rng.integers(1, 400))
    pts = np.stack([rng.uniform(-20, w + 20, n), rng.uniform(-20, h + 20, n)], 1).astype(np.float32)
    pts[: n // 4] = np.rint(pts[: n // 4])                           # integer positions: zero fractional weights
    init = None
    if seed % 2:
        init = (pts + rng.uniform(-12, 12, (n, 2))).astype(np.float32)
    p = Pyramid(ctx, w, h, 2, win, ml)
    p.upload(np.stack([A, B]), 0); p.build(0, 2)
    flags = LK_GET_MIN_EIGENVALS | (LK_USE_INITIAL_FLOW if init is not None else 0)
    lk = LK(win, ml, 99, 0.001, flags, 1e-4)
    out = klt_track(p, dev(ctx, np.array([0], np.int32)), dev(ctx, np.array([1], np.int32)), dev(ctx, pts[None]),
                    dev(ctx, np.array([n], np.int32)), lk, None if init is None else dev(ctx, init[None].copy()), 1.0)
    p1, st, err, keep = [o[0].cpu().numpy() for o in out]
    PA, PB = oracle.Pyramid(A, win, ml), oracle.Pyramid(B, win, ml)
    o1, os_, oe = oracle.lk_track(PA, PB, pts, init, win, ml, flags=flags)
    ob, osb, _ = oracle.lk_track(PB, PA, o1, None, win, ml)
    assert np.array_equal(st[:n], os_) and np.array_equal(p1[:n], o1) and np.array_equal(err[:n], oe), (win, ml, w, h)
    assert np.array_equal(keep[:n].astype(bool), oracle.fb_check(pts, ob, os_, osb, 1.0))
    p.close()
