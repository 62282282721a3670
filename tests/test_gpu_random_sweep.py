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
def test_klt_random(ctx, switches, seed):
    from zenslam_b200 import LK_GET_MIN_EIGENVALS, LK_USE_INITIAL_FLOW
    from zenslam_b200.runtime import LK, Pyramid, klt_track
    if seed % 4 == 3:
        switches.set(ctx, "ZS_KLT_PERSIST_MIN", 1)                       # the persistent launch form on a small problem
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


@pytest.mark.parametrize("form", ["two_tiles", "four_warps", "two_tiles_persistent", "four_warps_persistent",
                                  "packed6", "packed7_persistent", "unpacked", "unpacked_persistent"])
@pytest.mark.parametrize("seed", range(6))
def test_klt_random_63(ctx, switches, seed, form):
    """63 x 63 windows (tumvi.yaml:45) on random frames against the oracle, through the tiled forms of the tracker -- two
    tiles per warp with the packed template (the default; also at 6 / 7 CTAs per SM), two tiles per warp with one register per
    pixel (ZS_KLT63_UNPACKED) and one tile per warp (ZS_KLT63_FOUR_WARPS) -- each also in its persistent launch form: frames
    smaller than the window at the coarse levels, points outside the frame, integer positions, initial flow, several jobs in
    one launch (so that the persistent form has items to distribute)."""
    from zenslam_b200 import LK_GET_MIN_EIGENVALS, LK_USE_INITIAL_FLOW
    from zenslam_b200.runtime import LK, Pyramid, klt_track
    if form.startswith("four_warps"):
        switches.set(ctx, "ZS_KLT63_FOUR_WARPS")
    if form.startswith("packed"):                                        # the packed-template two-tile kernel at 6 / 7 CTAs per SM (default: 8)
        switches.set(ctx, "ZS_KLT63_PACKED", form[6])
    if form.startswith("unpacked"):                                      # register-per-pixel template (the round-2 form before the packed one)
        switches.set(ctx, "ZS_KLT63_UNPACKED")
    if form.endswith("persistent"):
        switches.set(ctx, "ZS_KLT_PERSIST_MIN", 1)                       # small launches do not take the persistent form by themselves
    rng = np.random.default_rng(900 + seed)
    win, ml = (63, 63), int(rng.integers(0, 5))
    w, h = int(rng.integers(130, 700)), int(rng.integers(130, 520))
    base = syn.base_texture(w, h, int(rng.integers(1, 1 << 30)))
    dx, dy = rng.uniform(-9, 9, 2)
    A = syn.crop(base, w, h, 0, 0); B = syn.crop(base, w, h, float(dx), float(dy))
    if seed % 3 == 1:
        A = A.copy(); A[h // 4:h // 2, w // 4:w // 2] = 77               # textureless block: min-eigenvalue rejections
    jobs = 3 if seed % 2 else 1
    n = int(rng.integers(40, 700))
    pts = np.stack([rng.uniform(-40, w + 40, (jobs, n)), rng.uniform(-40, h + 40, (jobs, n))], -1).astype(np.float32)
    pts[:, : n // 4] = np.rint(pts[:, : n // 4])
    init = (pts + rng.uniform(-15, 15, pts.shape)).astype(np.float32) if seed % 2 == 0 else None
    p = Pyramid(ctx, w, h, 2, win, ml)
    p.upload(np.stack([A, B]), 0); p.build(0, 2)
    flags = LK_GET_MIN_EIGENVALS | (LK_USE_INITIAL_FLOW if init is not None else 0)
    lk = LK(win, ml, 99, 0.001, flags, 1e-4)
    prev_slot = dev(ctx, np.array([0, 1, 0][:jobs], np.int32)); next_slot = dev(ctx, np.array([1, 0, 1][:jobs], np.int32))
    cnt = np.array([n, n - 7, n // 2][:jobs], np.int32)
    out = klt_track(p, prev_slot, next_slot, dev(ctx, pts), dev(ctx, cnt), lk, None if init is None else dev(ctx, init.copy()), 2.0)
    p1, st, err, keep = [o.cpu().numpy() for o in out]
    PA, PB = oracle.Pyramid(A, win, ml), oracle.Pyramid(B, win, ml)
    for j in range(jobs):
        P0, P1 = (PA, PB) if j != 1 else (PB, PA)
        m = int(cnt[j])
        o1, os_, oe = oracle.lk_track(P0, P1, pts[j, :m], None if init is None else init[j, :m], win, ml, flags=flags)
        ob, osb, _ = oracle.lk_track(P1, P0, o1, None, win, ml)
        assert np.array_equal(st[j, :m], os_) and np.array_equal(p1[j, :m], o1) and np.array_equal(err[j, :m], oe), (form, j, ml, w, h)
        assert np.array_equal(keep[j, :m].astype(bool), oracle.fb_check(pts[j, :m], ob, os_, osb, 2.0)), (form, j)
    p.close()


@pytest.mark.parametrize("seed", range(6))
def test_preprocessing_random(ctx, seed):
    """BGR->gray, CLAHE(clip), remap with maps that leave the frame (constant-0 border), odd sizes, against the oracle"""
    from zenslam_b200.processing import processor
    rng = np.random.default_rng(300 + seed)
    w, h = int(rng.integers(64, 500)), int(rng.integers(64, 360))
    tex = _image(rng, w, h, seed % 4)
    bgr = np.stack([tex, np.roll(tex, int(rng.integers(1, 9)), 1), (255 - tex // 2)], -1).astype(np.uint8)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    k1 = float(rng.uniform(-0.3, 0.3)); sx, sy = rng.uniform(-6, 6, 2)
    r2 = ((xx - w / 2) ** 2 + (yy - h / 2) ** 2) / (w * w / 4)
    mx = (w / 2 + (xx - w / 2) * (1 + k1 * r2) + sx).astype(np.float32)
    my = (h / 2 + (yy - h / 2) * (1 + k1 * r2) + sy).astype(np.float32)
    clip = float(rng.choice([2.0, 4.0, 40.0]))
    gray = oracle.bgr2gray(bgr)
    assert np.array_equal(processor(ctx).process_image(bgr), gray)
    assert np.array_equal(processor(ctx, clahe_enabled=True, clahe_clip_limit=clip).process_image(bgr), oracle.clahe(gray, clip))
    want = oracle.remap_linear(oracle.clahe(gray, clip), mx, my)
    assert np.array_equal(processor(ctx, clahe_enabled=True, clahe_clip_limit=clip, maps=[(mx, my)]).process_image(bgr), want)
    assert np.array_equal(processor(ctx, maps=[(mx, my)]).process_image(gray), oracle.remap_linear(gray, mx, my))


@pytest.mark.parametrize("seed", range(6))
def test_corner_subpix_random(ctx, seed):
    from zenslam_b200.runtime import Pyramid, corner_subpix
    rng = np.random.default_rng(400 + seed)
    w, h = int(rng.integers(60, 400)), int(rng.integers(60, 300))
    img = _image(rng, w, h, seed % 4)
    n = int(rng.integers(1, 300))
    pts = np.stack([rng.uniform(0, w - 1, n), rng.uniform(0, h - 1, n)], 1).astype(np.float32)
    pts[: n // 5] = np.rint(pts[: n // 5])
    win = [(5, 5), (3, 3), (7, 7), (2, 6), (5, 5), (1, 1)][seed]
    its, eps = int(rng.choice([1, 5, 30])), float(rng.choice([0.001, 0.01, 0.1]))
    pyr = Pyramid(ctx, w, h, 1, (31, 31), 0)
    pyr.upload(img[None], 0); pyr.build(0, 1)
    cap = max(n, 1)
    xy = dev(ctx, pts[None].copy())
    corner_subpix(pyr, 0, 1, xy, dev(ctx, np.array([n], np.int32)), win, its, eps)
    assert cap >= n
    assert np.array_equal(xy[0, :n].cpu().numpy(), oracle.corner_subpix(img, pts, win, its, eps))
    pyr.close()


@pytest.mark.parametrize("seed", range(6))
def test_l2_random(ctx, seed):
    """ragged SIFT-like and low-entropy pairs: tensor-core kNN-2 + ratio and cross-check against the oracle"""
    from zenslam_b200.runtime import match_l2_cross, match_l2_knn2
    rng = np.random.default_rng(500 + seed)
    pairs = int(rng.integers(1, 5))
    cap_q, cap_t = int(rng.integers(1, 700)), int(rng.integers(1, 900))
    hi = [256, 256, 4, 256, 2, 64][seed]
    q = rng.integers(0, hi, (pairs, cap_q, 128)).astype(np.float32); t = rng.integers(0, hi, (pairs, cap_t, 128)).astype(np.float32)
    nq = rng.integers(0, cap_q + 1, pairs).astype(np.int32); nt = rng.integers(0, cap_t + 1, pairs).astype(np.int32)
    nq[0], nt[0] = cap_q, cap_t
    for k in range(pairs):                                 # plant exact duplicates: distance-0 ties across tiles
        if nq[k] > 2 and nt[k] > 140:
            t[k, 3] = t[k, 139] = q[k, 1]
    idx, dist, ps = [x.cpu().numpy() for x in match_l2_knn2(ctx, dev(ctx, q), dev(ctx, nq), dev(ctx, t), dev(ctx, nt), 0.8)]
    cidx, cdist = [x.cpu().numpy() for x in match_l2_cross(ctx, dev(ctx, q), dev(ctx, nq), dev(ctx, t), dev(ctx, nt))]
    for k in range(pairs):
        a, b = int(nq[k]), int(nt[k])
        if a == 0 or b == 0:
            assert np.all(idx[k, :a] == -1) and np.all(cidx[k, :a] == -1)
            continue
        oi, od = oracle.match_l2_knn2(q[k, :a], t[k, :b])
        assert np.array_equal(idx[k, :a], oi), (k, a, b)
        assert np.array_equal(dist[k, :a][oi >= 0], od[oi >= 0])
        if b >= 2:
            assert np.array_equal(np.nonzero(ps[k, :a])[0], oracle.ratio_test(oi, od, 0.8)[0])
        oq, ot, odd = oracle.match_l2_cross(q[k, :a], t[k, :b])
        keep = np.nonzero(cidx[k, :a] >= 0)[0]
        assert np.array_equal(keep, oq) and np.array_equal(cidx[k, keep], ot) and np.array_equal(cdist[k, keep], odd)


@pytest.mark.parametrize("seed", range(4))
def test_orb_detector_random(ctx, seed):
    from zenslam_b200.runtime import OrbDetector
    rng = np.random.default_rng(600 + seed)
    w, h = int(rng.integers(150, 700)), int(rng.integers(120, 500))
    kw = [dict(), dict(nfeatures=120, nlevels=4), dict(scale_factor=1.44, nlevels=5, nfeatures=900), dict(nfeatures=40)][seed]
    thr = int(rng.choice([5, 10, 20]))
    imgs = np.stack([_image(rng, w, h, (seed + k) % 4) for k in range(2)])
    mask = np.full((2, h, w), 255, np.uint8)
    mask[0, : h // 3] = 0
    mask[1][rng.random((h, w)) < 0.3] = 0
    det = OrbDetector(ctx, w, h, 2, fast_threshold=thr, **kw)
    r = det.detect_and_compute(imgs, mask)
    for k in range(2):
        o = oracle.orb_detect(imgs[k], mask[k], fast_threshold=thr, **kw)
        n = int(r["n"][k])
        assert n == len(o["x"]), (k, n, len(o["x"]))
        assert np.array_equal(r["xy"][k, :n].cpu().numpy(), np.stack([o["x"], o["y"]], 1))
        for key in ("size", "angle", "response", "octave", "desc"):
            assert np.array_equal(r[key][k, :n].cpu().numpy(), o[key]), (k, key)
    det.close()


@pytest.mark.parametrize("w,h,win,ml", [(32, 32, (31, 31), 3), (16, 16, (15, 15), 2), (8, 200, (5, 5), 3), (200, 8, (7, 7), 4),
                                        (63, 63, (31, 31), 3), (64, 64, (31, 31), 3), (4096, 16, (9, 9), 2)])
@pytest.mark.parametrize("path", ["ZS_PYR_FUSED", "ZS_PYR_SPLIT"])
def test_extreme_frame_shapes(ctx, w, h, win, ml, path, switches):
    """frames barely larger than the LK window, one-cell-high strips, very wide rows: level counts, pyramids, Scharr planes,
    grid detection, the ORB border filter and KLT on white noise against the oracle"""
    from zenslam_b200 import LK_GET_MIN_EIGENVALS
    from zenslam_b200.runtime import LK, Pyramid, fast_grid_detect, klt_track, orb_compute
    switches.set(ctx, path)
    rng = np.random.default_rng(w * 7 + h)
    img = rng.integers(0, 256, (h, w)).astype(np.uint8)
    img2 = np.roll(img, 1, 1)
    p = Pyramid(ctx, w, h, 2, win, ml)
    p.upload(np.stack([img, img2]), 0); p.build(0, 2)
    P, P2 = oracle.Pyramid(img, win, ml), oracle.Pyramid(img2, win, ml)
    assert p.levels == P.levels
    for l in range(P.levels):
        assert np.array_equal(p.image(0, l), P.image(l)) and np.array_equal(p.deriv(0, l), P.deriv(l)), l
    cell = (16, 16) if min(w, h) >= 16 else (8, 8)
    xy, resp, n = fast_grid_detect(p, 0, 2, cell, 10)
    x, y, s = oracle.grid_detect(img, cell, 10)
    assert int(n[0]) == len(x) and np.array_equal(xy[0, :len(x)].cpu().numpy(), np.stack([x, y], 1).astype(np.float32).reshape(-1, 2))
    _, _, _, on, _ = orb_compute(p, 0, 2, xy, resp, n)
    assert int(on[0]) == len(oracle.orb_compute(img, x, y)[0])
    pts = np.stack([rng.uniform(-3, w + 3, 40), rng.uniform(-3, h + 3, 40)], 1).astype(np.float32)
    lk = LK(win, ml, 99, 0.001, LK_GET_MIN_EIGENVALS, 1e-4)
    out = klt_track(p, dev(ctx, np.array([0], np.int32)), dev(ctx, np.array([1], np.int32)), dev(ctx, pts[None]),
                    dev(ctx, np.array([40], np.int32)), lk, None, 1.0)
    p1, st, err, keep = [o[0].cpu().numpy() for o in out]
    o1, os_, oe = oracle.lk_track(P, P2, pts, None, win, ml)
    assert np.array_equal(st, os_) and np.array_equal(p1, o1) and np.array_equal(err, oe)
    p.close()
