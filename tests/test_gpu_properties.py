"""GPU: size-independent properties of the hot path at the BASELINE frame sizes (752x480, 1280x1024, 3840x2160), where the
oracle is too slow to sweep every case -- what must hold whatever the data:

  * identity: LK of a frame against itself leaves every trackable point where it is, keeps every forward-backward pair, and a
    descriptor set matched against itself is the identity at distance 0;
  * translation covariance: a frame shifted by whole pixels moves the grid detector's corners and their LK tracks by exactly
    that shift, wherever the cell grid and the window see the same pixels;
  * batch independence: what a frame yields does not depend on which slot of a batch it sits in, nor on the batch size;
  * determinism: two runs of the whole batched front-end give byte-identical results (no atomics-order dependence).
"""
import numpy as np
import pytest

from zenslam_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

SIZES = [(752, 480, 16), (1280, 1024, 32), (3840, 2160, 32)]


@pytest.fixture(scope="module")
def ctx():
    from zenslam_b200.runtime import Context
    c = Context()
    yield c
    c.close()


def _detect(ctx, imgs, cell, win=(31, 31), ml=3):
    from zenslam_b200.runtime import Pyramid, fast_grid_detect, orb_compute
    n, h, w = imgs.shape
    p = Pyramid(ctx, w, h, n, win, ml)
    p.upload(np.ascontiguousarray(imgs), 0)
    p.build(0, n)
    xy, resp, cnt = fast_grid_detect(p, 0, n, (cell, cell), 10)
    oxy, oresp, _, on, desc = orb_compute(p, 0, n, xy, resp, cnt)
    return p, oxy, oresp, on, desc


@pytest.mark.parametrize("w,h,cell", SIZES)
def test_identity_tracking_and_matching(ctx, w, h, cell):
    import torch

    from zenslam_b200 import LK_GET_MIN_EIGENVALS
    from zenslam_b200.runtime import LK, klt_track, match_hamming_cross, match_hamming_knn2
    seq, _ = syn.stereo_sequence(w, h, 1, 7100 + w, subpixel=True)
    imgs = np.stack([seq[0, 0], seq[0, 0]])
    p, xy, resp, n, desc = _detect(ctx, imgs, cell)
    k = int(n[0])
    assert k > 500 and int(n[1]) == k and torch.equal(xy[0, :k], xy[1, :k]) and torch.equal(desc[0, :k], desc[1, :k])
    lk = LK((31, 31), 3, 99, 0.001, LK_GET_MIN_EIGENVALS, 1e-4)
    zero, one = ctx.to_device(np.array([0], np.int32)), ctx.to_device(np.array([1], np.int32))
    nxt, st, err, keep = klt_track(p, zero, one, xy[:1].contiguous(), n[:1].contiguous(), lk, fb_threshold=1.0)
    st = st[0, :k].cpu().numpy().astype(bool); keep = keep[0, :k].cpu().numpy().astype(bool)
    d = (nxt[0, :k] - xy[0, :k]).abs().cpu().numpy()
    assert st.mean() > 0.99 and np.array_equal(keep, st)
    assert d[st].max() <= 1e-3, d[st].max()                              # zero mismatch: the first Gauss-Newton step is zero
    idx, dist, ps = match_hamming_knn2(ctx, desc[:1].contiguous(), n[:1].contiguous(), desc[1:2].contiguous(), n[1:2].contiguous(), 0.8)
    idx, dist = idx[0, :k].cpu().numpy(), dist[0, :k].cpu().numpy()
    # a descriptor's nearest neighbour in its own set is itself (or an identical earlier row: ties -> smaller index) at distance 0
    assert (dist[:, 0] == 0).all() and (idx[:, 0] <= np.arange(k)).all()
    d_np = desc[0, :k].cpu().numpy()
    assert np.array_equal(d_np[idx[:, 0]], d_np)
    cidx, cdist = match_hamming_cross(ctx, desc[:1].contiguous(), n[:1].contiguous(), desc[1:2].contiguous(), n[1:2].contiguous())
    cidx = cidx[0, :k].cpu().numpy()
    kept = cidx >= 0
    assert kept.mean() > 0.99 and np.array_equal(d_np[cidx[kept]], d_np[kept]) and (cdist[0, :k].cpu().numpy()[kept] == 0).all()


@pytest.mark.parametrize("w,h,cell", SIZES[:2])
def test_translation_covariance(ctx, w, h, cell):
    """the same texture shifted by (2 cells, 1 cell): corners of cells that exist in both frames move by exactly the shift,
    and so do their LK tracks into a third frame shifted the same way (windows far enough from the border)"""
    import torch

    from zenslam_b200 import LK_GET_MIN_EIGENVALS
    from zenslam_b200.runtime import LK, klt_track
    sx, sy = 2 * cell, cell
    big, _ = syn.stereo_sequence(w + sx, h + sy, 2, 7300 + w, subpixel=True)
    A0, B0 = big[0, 0][sy:, sx:], big[1, 0][sy:, sx:]              # the un-shifted view
    A1, B1 = big[0, 0][:h, :w], big[1, 0][:h, :w]                  # everything moved by (+sx, +sy)
    p, xy, resp, n, desc = _detect(ctx, np.stack([A0, B0, A1, B1]).copy(), cell)
    k0, k1 = int(n[0]), int(n[2])
    a = {(float(x), float(y)): float(r) for (x, y), r in zip(xy[0, :k0].cpu().numpy(), resp[0, :k0].cpu().numpy())}
    b = {(float(x) - sx, float(y) - sy): float(r) for (x, y), r in zip(xy[2, :k1].cpu().numpy(), resp[2, :k1].cpu().numpy())}
    common = [q for q in a if q in b]
    assert len(common) > 0.8 * min(k0, k1) and all(a[q] == b[q] for q in common)
    # LK: A0 -> B0 from the common corners vs A1 -> B1 from the shifted corners, away from the borders
    m = 80
    pts = np.array([q for q in common if m <= q[0] < w - sx - m and m <= q[1] < h - sy - m], np.float32)
    assert len(pts) > 100
    cap = len(pts)
    both = np.stack([pts, pts + np.array([sx, sy], np.float32)])
    prev_slot, next_slot = ctx.to_device(np.array([0, 2], np.int32)), ctx.to_device(np.array([1, 3], np.int32))
    cnt = ctx.to_device(np.array([cap, cap], np.int32))
    shift = np.array([sx, sy], np.float32)
    # one level: every window sees exactly the same pixels in both views -> same statuses, same flow (the coordinates differ
    # by whole pixels, so only the float rounding of `position + flow` can differ)
    nxt, st, err = klt_track(p, prev_slot, next_slot, ctx.to_device(both), cnt, LK((31, 31), 0, 99, 0.001, LK_GET_MIN_EIGENVALS, 1e-4))
    nxt, st, err = nxt.cpu().numpy(), st.cpu().numpy(), err.cpu().numpy()
    assert np.array_equal(st[0], st[1]) and st[0].mean() > 0.9 and np.array_equal(err[0], err[1])
    ok = st[0].astype(bool)
    d1 = np.abs((nxt[1] - shift) - nxt[0])[ok].max(axis=1)             # float steps at coordinates of different magnitude can
    assert np.median(d1) <= 1e-4 and d1.max() <= 0.01, (np.median(d1), d1.max())   # end a slow track one iteration apart
    # four levels: the coarse levels of the two views differ near the image border (REFLECT_101 sits elsewhere), which only
    # changes the initial guess handed down -- the tracks end at the same optimum
    nxt, st, err = klt_track(p, prev_slot, next_slot, ctx.to_device(both), cnt, LK((31, 31), 3, 99, 0.001, LK_GET_MIN_EIGENVALS, 1e-4))
    nxt, st = nxt.cpu().numpy(), st.cpu().numpy()
    assert (st[0] == st[1]).mean() > 0.99
    ok = (st[0] & st[1]).astype(bool)
    d = np.abs((nxt[1] - shift) - nxt[0])[ok].max(axis=1)
    assert np.median(d) <= 1e-3 and np.quantile(d, 0.99) <= 0.05, (np.median(d), np.quantile(d, 0.99))


def test_batch_slot_independence_and_determinism(ctx):
    """zs_frontend: frame k of a batch of 6 gives what it gives as a batch of 1 (given the same previous frame), and two runs
    of the same batch are byte-identical"""
    from zenslam_b200 import detection_options, slam_options, tracking_options
    from zenslam_b200.frontend import StereoFrontend
    w, h, B = 752, 480, 6
    opts = slam_options(matcher="KNN", detection=detection_options(), tracking=tracking_options())
    seq, _ = syn.stereo_sequence(w, h, B, 7500, subpixel=True)
    L, R = np.ascontiguousarray(seq[:, 0]), np.ascontiguousarray(seq[:, 1])
    fe = StereoFrontend(ctx, w, h, B, opts)
    r1 = {k: np.array(v, copy=True) for k, v in fe.process(L, R).items()}
    fe.close()
    fe = StereoFrontend(ctx, w, h, B, opts)
    r2 = fe.process(L, R)
    for k in r1:
        assert np.array_equal(r1[k], np.asarray(r2[k])), k
    fe.close()
    one = StereoFrontend(ctx, w, h, 1, opts)
    for k in range(B):
        r = one.process(L[k:k + 1], R[k:k + 1])
        n = int(r["n_left"][0])
        assert n == int(r1["n_left"][k])
        assert np.array_equal(r["kp_left"][0, :n], r1["kp_left"][k, :n]) and np.array_equal(r["desc_left"][0, :n], r1["desc_left"][k, :n])
        assert np.array_equal(r["match_idx"][0, :n], r1["match_idx"][k, :n])
        if k > 0:                                                    # temporal jobs start from frame k-1's keypoints, carried over
            for job in range(r["track_pts"].shape[0]):
                m = int(r["track_n"][job, 0])
                assert m == int(r1["track_n"][job, k])
                assert np.array_equal(r["track_pts"][job, 0, :m], r1["track_pts"][job, k, :m]), (k, job)
                assert np.array_equal(r["track_keep"][job, 0, :m], r1["track_keep"][job, k, :m]), (k, job)
    one.close()
