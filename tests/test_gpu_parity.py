"""GPU parity: the CUDA path (through the C ABI) against the oracle and the committed cv2 golden vectors.
Bit-exact for pyramids, keypoints, descriptors, match indices/distances; KLT within 0.01 px with identical status."""
import numpy as np
import pytest

import oracle
from zenslam_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

LK_TOL = 0.01


@pytest.fixture(scope="module")
def ctx():
    from zenslam_b200.runtime import Context
    c = Context()
    yield c
    c.close()


def dev(ctx, a, dtype=None):
    import torch
    return ctx.to_device(np.ascontiguousarray(a), dtype)


# ---------------------------------------------------------------------------------------------
# pyramid
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("w,h,win,ml", [(256, 192, (15, 15), 3), (752, 480, (31, 31), 3), (752, 480, (63, 63), 4),
                                        (321, 243, (21, 21), 5), (1280, 720, (31, 21), 3)])
@pytest.mark.parametrize("path", ["ZS_PYR_FUSED", "ZS_PYR_SPLIT"])
def test_pyramid_vs_oracle(ctx, w, h, win, ml, path, switches):
    """both builders -- one fused launch per level (few images) and pad / pyrDown / Scharr as separate passes (many)"""
    from zenslam_b200.runtime import Pyramid
    switches.set(ctx, path)
    imgs = np.stack([syn.stereo_pair(w, h, 10 + w)[0], syn.stereo_pair(w, h, 11 + w)[1],
                     np.random.default_rng(w).integers(0, 256, (h, w), dtype=np.uint8)])
    p = Pyramid(ctx, w, h, 3, win, ml)
    p.upload(imgs, 0)
    p.build(0, 3)
    for s in range(3):
        P = oracle.Pyramid(imgs[s], win, ml)
        assert p.levels == P.levels
        for l in range(P.levels):
            assert p.level_size(l) == P.level_size(l)
            assert np.array_equal(p.image(s, l), P.image(l)), (s, l)
            assert np.array_equal(p.deriv(s, l), P.deriv(l)), (s, l)


@pytest.mark.parametrize("path", ["ZS_PYR_FUSED", "ZS_PYR_SPLIT"])
def test_pyramid_every_width_residue(ctx, path, switches):
    """the fused builder patches the out-of-image columns of a row's first / last 4-pixel item in registers: every
    residue of the width mod 8 (and both parities of the height) at every level, on white noise"""
    from zenslam_b200.runtime import Pyramid
    switches.set(ctx, path)
    for w in range(64, 81):
        h = 40 + (w & 1) + (w >> 2 & 1) * 2
        img = np.random.default_rng(w).integers(0, 256, (2, h, w), dtype=np.uint8)
        p = Pyramid(ctx, w, h, 2, (9, 9), 2)
        p.upload(img, 0)
        p.build(0, 2)
        for s in range(2):
            P = oracle.Pyramid(img[s], (9, 9), 2)
            assert p.levels == P.levels
            for l in range(P.levels):
                assert np.array_equal(p.image(s, l), P.image(l)), (w, h, s, l)
                assert np.array_equal(p.deriv(s, l), P.deriv(l)), (w, h, s, l)


def test_pyramid_golden(ctx, golden):
    from zenslam_b200.runtime import Pyramid
    g = golden("pyramid")
    L = g["L"]
    h, w = L.shape
    p = Pyramid(ctx, w, h, 2, (15, 15), 3)
    p.upload(L, 1)        # slot 1: exercises the slot offset
    p.build(1, 1)
    assert p.levels == int(g["cv_levels_w15"])
    for l in range(p.levels):
        assert np.array_equal(p.image(1, l), g[f"cv_pyr_img{l}"])
        assert np.array_equal(p.deriv(1, l), g[f"cv_pyr_der{l}"])


# ---------------------------------------------------------------------------------------------
# detection
# ---------------------------------------------------------------------------------------------
def _grid(ctx, imgs, cell, thr, occ=None):
    from zenslam_b200.runtime import Pyramid, fast_grid_detect
    n, h, w = imgs.shape
    p = Pyramid(ctx, w, h, n, (16, 16), 0)
    p.upload(imgs, 0)
    p.build(0, n)
    xy, resp, cnt = fast_grid_detect(p, 0, n, cell, thr, occ)
    return p, xy, resp, cnt


@pytest.mark.parametrize("name", ["L", "Lq"])
@pytest.mark.parametrize("cell,thr", [((16, 16), 10), ((32, 32), 10), ((64, 64), 1), ((24, 16), 5)])
def test_grid_detect_golden(ctx, golden, name, cell, thr):
    g = golden("detect")
    _, xy, resp, cnt = _grid(ctx, g[name][None], cell, thr)
    k = f"cv_grid_{name}_c{cell[0]}x{cell[1]}_t{thr}"
    n = int(cnt[0])
    assert n == len(g[k + "_x"])
    xy = xy[0, :n].cpu().numpy()
    assert np.array_equal(xy[:, 0], g[k + "_x"]) and np.array_equal(xy[:, 1], g[k + "_y"])
    assert np.array_equal(resp[0, :n].cpu().numpy(), g[k + "_r"])


def test_grid_detect_occupancy_golden(ctx, golden):
    g = golden("detect")
    _, xy, resp, cnt = _grid(ctx, g["L"][None], (16, 16), 10, g["occ"][None])
    n = int(cnt[0])
    xy = xy[0, :n].cpu().numpy()
    assert np.array_equal(xy[:, 0], g["cv_grid_occ_x"]) and np.array_equal(xy[:, 1], g["cv_grid_occ_y"])
    assert np.array_equal(resp[0, :n].cpu().numpy(), g["cv_grid_occ_r"])


@pytest.mark.parametrize("w,h,cell,thr", [(752, 480, (16, 16), 10), (752, 480, (16, 16), 1), (1280, 1024, (32, 32), 10),
                                          (752, 480, (64, 64), 1), (640, 400, (20, 28), 7)])
def test_grid_detect_batch_vs_oracle(ctx, w, h, cell, thr):
    seq, _ = syn.stereo_sequence(w, h, 3, 2000 + w, subpixel=True)
    imgs = seq.reshape(-1, h, w)
    imgs[1] = imgs[1] // 8 * 8           # heavy ties
    _, xy, resp, cnt = _grid(ctx, imgs, cell, thr)
    for i in range(len(imgs)):
        ox, oy, osc = oracle.grid_detect(imgs[i], cell, thr)
        n = int(cnt[i])
        assert n == len(ox)
        a = xy[i, :n].cpu().numpy()
        assert np.array_equal(a[:, 0], ox) and np.array_equal(a[:, 1], oy)
        assert np.array_equal(resp[i, :n].cpu().numpy(), osc)


@pytest.mark.parametrize("thr", [1, 10, 40])
def test_fast_full_frame_golden(ctx, golden, thr):
    from zenslam_b200.runtime import Pyramid, fast_detect
    g = golden("detect")
    L = g["L"]
    h, w = L.shape
    p = Pyramid(ctx, w, h, 1, (16, 16), 0)
    p.upload(L, 0); p.build(0, 1)
    xy, resp, cnt = fast_detect(p, 0, 1, thr, None, 65536)
    n = int(cnt[0])
    assert n == len(g[f"cv_fast_t{thr}_x"])
    a = xy[0, :n].cpu().numpy()
    assert np.array_equal(a[:, 0], g[f"cv_fast_t{thr}_x"]) and np.array_equal(a[:, 1], g[f"cv_fast_t{thr}_y"])
    assert np.array_equal(resp[0, :n].cpu().numpy(), g[f"cv_fast_t{thr}_r"])


# ---------------------------------------------------------------------------------------------
# ORB
# ---------------------------------------------------------------------------------------------
def test_orb_golden(ctx, golden):
    from zenslam_b200.runtime import orb_compute
    g = golden("detect")
    L = g["L"]
    p, xy, resp, cnt = _grid(ctx, L[None], (16, 16), 10)
    oxy, oresp, src, on, desc = orb_compute(p, 0, 1, xy, resp, cnt)
    assert np.array_equal(p.blur(0), g["cv_orb_blur"])
    n = int(on[0])
    assert n == len(g["cv_orb_kx"])
    a = oxy[0, :n].cpu().numpy()
    assert np.array_equal(a[:, 0], g["cv_orb_kx"]) and np.array_equal(a[:, 1], g["cv_orb_ky"])
    assert np.array_equal(desc[0, :n].cpu().numpy(), g["cv_orb_desc"])


def test_orb_rotated_subpixel_golden(ctx, golden):
    import torch
    from zenslam_b200.runtime import Pyramid, orb_compute
    g = golden("detect")
    L = g["L"]
    h, w = L.shape
    p = Pyramid(ctx, w, h, 1, (16, 16), 0)
    p.upload(L, 0); p.build(0, 1)
    xy = dev(ctx, np.stack([g["orb_in_x"], g["orb_in_y"]], 1)[None])
    n = dev(ctx, np.array([len(g["orb_in_x"])], np.int32))
    resp = dev(ctx, np.zeros((1, len(g["orb_in_x"])), np.float32))
    ang = dev(ctx, g["orb_in_a"][None])
    oxy, _, src, on, desc = orb_compute(p, 0, 1, xy, resp, n, ang)
    k = int(on[0])
    assert k == len(g["cv_orb_rot_kx"])
    a = oxy[0, :k].cpu().numpy()
    assert np.array_equal(a[:, 0], g["cv_orb_rot_kx"]) and np.array_equal(a[:, 1], g["cv_orb_rot_ky"])
    assert np.array_equal(desc[0, :k].cpu().numpy(), g["cv_orb_rot_desc"])


@pytest.mark.parametrize("w,h,cell", [(752, 480, (16, 16)), (1280, 720, (32, 32))])
def test_orb_batch_vs_oracle(ctx, w, h, cell):
    from zenslam_b200.runtime import orb_compute
    seq, _ = syn.stereo_sequence(w, h, 2, 3000 + w)
    imgs = seq.reshape(-1, h, w)
    p, xy, resp, cnt = _grid(ctx, imgs, cell, 10)
    oxy, oresp, src, on, desc = orb_compute(p, 0, len(imgs), xy, resp, cnt)
    for i in range(len(imgs)):
        assert np.array_equal(p.blur(i), oracle.orb_blur(imgs[i]))
        ox, oy, osc = oracle.grid_detect(imgs[i], cell, 10)
        kept, odesc = oracle.orb_compute(imgs[i], ox, oy)
        n = int(on[i])
        assert n == len(kept)
        a = oxy[i, :n].cpu().numpy()
        assert np.array_equal(a[:, 0], ox[kept]) and np.array_equal(a[:, 1], oy[kept])
        assert np.array_equal(src[i, :n].cpu().numpy(), kept)
        assert np.array_equal(oresp[i, :n].cpu().numpy(), osc[kept])
        assert np.array_equal(desc[i, :n].cpu().numpy(), odesc)


# ---------------------------------------------------------------------------------------------
# matching
# ---------------------------------------------------------------------------------------------
def _pad(descs, cap, dtype):
    out = np.zeros((len(descs), cap) + descs[0].shape[1:], dtype)
    for i, d in enumerate(descs):
        out[i, :len(d)] = d
    return out


@pytest.mark.parametrize("nm,qk,tk", [("orb", "dl", "dr"), ("b16", "q16", "t16"), ("one", "dl", "dr")])
def test_hamming_golden(ctx, golden, nm, qk, tk):
    from zenslam_b200.runtime import match_hamming_cross, match_hamming_knn2
    g = golden("match")
    q, t = g[qk], g[tk]
    if nm == "one":
        t = t[:1]
    dq, dt = dev(ctx, q[None]), dev(ctx, t[None])
    nq, nt = dev(ctx, np.array([len(q)], np.int32)), dev(ctx, np.array([len(t)], np.int32))
    idx, dist, ps = match_hamming_knn2(ctx, dq, nq, dt, nt, 0.8)
    idx, dist, ps = idx[0].cpu().numpy(), dist[0].cpu().numpy(), ps[0].cpu().numpy()
    assert np.array_equal(idx, g[f"cv_knn_{nm}_idx"])
    valid = idx >= 0
    assert np.array_equal(dist[valid], g[f"cv_knn_{nm}_dist"][valid])
    assert np.array_equal(np.nonzero(ps)[0], g[f"cv_ratio_{nm}_q"])
    cidx, cdist = match_hamming_cross(ctx, dq, nq, dt, nt)
    cidx, cdist = cidx[0].cpu().numpy(), cdist[0].cpu().numpy()
    keep = np.nonzero(cidx >= 0)[0]
    assert np.array_equal(keep, g[f"cv_cross_{nm}_q"]) and np.array_equal(cidx[keep], g[f"cv_cross_{nm}_t"])
    assert np.array_equal(cdist[keep], g[f"cv_cross_{nm}_d"])


@pytest.mark.parametrize("path", ["cuda_core", "tensor"])
def test_hamming_batch_ragged_vs_oracle(ctx, switches, path):
    """ragged batch (empty sides, one row, tile tails, duplicates, 16-bit descriptors) through both Hamming paths: the CUDA-core
    kernel and the tcgen05 kind::i8 kernel fed with the descriptors expanded to one byte per bit (forced here: by default
    only calls of at least 12 M distances take it)"""
    from zenslam_b200.runtime import match_hamming_cross, match_hamming_knn2
    if path == "tensor":
        switches.set(ctx, "ZS_HAMMING_TENSOR_MIN", 1)
    else:
        switches.set(ctx, "ZS_HAMMING_NO_TENSOR")
    rng = np.random.default_rng(5)
    sizes = [(1100, 1000), (0, 50), (37, 0), (1, 1), (300, 1411), (129, 128)]
    qs = [rng.integers(0, 256, (a, 32), dtype=np.uint8) for a, _ in sizes]
    ts = [rng.integers(0, 256, (b, 32), dtype=np.uint8) for _, b in sizes]
    ts[0][:40] = qs[0][100:140]; ts[0][500] = ts[0][3]           # duplicates -> ties
    qs[4][:, 2:] = 0; ts[4][:, 2:] = 0                           # 16-bit descriptors: massive ties
    cap_q, cap_t = 1100, 1411
    dq, dt = dev(ctx, _pad(qs, cap_q, np.uint8)), dev(ctx, _pad(ts, cap_t, np.uint8))
    nq = dev(ctx, np.array([a for a, _ in sizes], np.int32)); nt = dev(ctx, np.array([b for _, b in sizes], np.int32))
    idx, dist, ps = match_hamming_knn2(ctx, dq, nq, dt, nt, 0.8)
    cidx, cdist = match_hamming_cross(ctx, dq, nq, dt, nt)
    idx, dist, ps, cidx, cdist = [a.cpu().numpy() for a in (idx, dist, ps, cidx, cdist)]
    for k, (a, b) in enumerate(sizes):
        oi, od = oracle.match_hamming_knn2(qs[k], ts[k])
        assert np.array_equal(idx[k, :a], oi)
        assert np.array_equal(dist[k, :a][oi >= 0], od.astype(np.float32)[oi >= 0])
        rq, _, _ = oracle.ratio_test(oi, od.astype(np.float32), 0.8)
        assert np.array_equal(np.nonzero(ps[k, :a])[0], rq)
        assert not ps[k, a:].any() and (idx[k, a:] == -1).all()
        oq, ot, odd = oracle.match_hamming_cross(qs[k], ts[k])
        keep = np.nonzero(cidx[k, :a] >= 0)[0]
        assert np.array_equal(keep, oq) and np.array_equal(cidx[k, keep], ot)
        assert np.array_equal(cdist[k, keep], odd.astype(np.float32))


@pytest.mark.parametrize("variant,splits", [(0, 0), (0, 3), (2, 0), (3, 2), (4, 0), (5, 0), (5, 5), (-1, 0)])
def test_hamming_kernel_variants_vs_oracle(ctx, monkeypatch, variant, splits):
    """every instantiation of k_hamming_top2 behind ZS_HAMMING_VARIANT (queries per thread x carry-save / plain popcount) and
    train-side split counts, packed (distance << 22 | row) keys included, and (variant -1) the tensor-core path: ragged batch with duplicate rows and 16-bit
    descriptors (tie floods) against the oracle's stable top-2 / cross-check"""
    from zenslam_b200.runtime import match_hamming_cross, match_hamming_knn2
    if variant > 0:
        monkeypatch.setenv("ZS_HAMMING_VARIANT", str(variant))
    if variant == -1:
        monkeypatch.setenv("ZS_HAMMING_TENSOR_MIN", "1")             # the tensor-core path (bits expanded to bytes), whatever the size
    else:
        monkeypatch.setenv("ZS_HAMMING_NO_TENSOR", "1")
    if splits:
        monkeypatch.setenv("ZS_HAMMING_SPLITS", str(splits))
    ctx.reload_switches()
    try:
        rng = np.random.default_rng(50 + abs(variant))
        sizes = [(700, 900), (0, 50), (37, 0), (1, 1), (300, 1411), (129, 128), (513, 2)]
        qs = [rng.integers(0, 256, (a, 32), dtype=np.uint8) for a, _ in sizes]
        ts = [rng.integers(0, 256, (b, 32), dtype=np.uint8) for _, b in sizes]
        ts[0][:40] = qs[0][100:140]; ts[0][500] = ts[0][3]; ts[0][899] = ts[0][3]
        qs[4][:, 2:] = 0; ts[4][:, 2:] = 0
        ts[6][1] = ts[6][0]                                          # two identical train rows: best and second tie on every query
        cap_q, cap_t = 700, 1411
        dq, dt = dev(ctx, _pad(qs, cap_q, np.uint8)), dev(ctx, _pad(ts, cap_t, np.uint8))
        nq = dev(ctx, np.array([a for a, _ in sizes], np.int32)); nt = dev(ctx, np.array([b for _, b in sizes], np.int32))
        idx, dist, ps = match_hamming_knn2(ctx, dq, nq, dt, nt, 0.8)
        cidx, cdist = match_hamming_cross(ctx, dq, nq, dt, nt)
        idx, dist, ps, cidx, cdist = [a.cpu().numpy() for a in (idx, dist, ps, cidx, cdist)]
        for k, (a, b) in enumerate(sizes):
            oi, od = oracle.match_hamming_knn2(qs[k], ts[k])
            assert np.array_equal(idx[k, :a], oi), (variant, splits, k)
            assert np.array_equal(dist[k, :a][oi >= 0], od.astype(np.float32)[oi >= 0])
            rq, _, _ = oracle.ratio_test(oi, od.astype(np.float32), 0.8)
            assert np.array_equal(np.nonzero(ps[k, :a])[0], rq)
            oq, ot, odd = oracle.match_hamming_cross(qs[k], ts[k])
            keep = np.nonzero(cidx[k, :a] >= 0)[0]
            assert np.array_equal(keep, oq) and np.array_equal(cidx[k, keep], ot)
            assert np.array_equal(cdist[k, keep], odd.astype(np.float32))
    finally:
        monkeypatch.delenv("ZS_HAMMING_VARIANT", raising=False)
        monkeypatch.delenv("ZS_HAMMING_SPLITS", raising=False)
        monkeypatch.delenv("ZS_HAMMING_TENSOR_MIN", raising=False)
        monkeypatch.delenv("ZS_HAMMING_NO_TENSOR", raising=False)
        ctx.reload_switches()


def test_l2_golden_sift(ctx, golden):
    from zenslam_b200.runtime import match_l2_cross, match_l2_knn2
    g = golden("match")
    q, t = g["sift0"].astype(np.float32), g["sift1"].astype(np.float32)
    dq, dt = dev(ctx, q[None]), dev(ctx, t[None])
    nq, nt = dev(ctx, np.array([len(q)], np.int32)), dev(ctx, np.array([len(t)], np.int32))
    idx, dist, ps = match_l2_knn2(ctx, dq, nq, dt, nt, 0.8)
    idx, dist, ps = idx[0].cpu().numpy(), dist[0].cpu().numpy(), ps[0].cpu().numpy()
    assert np.array_equal(idx, g["cv_knn_sift_idx"])
    assert np.array_equal(dist, g["cv_knn_sift_dist"])
    assert np.array_equal(np.nonzero(ps)[0], g["cv_ratio_sift_q"])
    cidx, cdist = match_l2_cross(ctx, dq, nq, dt, nt)
    cidx, cdist = cidx[0].cpu().numpy(), cdist[0].cpu().numpy()
    keep = np.nonzero(cidx >= 0)[0]
    assert np.array_equal(keep, g["cv_cross_sift_q"]) and np.array_equal(cidx[keep], g["cv_cross_sift_t"])
    assert np.array_equal(cdist[keep], g["cv_cross_sift_d"])


def test_l2_rejects_non_integer(ctx):
    from zenslam_b200 import ZenslamCudaError
    from zenslam_b200.runtime import match_l2_cross, match_l2_knn2
    q = np.random.default_rng(0).random((1, 8, 128), dtype=np.float32)
    n = dev(ctx, np.array([8], np.int32))
    # the device entry never waits for the GPU: it reports through the data (no neighbour anywhere) ...
    idx, dist, ps = match_l2_knn2(ctx, dev(ctx, q), n, dev(ctx, q), n, 0.8)
    assert (idx.cpu().numpy() == -1).all() and not ps.cpu().numpy().any()
    cidx, _ = match_l2_cross(ctx, dev(ctx, q), n, dev(ctx, q), n)
    assert (cidx.cpu().numpy() == -1).all()
    # a later call with valid rows is not affected by the earlier bad one ...
    good = np.rint(q * 200).astype(np.float32)
    gidx, _, _ = match_l2_knn2(ctx, dev(ctx, good), n, dev(ctx, good), n, 0.8)
    assert np.array_equal(gidx.cpu().numpy()[0, :, 0], np.arange(8))
    # ... and the bad one is still reported through the context's asynchronous error, once
    with pytest.raises(ZenslamCudaError):
        ctx.async_error()
    ctx.async_error()
    # the host entries (what the C++ bf_matcher calls) return the status themselves
    from zenslam_b200 import _lib
    import ctypes as C
    out_i = np.zeros((8, 2), np.int32); out_d = np.zeros((8, 2), np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert _lib.lib().zs_knn_match_host(ctx._h, p(q[0]), 8, p(q[0]), 8, 128, 1, 2, 0, p(out_i), p(out_d)) == -5
    ok = np.rint(q[0] * 200).astype(np.float32)
    assert _lib.lib().zs_knn_match_host(ctx._h, p(ok), 8, p(ok), 8, 128, 1, 2, 0, p(out_i), p(out_d)) == 0
    assert np.array_equal(out_i[:, 0], np.arange(8))


@pytest.mark.parametrize("nq,nt,dim", [(2000, 1700, 128), (300, 517, 128), (90, 75, 64)])
def test_l2_u8_rows_equal_float_rows(ctx, nq, nt, dim):
    """zs_match_l2_*_u8 (no conversion pass) == the float entries on the same values == the oracle; also through the host entry
    with norm 2"""
    import ctypes as C

    from zenslam_b200 import _lib
    from zenslam_b200.runtime import match_l2_cross, match_l2_knn2
    rng = np.random.default_rng(nq + dim)
    q = rng.integers(0, 256, (nq, dim)).astype(np.uint8); t = rng.integers(0, 256, (nt, dim)).astype(np.uint8)
    t[:40] = q[10:50]; t[41] = t[3]
    dnq, dnt = dev(ctx, np.array([nq], np.int32)), dev(ctx, np.array([nt], np.int32))
    oi, od = oracle.match_l2_knn2(q.astype(np.float32), t.astype(np.float32))
    for cast in (np.uint8, np.float32):
        idx, dist, ps = match_l2_knn2(ctx, dev(ctx, q[None].astype(cast)), dnq, dev(ctx, t[None].astype(cast)), dnt, 0.8)
        assert np.array_equal(idx[0].cpu().numpy(), oi) and np.array_equal(dist[0].cpu().numpy(), od), cast
        cidx, cdist = match_l2_cross(ctx, dev(ctx, q[None].astype(cast)), dnq, dev(ctx, t[None].astype(cast)), dnt)
        oq, ot, odd = oracle.match_l2_cross(q.astype(np.float32), t.astype(np.float32))
        keep = np.nonzero(cidx[0].cpu().numpy() >= 0)[0]
        assert np.array_equal(keep, oq) and np.array_equal(cidx[0].cpu().numpy()[keep], ot) and np.array_equal(cdist[0].cpu().numpy()[keep], odd)
    ctx.async_error()
    out_i = np.zeros((nq, 2), np.int32); out_d = np.zeros((nq, 2), np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert _lib.lib().zs_knn_match_host(ctx._h, p(q), nq, p(t), nt, dim, 2, 2, 0, p(out_i), p(out_d)) == 0
    assert np.array_equal(out_i, oi) and np.array_equal(out_d, od)


@pytest.mark.parametrize("nq,nt", [(2000, 2000), (517, 1033), (1, 300), (256, 128)])
def test_l2_sift_like_vs_oracle(ctx, nq, nt):
    from zenslam_b200.runtime import match_l2_cross, match_l2_knn2
    rng = np.random.default_rng(nq)
    # SIFT-like: integer-valued, row norm ~512, clipped at 255; planted duplicates for ties
    def mk(n):
        a = rng.gamma(0.6, 40.0, (n, 128))
        a = a / np.linalg.norm(a, axis=1, keepdims=True) * 512
        return np.clip(np.rint(a), 0, 255).astype(np.float32)
    q, t = mk(nq), mk(nt)
    if nt > 40 and nq > 10:
        t[7] = q[3]; t[31] = q[3]; t[5] = t[6]
    dq, dt = dev(ctx, q[None]), dev(ctx, t[None])
    dnq, dnt = dev(ctx, np.array([nq], np.int32)), dev(ctx, np.array([nt], np.int32))
    idx, dist, ps = match_l2_knn2(ctx, dq, dnq, dt, dnt, 0.8)
    oi, od = oracle.match_l2_knn2(q, t)
    assert np.array_equal(idx[0].cpu().numpy(), oi)
    assert np.array_equal(dist[0].cpu().numpy()[oi >= 0], od[oi >= 0])
    cidx, cdist = match_l2_cross(ctx, dq, dnq, dt, dnt)
    cidx, cdist = cidx[0].cpu().numpy(), cdist[0].cpu().numpy()
    oq, ot, odd = oracle.match_l2_cross(q, t)
    keep = np.nonzero(cidx >= 0)[0]
    assert np.array_equal(keep, oq) and np.array_equal(cidx[keep], ot) and np.array_equal(cdist[keep], odd)


# ---------------------------------------------------------------------------------------------
# KLT
# ---------------------------------------------------------------------------------------------
LK_CASES = [((15, 15), 3), ((21, 21), 2), ((31, 31), 3), ((31, 31), 0), ((63, 63), 3), ((31, 21), 3)]


def _klt(ctx, A, B, pts, init, win, ml, fb=None):
    from zenslam_b200 import LK_GET_MIN_EIGENVALS, LK_USE_INITIAL_FLOW
    from zenslam_b200.runtime import LK, Pyramid, klt_track
    h, w = A.shape
    p = Pyramid(ctx, w, h, 2, win, ml)
    p.upload(np.stack([A, B]), 0); p.build(0, 2)
    n = len(pts)
    lk = LK(win, ml, 99, 0.001, LK_GET_MIN_EIGENVALS | (LK_USE_INITIAL_FLOW if init is not None else 0), 1e-4)
    out = klt_track(p, dev(ctx, np.array([0], np.int32)), dev(ctx, np.array([1], np.int32)), dev(ctx, pts[None]),
                    dev(ctx, np.array([n], np.int32)), lk, None if init is None else dev(ctx, init[None].copy()), fb)
    return [o[0].cpu().numpy() for o in out]


@pytest.mark.parametrize("win,ml", LK_CASES)
@pytest.mark.parametrize("init", [False, True])
def test_klt_golden(ctx, golden, win, ml, init):
    g = golden("klt")
    k = f"{'cv_lki' if init else 'cv_lk'}_w{win[0]}x{win[1]}_l{ml}"
    p1, st, err = _klt(ctx, g["A"], g["B"], g["pts"], g["init"] if init else None, win, ml)
    assert np.array_equal(st, g[k + "_st"])
    ok = st > 0
    assert np.abs(p1 - g[k + "_p1"])[ok].max() < LK_TOL
    assert np.allclose(err, g[k + "_err"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("win,ml", LK_CASES)
@pytest.mark.parametrize("init", [False, True])
def test_klt_bit_exact_vs_oracle(ctx, golden, win, ml, init):
    """The CUDA kernel and the C oracle implement the same exact-integer arithmetic: identical bits."""
    g = golden("klt")
    PA, PB = oracle.Pyramid(g["A"], win, ml), oracle.Pyramid(g["B"], win, ml)
    flags = oracle.LK_GET_MIN_EIGENVALS | (oracle.LK_USE_INITIAL_FLOW if init else 0)
    o1, os_, oe = oracle.lk_track(PA, PB, g["pts"], g["init"] if init else None, win, ml, flags=flags)
    p1, st, err = _klt(ctx, g["A"], g["B"], g["pts"], g["init"] if init else None, win, ml)
    assert np.array_equal(st, os_)
    assert np.array_equal(p1, o1)
    assert np.array_equal(err, oe)


def test_klt_fb_gate_golden(ctx, golden):
    g = golden("klt")
    p1, st, err, keep = _klt(ctx, g["A"], g["B"], g["pts"], None, (31, 31), 3, fb=1.0)
    assert np.array_equal(keep.astype(bool), g["cv_fb_keep"])
    assert np.abs(p1 - g["cv_fb_p1"])[g["cv_fb_keep"]].max() < LK_TOL


def test_klt_multi_job_752(ctx):
    """C2-shaped: 4 jobs over a 4-slot pyramid (temporal L/R, stereo both ways), ragged counts."""
    from zenslam_b200 import LK_GET_MIN_EIGENVALS
    from zenslam_b200.runtime import LK, Pyramid, klt_track
    w, h, win, ml = 752, 480, (31, 31), 3
    seq, _ = syn.stereo_sequence(w, h, 2, 4000, subpixel=True)
    imgs = seq.reshape(4, h, w)          # slots: 0 = L0, 1 = R0, 2 = L1, 3 = R1
    p = Pyramid(ctx, w, h, 4, win, ml)
    p.upload(imgs, 0); p.build(0, 4)
    rng = np.random.default_rng(1)
    cap = 1200
    counts = np.array([1200, 800, 0, 37], np.int32)
    pts = np.stack([rng.uniform(0, w, (4, cap)), rng.uniform(0, h, (4, cap))], -1).astype(np.float32)
    prev = np.array([0, 1, 2, 3], np.int32); nxt = np.array([2, 3, 3, 2], np.int32)
    lk = LK(win, ml, 99, 0.001, LK_GET_MIN_EIGENVALS, 1e-4)
    p1, st, err, keep = klt_track(p, dev(ctx, prev), dev(ctx, nxt), dev(ctx, pts), dev(ctx, counts), lk, None, 1.0)
    p1, st, err, keep = [a.cpu().numpy() for a in (p1, st, err, keep)]
    pyr = [oracle.Pyramid(imgs[i], win, ml) for i in range(4)]
    for j in range(4):
        n = counts[j]
        o1, os_, oe = oracle.lk_track(pyr[prev[j]], pyr[nxt[j]], pts[j, :n], None, win, ml)
        ob, osb, _ = oracle.lk_track(pyr[nxt[j]], pyr[prev[j]], o1, None, win, ml)
        okeep = oracle.fb_check(pts[j, :n], ob, os_, osb, 1.0)
        assert np.array_equal(st[j, :n], os_) and np.array_equal(p1[j, :n], o1) and np.array_equal(err[j, :n], oe)
        assert np.array_equal(keep[j, :n].astype(bool), okeep)
        assert okeep.mean() > 0.5 if n > 100 else True


@pytest.mark.parametrize("groups", [1, 2, 4, "2_chains"])
def test_l2_tensor_core_ragged_pairs_vs_oracle_and_cuda_core(ctx, switches, groups):
    """several pairs of different sizes in one launch (tile tails, pairs with an empty side, counts that are not
    multiples of the 128-row tiles, low-entropy rows full of distance ties), checked against the oracle and against
    the CUDA-core dp4a kernel, for 1 / 2 / 4 epilogue warps per TMEM lane quarter (the column groups of a tile are
    folded with an explicit index tie-break); two warps per quarter is the default and drains a chunk with two min3 trees,
    "2_chains" (ZS_L2_CHAINS) is the same grouping with the four serial (best, second) chains the others use"""
    from zenslam_b200.runtime import match_l2_cross, match_l2_knn2
    if groups == "2_chains":
        switches.set(ctx, "ZS_L2_CHAINS")
    else:
        switches.set(ctx, "ZS_L2_EPI_GROUPS", groups)
    rng = np.random.default_rng(99)
    sizes = [(300, 129), (1, 1), (128, 128), (257, 511), (0, 40), (40, 0), (130, 2), (320, 500)]
    cap_q, cap_t = 320, 512
    q = np.zeros((len(sizes), cap_q, 128), np.float32); t = np.zeros((len(sizes), cap_t, 128), np.float32)
    for k, (a, b) in enumerate(sizes):
        hi = 2 if k == len(sizes) - 1 else 256                      # last pair: 0/1 rows -> many equal distances
        q[k, :a] = rng.integers(0, hi, (a, 128)); t[k, :b] = rng.integers(0, hi, (b, 128))
        if hi == 2:
            q[k, :a, 8:] = 0; t[k, :b, 8:] = 0                      # 8 informative bits: every row has dozens of ties
            t[k, 300] = t[k, 40] = t[k, 170] = q[k, 7]
        # garbage beyond the counts must be ignored
        q[k, a:] = rng.integers(0, 256, (cap_q - a, 128)); t[k, b:] = rng.integers(0, 256, (cap_t - b, 128))
        if a > 5 and b > 100:
            t[k, 100] = q[k, 5]; t[k, 20] = q[k, 5]                 # exact duplicates: tie on distance 0
    nq = np.array([s[0] for s in sizes], np.int32); nt = np.array([s[1] for s in sizes], np.int32)
    dq, dt, dnq, dnt = dev(ctx, q), dev(ctx, t), dev(ctx, nq), dev(ctx, nt)
    idx, dist, ps = [x.cpu().numpy() for x in match_l2_knn2(ctx, dq, dnq, dt, dnt, 0.8)]
    cidx, cdist = [x.cpu().numpy() for x in match_l2_cross(ctx, dq, dnq, dt, dnt)]
    switches.set(ctx, "ZS_L2_NO_TENSOR")
    idx2, dist2, ps2 = [x.cpu().numpy() for x in match_l2_knn2(ctx, dq, dnq, dt, dnt, 0.8)]
    cidx2, cdist2 = [x.cpu().numpy() for x in match_l2_cross(ctx, dq, dnq, dt, dnt)]
    switches.clear(ctx, "ZS_L2_NO_TENSOR")
    for k, (a, b) in enumerate(sizes):
        assert np.array_equal(idx[k, :a], idx2[k, :a]) and np.array_equal(dist[k, :a], dist2[k, :a]), k
        assert np.array_equal(ps[k, :a], ps2[k, :a]) and np.array_equal(cidx[k, :a], cidx2[k, :a])
        assert np.all(idx[k, a:] == -1) and np.all(cidx[k, a:] == -1)
        if a == 0 or b == 0:
            assert np.all(idx[k, :a] == -1)
            continue
        oi, od = oracle.match_l2_knn2(q[k, :a], t[k, :b])
        assert np.array_equal(idx[k, :a], oi), k
        assert np.array_equal(dist[k, :a][oi >= 0], od[oi >= 0])
        oq, ot, odd = oracle.match_l2_cross(q[k, :a], t[k, :b])
        keep = np.nonzero(cidx[k, :a] >= 0)[0]
        assert np.array_equal(keep, oq) and np.array_equal(cidx[k, keep], ot) and np.array_equal(cdist[k, keep], odd)


# ---------------------------------------------------------------------------------------------
# cornerSubPix (PARALLEL_GRID detector)
# ---------------------------------------------------------------------------------------------
def test_corner_subpix_golden(ctx, golden):
    """bit-exact against cv2.cornerSubPix (IPP off) on the fixture, image-border points included"""
    from zenslam_b200.runtime import Pyramid, corner_subpix
    g = golden("subpix")
    L, pts = g["L"], g["pts"]
    h, w = L.shape
    pyr = Pyramid(ctx, w, h, 2, (31, 31), 0)
    pyr.upload(L[None], 1)
    pyr.build(1, 1)
    cap = len(pts) + 7
    xy = np.zeros((1, cap, 2), np.float32); xy[0, :len(pts)] = pts; xy[0, len(pts):] = 9.5     # beyond count: untouched
    d = dev(ctx, xy)
    corner_subpix(pyr, 1, 1, d, dev(ctx, np.array([len(pts)], np.int32)))
    got = d.cpu().numpy()[0]
    assert np.array_equal(got[:len(pts)], g["cv_subpix"])
    assert np.all(got[len(pts):] == 9.5)
    assert np.abs(got[:len(pts)] - g["cv_subpix_ipp"]).max() < 5e-3


@pytest.mark.parametrize("w,h,cell,thr,win,its,eps", [(752, 480, (16, 16), 10, (5, 5), 30, 0.01), (640, 400, (32, 32), 5, (3, 7), 5, 0.0),
                                                      (320, 200, (16, 16), 1, (7, 2), 100, 0.1)])
def test_corner_subpix_batch_vs_oracle(ctx, w, h, cell, thr, win, its, eps):
    from zenslam_b200.runtime import Pyramid, corner_subpix, fast_grid_detect
    B = 3
    seq, _ = syn.stereo_sequence(w, h, B, 31 + w, subpixel=True)
    imgs = np.ascontiguousarray(seq[:, 0])
    pyr = Pyramid(ctx, w, h, B, (31, 31), 0)
    pyr.upload(imgs, 0)
    pyr.build(0, B)
    xy, resp, n = fast_grid_detect(pyr, 0, B, cell, thr)
    before = xy.cpu().numpy().copy()
    corner_subpix(pyr, 0, B, xy, n, win, its, eps)
    after, cnt = xy.cpu().numpy(), n.cpu().numpy()
    moved = 0
    for k in range(B):
        want = oracle.corner_subpix(imgs[k], before[k, :cnt[k]], win, its, eps)
        assert np.array_equal(after[k, :cnt[k]], want), k
        moved += int((np.abs(want - before[k, :cnt[k]]).max(1) > 0).sum())
    assert moved > 0


# ---------------------------------------------------------------------------------------------
# multi-scale ORB detector (`feature: ORB`)
# ---------------------------------------------------------------------------------------------
ORB_KEYS = ("x", "y", "size", "angle", "response", "octave", "desc")


def _orb_gpu(det, images, masks=None):
    r = det.detect_and_compute(images, masks)
    out = []
    for k in range(len(images)):
        n = int(r["n"][k])
        assert n <= det.cap
        xy = r["xy"][k, :n].cpu().numpy()
        out.append(dict(x=xy[:, 0].copy(), y=xy[:, 1].copy(), size=r["size"][k, :n].cpu().numpy(),
                        angle=r["angle"][k, :n].cpu().numpy(), response=r["response"][k, :n].cpu().numpy(),
                        octave=r["octave"][k, :n].cpu().numpy(), desc=r["desc"][k, :n].cpu().numpy()))
    return out


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_orb_multiscale_detector_golden(ctx, golden, case):
    """bit-exact against cv2's ORB detect + compute on the committed fixture (set in canonical order), with and without
    a mask; the pyramid levels equal cv2.resize(INTER_LINEAR_EXACT)"""
    from zenslam_b200.runtime import OrbDetector
    g = golden("orb_detect")
    img = g[case + "_img"]
    mask = g[case + "_mask"] if (case + "_mask") in g.files else None
    h, w = img.shape
    det = OrbDetector(ctx, w, h, 1, fast_threshold=int(g[case + "_thr"]))
    o = _orb_gpu(det, img[None], None if mask is None else mask[None])[0]
    for k in ORB_KEYS:
        assert np.array_equal(o[k], g[case + "_" + k]), k
    if case == "a":
        assert np.array_equal(det.download_level(0, 1), g["resize_313x200"])
        assert np.array_equal(det.download_level(0, 2), g["resize_261x167"])
    det.close()


@pytest.mark.parametrize("w,h,thr,batch", [(752, 480, 10, 3), (1280, 720, 20, 2), (200, 150, 5, 2)])
def test_orb_multiscale_detector_vs_oracle(ctx, w, h, thr, batch):
    """batches of frames (one masked) against the oracle at the reference's ORB parameters; small frames lose their
    top levels to the 31-px edge filter"""
    from zenslam_b200.runtime import OrbDetector
    imgs = np.stack([syn.stereo_pair(w, h, 8100 + i)[0] for i in range(batch)])
    rng = np.random.default_rng(5)
    masks = np.full_like(imgs, 255)
    yy, xx = np.mgrid[0:h, 0:w]
    for _ in range(150):
        cx, cy = rng.integers(0, w), rng.integers(0, h)
        masks[0][(xx - cx) ** 2 + (yy - cy) ** 2 <= 64] = 0
    det = OrbDetector(ctx, w, h, batch, fast_threshold=thr)
    got = _orb_gpu(det, imgs, masks)
    for k in range(batch):
        o = oracle.orb_detect(imgs[k], masks[k], fast_threshold=thr)
        assert len(o["x"]) > 0
        for key in ORB_KEYS:
            assert np.array_equal(got[k][key], o[key]), (k, key)
    # detect only, no mask: same keypoints as the unmasked oracle
    r = det.detect_and_compute(imgs[:1], None, describe=False)
    o = oracle.orb_detect(imgs[0], None, fast_threshold=thr, describe=False)
    n = int(r["n"][0])
    assert n == len(o["x"]) and np.array_equal(r["xy"][0, :n, 0].cpu().numpy(), o["x"])
    assert np.array_equal(r["angle"][0, :n].cpu().numpy(), o["angle"])
    det.close()


def test_simple_detector_feature_orb_mirror(ctx):
    """keypoint_detector_simple with `feature: ORB` through the host-pointer entry (zs_detect_keypoints_orb_host):
    existing keypoints mask discs of radius min(cell)/2 (keypoint_detector_simple.cpp:41-48)"""
    from zenslam_b200 import detection_options
    from zenslam_b200.detection import keypoint_detector_simple
    from zenslam_b200.types import keypoint
    w, h = 640, 400
    img = syn.stereo_pair(w, h, 8200)[0]
    opt = detection_options(feature_detector="ORB", fast_threshold=12, algorithm="SIMPLE")
    det = keypoint_detector_simple(opt, ctx)
    first = det.detect_keypoints(img, None)
    o = oracle.orb_detect(img, None, fast_threshold=12)
    assert len(first) == len(o["x"]) > 100
    assert np.array_equal(np.array([k.pt for k in first], np.float32), np.stack([o["x"], o["y"]], 1))
    assert np.array_equal(np.stack([k.descriptor for k in first]), o["desc"])
    assert [k.octave for k in first] == o["octave"].tolist()
    existing = {k.index: k for k in first[::3]}
    second = det.detect_keypoints(img, existing)
    mask = det._mask(h, w, existing)
    o2 = oracle.orb_detect(img, mask, fast_threshold=12)
    assert np.array_equal(np.array([k.pt for k in second], np.float32), np.stack([o2["x"], o2["y"]], 1))
    assert np.array_equal(np.array([k.angle for k in second], np.float32), o2["angle"])
    assert np.array_equal(np.array([k.response for k in second], np.float32), o2["response"])
    assert second[0].index == first[-1].index + 1                      # sequential global indices


def test_simple_detector_host_entry(ctx):
    """zs_detect_keypoints_simple_host (what the C++ adapter binds for algorithm SIMPLE + feature FAST): full-frame FAST +
    mask + ORB::compute in raster order, and ZS_ERR_CAPACITY instead of a silent truncation"""
    import ctypes as C
    from zenslam_b200._lib import ZenslamCudaError, check, lib
    w, h, thr = 320, 240, 10
    img = syn.stereo_pair(w, h, 8300)[0]
    mask = np.full((h, w), 255, np.uint8); mask[40:90, 100:220] = 0
    cap = ((w + 1) // 2) * ((h + 1) // 2)
    x = np.empty(cap, np.float32); y = np.empty(cap, np.float32); r = np.empty(cap, np.float32)
    d = np.empty((cap, 32), np.uint8); n = C.c_int(0)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    check(lib().zs_detect_keypoints_simple_host(ctx._h, p(img), w, h, w, p(mask), w, thr, p(x), p(y), p(r), p(d), cap, C.byref(n)))
    fx, fy, fs = oracle.fast_detect(img, thr)
    keep = mask[fy, fx] != 0
    fx, fy, fs = fx[keep], fy[keep], fs[keep]
    kept, desc = oracle.orb_compute(img, fx, fy)
    assert n.value == len(kept) > 200
    assert np.array_equal(x[:n.value], fx[kept].astype(np.float32)) and np.array_equal(y[:n.value], fy[kept].astype(np.float32))
    assert np.array_equal(r[:n.value], fs[kept].astype(np.float32)) and np.array_equal(d[:n.value], desc)
    with pytest.raises(ZenslamCudaError, match="capacity"):
        check(lib().zs_detect_keypoints_simple_host(ctx._h, p(img), w, h, w, None, w, thr, p(x), p(y), p(r), p(d), 50, C.byref(n)))
    assert n.value > 50


def test_orb_multiscale_detector_large_frame_and_degenerate_options(ctx):
    """BASELINE config 5 frame size (3840 x 2160): hundreds of thousands of FAST corners per level funnel through the
    histogram / rank-count selection; plus the degenerate settings OpenCV accepts (nfeatures 0, one level)"""
    from zenslam_b200._lib import ZenslamCudaError
    from zenslam_b200.runtime import OrbDetector
    w, h = 3840, 2160
    img = syn.stereo_pair(w, h, 8400)[0]
    det = OrbDetector(ctx, w, h, 1, fast_threshold=10)
    got = _orb_gpu(det, img[None])[0]
    o = oracle.orb_detect(img, None, fast_threshold=10)
    assert len(o["x"]) >= 500
    for key in ORB_KEYS:
        assert np.array_equal(got[key], o[key]), key
    det.close()
    small = img[:300, :400].copy()
    for kw in (dict(nfeatures=0), dict(nlevels=1, nfeatures=50), dict(nlevels=3, nfeatures=7, scale_factor=1.5)):
        det = OrbDetector(ctx, 400, 300, 1, fast_threshold=10, **kw)
        got = _orb_gpu(det, small[None])[0]
        o = oracle.orb_detect(small, None, fast_threshold=10, **kw)
        for key in ORB_KEYS:
            assert np.array_equal(got[key], o[key]), (kw, key)
        det.close()
    with pytest.raises(ZenslamCudaError):
        OrbDetector(ctx, 20, 20, 1, nlevels=16, scale_factor=2.0)          # top levels would be empty
