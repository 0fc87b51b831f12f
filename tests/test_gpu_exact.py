"""GPU parity: the exact CUDA-core kernels against the oracle, through the C-ABI (bit-exact)."""
import numpy as np
import pytest

from util import bits, make_data

pytestmark = pytest.mark.gpu

CASES = [  # metric, n, nq, d, classes
    ("l2", 1000, 70, 128, 10),
    ("l2", 333, 65, 100, 7),      # ragged: D % 32 != 0, N % 64 != 0
    ("chi2", 700, 40, 96, 9),
    ("kl", 500, 33, 80, 5),
    ("l2", 5, 3, 8, 2),           # tiny
]


@pytest.mark.parametrize("metric,n,nq,d,c", CASES)
def test_topk_matches_oracle(fir, port, metric, n, nq, d, c):
    g, gl, q, ql = make_data(port, metric, n, nq, d, c)
    gal = fir.Gallery(g, gl, metric)
    for k in (1, 10):
        if k > n:
            continue
        idx, dist = gal.search(q, k=k, path=fir.PATH_EXACT)
        oi, od = port.topk(metric, g, q, k)
        assert np.array_equal(idx, oi)
        assert np.array_equal(bits(dist), bits(od))
    bi, bd = port.bf(metric, g, q)
    idx, dist = gal.search(q, k=1, path=fir.PATH_EXACT)
    assert np.array_equal(idx[:, 0], bi) and np.array_equal(bits(dist[:, 0]), bits(bd))
    gal.close()


def test_topk_matches_reference_build(fir, port, ref_l2, ref_chi2, ref_kl):
    for metric, ref in (("l2", ref_l2), ("chi2", ref_chi2), ("kl", ref_kl)):
        g, gl, q, ql = make_data(port, metric, 900, 50, 160, 8, seed=3)
        gal = fir.Gallery(g, gl, metric)
        idx, dist = gal.search(q, k=1, path=fir.PATH_EXACT)
        ri, rd = ref.bf(g, q, gl)
        assert np.array_equal(idx[:, 0], ri)
        assert np.array_equal(bits(dist[:, 0]), bits(rd))
        gal.close()


def test_ties_lowest_index_wins(fir, port):
    g, gl, q, ql = make_data(port, "l2", 300, 20, 64, 4, seed=1)
    g[150:300] = g[0:150]          # every row duplicated ⇒ exact ties
    gal = fir.Gallery(g, None, "l2")
    idx, dist = gal.search(q, k=4, path=fir.PATH_EXACT)
    oi, od = port.topk("l2", g, q, 4)
    assert np.array_equal(idx, oi) and np.array_equal(bits(dist), bits(od))
    assert (idx[:, 0] < 150).all() and (idx[:, 1] == idx[:, 0] + 150).all()
    gal.close()


def test_prefix_max_features(fir, port):
    g, gl, q, ql = make_data(port, "l2", 400, 30, 96, 6, seed=2)
    gal = fir.Gallery(g, gl, "l2")
    for mf in (32, 50, 96):
        idx, dist = gal.search(q, k=1, max_features=mf, path=fir.PATH_EXACT)
        oi, od = port.bf("l2", g, q, max_features=mf)
        assert np.array_equal(idx[:, 0], oi) and np.array_equal(bits(dist[:, 0]), bits(od))
    gal.close()


def test_zero_vectors_chi2_kl(fir, port):
    rng = np.random.default_rng(5)
    for metric in ("chi2", "kl"):
        g = np.maximum(rng.standard_normal((200, 64)), 0).astype(np.float32)
        g[:, ::3] = 0                      # shared zero columns exercise the l+r>0 guard (db_features.cpp:29)
        q = np.maximum(rng.standard_normal((17, 64)), 0).astype(np.float32)
        q[:, ::3] = 0
        g, q = port.normalize_rows(metric, g), port.normalize_rows(metric, q)
        gal = fir.Gallery(g, None, metric)
        idx, dist = gal.search(q, k=3, path=fir.PATH_EXACT)
        oi, od = port.topk(metric, g, q, 3)
        assert np.array_equal(idx, oi) and np.array_equal(bits(dist), bits(od))
        gal.close()


def test_kl_tile_special_values(fir, port):
    """The KL tile kernel decides per warp whether a step's query element is zero (then the step is r*logf(2), no division, no
    logf) — the shortcut must be invisible: signed zeros, negative and mixed-sign elements, denormals, values whose double
    overflows (2r = inf), infinities and NaNs on either side all give the port's bits, through top-k, class minima and PNN sums."""
    rng = np.random.default_rng(17)
    n, nq, d = 300, 70, 96
    g = np.maximum(rng.standard_normal((n, d)), 0).astype(np.float32)
    q = np.maximum(rng.standard_normal((nq, d)), 0).astype(np.float32)
    g[:, 5] = 0; q[:, 7] = 0                                        # a zero column on each side
    q[3, :] = 0; g[11, :] = 0                                       # all-zero rows
    q[4, ::2] = -0.0; g[12, 1::2] = -0.0                            # negative zeros
    q[5, :6] = [-0.25, -1e-30, 1e-41, 3e38, 2e38, 1.7014118e38]     # negative, tiny, denormal, 2l overflows
    g[13, :6] = [1e-41, -0.5, 3e38, 1e-45, 1.7014120e38, 2.5e38]    # 2r overflows against zero and non-zero l
    g[14, 20] = np.inf; g[15, 21] = np.nan; q[6, 22] = np.inf; q[8, 23] = np.nan
    q[9, :] = 0; q[9, 2] = 1.0                                      # one non-zero element: every other step is the shortcut
    gl = np.sort(rng.integers(0, 9, n)).astype(np.int32)
    gal = fir.Gallery(g, gl, "kl")
    idx, dist = gal.search(q, k=4, path=fir.PATH_EXACT)
    oi, od = port.topk("kl", g, q, 4)
    fin = ~np.isnan(od)                                             # rows with NaN distances: the reference never accepts them; compare the rest bit for bit
    assert np.array_equal(np.isnan(dist), np.isnan(od))
    assert np.array_equal(idx[fin], oi[fin]) and np.array_equal(bits(dist)[fin], bits(od)[fin])
    mn, arg = gal.class_min(q)
    omn, oarg = port.class_min("kl", g, gl, gal.n_classes, q)
    assert np.array_equal(arg, oarg) and np.array_equal(bits(mn), bits(omn))
    gal.close()


def test_pair_distances_both_operand_orders(fir, port):
    rng = np.random.default_rng(7)
    for metric in ("l2", "chi2", "kl"):
        g, gl, q, ql = make_data(port, metric, 257, 9, 200, 5, seed=4)
        gal = fir.Gallery(g, gl, metric)
        cand = rng.integers(0, 257, (9, 40)).astype(np.int32)
        cand[0, 3] = -1
        for lhs in (False, True):
            out = gal.distances(q, cand, gallery_is_lhs=lhs)
            for i in range(9):
                for s in range(40):
                    j = cand[i, s]
                    if j < 0:
                        assert np.isinf(out[i, s])
                        continue
                    want = port.distance(metric, g[j], q[i]) if lhs else port.distance(metric, q[i], g[j])
                    assert bits(np.float32(out[i, s])) == bits(np.float32(want)), (metric, lhs, i, s)
        gal.close()


def test_class_min_and_pnn(fir, port):
    for metric, var in (("l2", 2e-4), ("chi2", 1e-3), ("kl", 5e-4)):
        g, gl, q, ql = make_data(port, metric, 640, 25, 96, 12, seed=6)
        gal = fir.Gallery(g, gl, metric)
        mn, arg = gal.class_min(q)
        omn, oarg = port.class_min(metric, g, gl, gal.n_classes, q)
        assert np.array_equal(arg, oarg) and np.array_equal(bits(mn), bits(omn))
        sc, lab = gal.pnn_scores(q, var)
        osc, olab = port.pnn_div(metric, g, gl, gal.n_classes, q, var)
        assert np.array_equal(lab, olab)
        np.testing.assert_allclose(sc, osc, rtol=1e-5, atol=0)   # BASELINE north_star: PNN scores within 1e-5 relative
        gal.close()


def test_normalize_rows_matches_loader(fir, port):
    rng = np.random.default_rng(11)
    for metric in ("l2", "chi2"):
        raw = rng.standard_normal((300, 77)).astype(np.float32)
        raw[rng.random(raw.shape) < 0.1] *= 1e-4        # values around the 1e-4 zeroing threshold (db_features.cpp:85)
        if metric != "l2":
            raw = np.abs(raw)
        want = port.normalize_rows(metric, raw)
        got = fir.normalize_rows(raw.copy(), metric)
        assert np.array_equal(bits(got), bits(want))


def test_device_pointers_and_merge(fir, port):
    import torch
    g, gl, q, ql = make_data(port, "l2", 1200, 64, 128, 10, seed=8)
    oi, od = port.topk("l2", g, q, 5)
    # two row shards with global index offsets, merged on the device
    parts_d, parts_i = [], []
    for lo, hi in ((0, 500), (500, 1200)):
        gal = fir.Gallery(torch.from_numpy(g[lo:hi]).cuda(), torch.from_numpy(gl[lo:hi]).cuda(), "l2", index_offset=lo)
        i, d = gal.search(torch.from_numpy(q).cuda(), k=5, path=fir.PATH_EXACT)
        parts_i.append(i)
        parts_d.append(d)
        torch.cuda.synchronize()
        gal.close()
    mi, md = fir.merge_topk(torch.stack(parts_d), torch.stack(parts_i))
    torch.cuda.synchronize()
    assert np.array_equal(mi.cpu().numpy(), oi) and np.array_equal(bits(md.cpu().numpy()), bits(od))


def test_error_behaviour(fir):
    with pytest.raises(fir.FirError):
        fir.Gallery(np.zeros((0, 8), np.float32))
    gal = fir.Gallery(np.ones((4, 8), np.float32))
    with pytest.raises(fir.FirError):
        gal.search(np.ones((2, 8), np.float32), k=0)
    idx, dist = gal.search(np.full((2, 8), 1e4, np.float32), k=1, path=fir.PATH_EXACT)   # all distances >= 1e5 ⇒ -1 (ann.cpp:115-116)
    assert (idx == -1).all()
    gal.close()


@pytest.mark.parametrize("metric", ["l2", "chi2", "kl"])
def test_small_batch_streaming_path(fir, port, metric):
    """<= 8 queries take the one-pass streaming kernels (latency mode); same bits as the reference."""
    g, gl, q, ql = make_data(port, metric, 5000, 8, 200, 12, seed=13)
    g[4000:4100] = g[10:110]                              # ties across segments
    gal = fir.Gallery(g, gl, metric)
    for nq in (1, 3, 5, 6, 7, 8):                        # (3, 5, 6, 7: the kernel runs 4 / 8 query slots and must drop the spare ones)
        for k in (1, 7, 16):
            idx, dist = gal.search(q[:nq], k=k, path=fir.PATH_EXACT)
            oi, od = port.topk(metric, g, q[:nq], k)
            assert np.array_equal(idx, oi) and np.array_equal(bits(dist), bits(od)), (nq, k)
        mn, arg = gal.class_min(q[:nq])
        omn, oarg = port.class_min(metric, g, gl, gal.n_classes, q[:nq])
        assert np.array_equal(arg, oarg) and np.array_equal(bits(mn), bits(omn))
        sc, lab = gal.pnn_scores(q[:nq], 5e-4)
        osc, olab = port.pnn_div(metric, g, gl, gal.n_classes, q[:nq], 5e-4)
        assert np.array_equal(lab, olab)
        np.testing.assert_allclose(sc, osc, rtol=1e-5, atol=0)
    idx, dist = gal.search(q[:2], k=1, max_features=96, path=fir.PATH_EXACT)
    oi, od = port.bf(metric, g, q[:2], max_features=96)
    assert np.array_equal(idx[:, 0], oi) and np.array_equal(bits(dist[:, 0]), bits(od))
    gal.close()


@pytest.mark.parametrize("metric,n,nq,d,c", [("chi2", 3000, 200, 160, 20), ("kl", 2500, 150, 96, 15), ("chi2", 777, 40, 100, 6)])
def test_approximate_pass_with_exact_rerank(fir, port, metric, n, nq, d, c):
    """chi2/KL batched search through the fast approximate tiles: the certificate + exact rerank must return the
    reference's bits, also with duplicated rows (exact ties ⇒ certificate failures ⇒ exact re-run)."""
    g, gl, q, ql = make_data(port, metric, n, nq, d, c, seed=n % 11)
    g[n - 60:] = g[:60]
    gal = fir.Gallery(g, gl, metric)
    for k in (1, 10, 20):
        idx, dist = gal.search(q, k=k, path=fir.PATH_APPROX)
        assert gal.stats()["path_used"] == fir.PATH_APPROX
        oi, od = port.topk(metric, g, q, k, nthreads=8)
        assert np.array_equal(idx, oi) and np.array_equal(bits(dist), bits(od)), (metric, k, gal.stats())
    gal.close()


def test_kl_approx_path_with_mixed_sign_rows(fir, port):
    """The KL error model of the approximate path assumes non-negative features; the reference's guards (`l + r > 0`, `l > 0`,
    db_features.cpp:33-36) also accept mixed-sign rows (PCA'd features), where log(2l/(l+r)) is unbounded.  Mixed-sign queries
    must be re-run exactly, and a mixed-sign gallery must not use the approximate kernels at all — answers stay the port's."""
    g, gl, q, ql = make_data(port, "kl", 5000, 96, 96, 12, seed=31)
    rng = np.random.default_rng(5)
    q2 = q.copy()
    rows = rng.choice(len(q2), 24, replace=False)
    for r in rows:                                               # negative entries that nearly cancel a gallery value: l + r -> 0+
        cols = rng.choice(96, 6, replace=False)
        q2[r, cols] = -g[rng.integers(0, len(g)), cols] * np.float32(1 - 1e-6)
    gal = fir.Gallery(g, gl, "kl")
    idx, dist = gal.search(q2, k=5, path=fir.PATH_APPROX)
    st = gal.stats()
    n_mixed = int((q2 < 0).any(axis=1).sum())                    # (a ReLU zero times -1 is -0.0: not a negative value)
    assert n_mixed >= 10
    assert st["path_used"] == fir.PATH_APPROX and st["n_fallback"] >= n_mixed        # every mixed-sign query went to the exact re-run
    oi, od = port.topk("kl", g, q2, 5)
    assert np.array_equal(idx, oi) and np.array_equal(bits(dist), bits(od))
    gal.close()
    g2 = g.copy()
    g2[rng.choice(len(g2), 200, replace=False), 3] *= -1         # mixed-sign gallery
    gal = fir.Gallery(g2, gl, "kl")
    idx, dist = gal.search(q, k=5, path=fir.PATH_APPROX)
    assert gal.stats()["path_used"] == fir.PATH_EXACT
    oi, od = port.topk("kl", g2, q, 5)
    assert np.array_equal(idx, oi) and np.array_equal(bits(dist), bits(od))
    gal.close()
