"""GPU parity: fp64 kNN / PNN (classification.cpp) and directed enumeration (ann.cpp) against the oracle."""
import numpy as np
import pytest

from util import bits, make_data

pytestmark = pytest.mark.gpu


def _cls_problem(port, n, d, c, seed):
    g, gl, q, ql = make_data(port, "l2", n, n // 3, d, c, seed=seed)
    rows = np.concatenate([g, q]).astype(np.float64)
    rows /= np.linalg.norm(rows, axis=1, keepdims=True)           # load_image_dataset normalises in double (:829-847)
    return rows, np.concatenate([gl, ql]).astype(np.int32)


def test_knn_pnn_match_reference_build(fir, port, ref_l2):
    rows, labels = _cls_problem(port, 900, 96, 12, seed=1)
    tr, trl, te, avg = ref_l2.cls_setup(rows, labels, 12, 20, seed=5)      # verbatim split_train_test
    clf = fir.Classifier(rows[tr], trl, 12, avg)
    for K in (1, 3, 7):
        assert np.array_equal(clf.knn(rows[te], K), ref_l2.cls_knn(K, 0, len(te)))
    lab, sc = clf.pnn(rows[te])
    rlab, rsc = ref_l2.cls_pnn(0, len(te))
    assert np.array_equal(lab, rlab)
    np.testing.assert_allclose(sc, rsc, rtol=1e-5, atol=0)        # north_star: PNN scores within 1e-5 relative
    clf.close()


def test_knn_pnn_match_port_ragged(fir, port):
    rows, labels = _cls_problem(port, 333, 50, 5, seed=2)
    order = np.argsort(labels[:250], kind="stable")
    tr, te = order, np.arange(250, len(rows))
    avg = rows[tr].mean(axis=0)
    clf = fir.Classifier(rows[tr], labels[tr], 5, avg)
    for K in (1, 5):
        assert np.array_equal(clf.knn(rows[te], K), port.knn(rows[tr], labels[tr], 5, avg, rows[te], K))
    lab, sc = clf.pnn(rows[te])
    psc, plab = port.pnn(rows[tr], labels[tr], 5, avg, rows[te])
    assert np.array_equal(lab, plab)
    np.testing.assert_allclose(sc, psc, rtol=1e-5, atol=0)
    clf.close()


def test_classifier_device_memspace_and_unsorted_rows(fir, port):
    """Queries and outputs in device memory (asynchronous on the classifier's stream), more queries than one 64-row tile,
    a training set that is NOT class-major (the fused Parzen epilogue folds runs of equal labels, whatever the order)."""
    import torch
    rows, labels = _cls_problem(port, 700, 70, 9, seed=4)
    perm = np.random.default_rng(1).permutation(500)
    tr, te = perm, np.arange(500, 700)
    avg = rows[tr].mean(axis=0)
    clf = fir.Classifier(rows[tr], labels[tr], 9, avg)
    qd = torch.from_numpy(rows[te]).cuda()
    lab, sc = clf.pnn(qd)
    knn = clf.knn(qd, 3)
    torch.cuda.synchronize()
    psc, plab = port.pnn(rows[tr], labels[tr], 9, avg, rows[te])
    assert np.array_equal(lab.cpu().numpy(), plab)
    np.testing.assert_allclose(sc.cpu().numpy(), psc, rtol=1e-5, atol=0)
    assert np.array_equal(knn.cpu().numpy(), port.knn(rows[tr], labels[tr], 9, avg, rows[te], 3))
    clf.profile(True)
    clf.pnn(qd)
    ms, n = clf.profile(False)
    assert n == 1 and ms > 0
    clf.close()


def _misleading_head(rows, seed, sigma):
    """Replace the first 32 dimensions with class-independent noise so the sequential PNN prunes on a misleading chunk."""
    r = np.random.default_rng(seed)
    rows = rows.copy()
    rows[:, :32] = r.normal(0, sigma, size=(len(rows), 32))
    return rows / np.linalg.norm(rows, axis=1, keepdims=True)


@pytest.mark.parametrize("n,d,c,per,head", [(900, 96, 12, 20, 0), (600, 200, 9, 15, 0), (700, 100, 10, 20, 0.04), (500, 40, 6, 25, 0.02)])
def test_pnn_sequential_matches_reference_build(fir, port, ref_l2, n, d, c, per, head):
    """PNNClassifier(bf=false) == predict_sequentional (classification.cpp:228-295); SURVEY §8(f) rank 1."""
    rows, labels = _cls_problem(port, n, d, c, seed=n + d)
    if head:
        rows = _misleading_head(rows, seed=d, sigma=head)
    tr, trl, te, avg = ref_l2.cls_setup(rows, labels, c, per, seed=3)
    clf = fir.Classifier(rows[tr], trl, c, avg)
    got = clf.pnn_sequential(rows[te])
    want = ref_l2.cls_pnn_seq(0, len(te))
    assert np.array_equal(got, want)
    assert np.array_equal(got, port.pnn_seq(rows[tr], trl, c, avg, rows[te]))
    if head:                                                           # the pruning walk is really exercised: it differs from predict_bf
        assert not np.array_equal(got, clf.pnn(rows[te], scores=False)[0])
    clf.close()


@pytest.mark.parametrize("metric", ["l2", "chi2"])
def test_dem_build_matches_reference(fir, port, metric, request):
    ref = request.getfixturevalue("ref_" + metric)
    g, gl, q, ql = make_data(port, metric, 3000, 100, 64, 25, seed=3)
    rd = ref.dem_create(g, gl, seed=11)                           # verbatim ctor (random_shuffle seeded through srand)
    gal = fir.Gallery(g, gl, metric)
    dem = fir.Dem(gal, pivot0=int(rd.pivots[0]))
    assert dem.n_pivots == rd.n_pivots and dem.chain_rows == 45
    assert np.array_equal(dem.pivots, rd.pivots)
    assert np.array_equal(bits(dem.P), bits(rd.P()))
    assert bits(np.float32(dem.threshold)) == bits(np.float32(rd.threshold))
    pb = port.dem_build(metric, g, gl, int(rd.pivots[0]))
    assert np.array_equal(bits(dem.min_other), bits(pb["min_other"]))
    for M in (0, 50, 300):
        out = dem.search(q, M)
        want = rd.search(q, M)
        for a, b in zip(out, want):
            assert np.array_equal(a, b)
    rd.close()
    dem.close()
    gal.close()


def test_dem_search_candidate_walk(fir, port, ref_l2):
    """Thresholds far below the data's distances force the ordered candidate walk (ann.cpp:469-477)."""
    g, gl, q, ql = make_data(port, "l2", 4000, 120, 64, 30, seed=4)
    gal = fir.Gallery(g, gl, "l2")
    dem0 = fir.Dem(gal, pivot0=17)
    piv, P = dem0.pivots, dem0.P
    d1, _ = port.bf("l2", g, q)
    nn = port.distance  # noqa
    bf_i, bf_d = port.bf("l2", g, q)
    for thr in (float(np.median(bf_d)), float(bf_d.min()) * 0.5):          # ~half the queries can hit / nobody can
        dem = fir.Dem(gal, state=(piv, P, thr))
        rdem = ref_l2.dem_create_injected(g, gl, piv, P, thr)
        for M in (20, 33, 100, 700, 2500, 0):
            out = dem.search(q, M)
            want = rdem.search(q, M)
            pw = port.dem_search("l2", g, piv, P, thr, M, q)
            for a, b, c in zip(out, want, pw):
                assert np.array_equal(a, c), (thr, M)                      # defined (likelihood, index) order
                assert np.array_equal(a, b), (thr, M)                      # and the verbatim reference
        rdem.close()
        dem.close()
    dem0.close()
    gal.close()


def test_dem_index_walk_quirk(fir, port, ref_l2):
    """A pivot whose gallery index is smaller than its ordinal makes the reference's index 'swap' drop one
    candidate and keep a pivot in the tail; the GPU replay must agree."""
    g, gl, q, ql = make_data(port, "l2", 1500, 80, 48, 10, seed=6)
    gal = fir.Gallery(g, gl, "l2")
    piv = np.array([900, 5, 1, 700, 2, 40, 3, 1200], np.int32)             # p_2 = 1 < 2, p_4 = 2 < 4, p_6 = 3 < 6
    P = ref_l2.all_distances(g, g[piv], gallery_is_lhs=True)
    thr = 1e-9
    dem = fir.Dem(gal, state=(piv, P, thr))
    rdem = ref_l2.dem_create_injected(g, gl, piv, P, thr)
    for M in (9, 30, 400, 0):
        for a, b in zip(dem.search(q, M), rdem.search(q, M)):
            assert np.array_equal(a, b), M
    rdem.close()
    dem.close()
    gal.close()


def test_dem_search_large_gallery_fast_first_round(fir, port):
    """40k rows: the first round of the candidate walk goes through the per-warp-minima bound (dem_fast_*), later rounds
    through the general radix path starting from where the fast round ended.  Both a realistic threshold (early exits) and
    a threshold far below the data (long ordered walks over several rounds) against the port."""
    g, gl, q, ql = make_data(port, "l2", 40000, 96, 32, 400, seed=6, sigma=1.5)
    g[20000:20040] = g[100:140]                                  # exact duplicates: ties at equal likelihood, resolved by row
    gal = fir.Gallery(g, gl, "l2")
    dem = fir.Dem(gal, pivot0=123, max_chain=40)
    piv, P, thr = dem.pivots, dem.P, float(dem.threshold)
    for M in (0, 200, 300, 3000):
        got = dem.search(q, M)
        want = port.dem_search("l2", g, piv, P, thr, M, q)
        for name, a, b in zip(("idx", "dist", "below", "evals"), got, want):
            assert np.array_equal(a, b), (M, name)
    low = fir.Dem(gal, state=(piv, P, thr * 0.05))               # nothing is ever under the threshold: full budget is walked
    for M in (100, 256, 257, 700, 2600):
        got = low.search(q, M)
        want = port.dem_search("l2", g, piv, P, thr * 0.05, M, q)
        for name, a, b in zip(("idx", "dist", "below", "evals"), got, want):
            assert np.array_equal(a, b), (M, name)
        assert int(got[3].max()) == M
    low.close()
    dem.close()
    gal.close()


def test_dem_tensor_first_round_matches_verbatim_reference(fir, port, ref_l2):
    """>= 8192 rows with 32 pivots: the first 24 candidates of every query come from the tensor-core brute force over the
    pivot-space gallery (dem_round0_merge_kernel).  The pivot list carries the index-walk quirk (pivots whose gallery index is
    below their ordinal, so some rows have truncated likelihoods and one stays in the tail), budgets end inside / exactly at /
    just after the first round, and a threshold far below the data forces the walk to resume in the general path behind the
    round's closing key.  Checked against the verbatim recognize(); a second gallery with duplicated rows (equal likelihoods —
    where the reference's unstable partial_sort is implementation-defined) is checked against the port's (likelihood, row) order."""
    g, gl, q, ql = make_data(port, "l2", 12000, 150, 64, 120, seed=9, sigma=1.2)
    rng = np.random.default_rng(3)
    piv = rng.choice(np.arange(40, 12000), 32, replace=False).astype(np.int32)
    piv[5], piv[9], piv[20], piv[31] = 2, 7, 11, 30               # p_i < i: the reference's 'swap' duplicates / drops an index
    for dup in (False, True):
        if dup:
            g = g.copy()
            g[7000:7030] = g[300:330]                            # exact duplicates: ties at equal likelihood, resolved by row
        gal = fir.Gallery(g, gl, "l2")
        P = ref_l2.all_distances(g, g[piv], gallery_is_lhs=True)
        other = np.array([P[i][gl != gl[piv[i]]].min() for i in range(32)])
        for thr in (float(np.sort(other)[0]), 1e-9):
            dem = fir.Dem(gal, state=(piv, P, thr))
            rdem = None if dup else ref_l2.dem_create_injected(g, gl, piv, P, thr)
            for M in (33, 35, 55, 56, 57, 80, 600, 0 if thr > 1e-6 else 1500):
                got = dem.search(q, M)
                assert dem.search_stats()["tensor_round"]
                want = port.dem_search("l2", g, piv, P, thr, M, q) if dup else rdem.search(q, M)
                for name, a, b in zip(("idx", "dist", "below", "evals"), got, want):
                    assert np.array_equal(a, b), (dup, thr, M, name)
            if rdem is not None:
                rdem.close()
            dem.close()
        gal.close()
