"""CPU: the C restatement against the live reference build (oracle/_ref) on fresh random inputs — wider
coverage than the frozen vectors.  Skipped where oracle/_ref was not built (no /root/reference and no prebuilt .so)."""
import numpy as np
import pytest

from util import bits, make_data


@pytest.mark.parametrize("metric", ["l2", "chi2", "kl"])
def test_distances_and_bf(port, metric, request):
    ref = request.getfixturevalue("ref_" + metric)
    for (n, nq, d, c, seed) in ((257, 19, 96, 7, 1), (64, 5, 1536, 4, 2), (500, 9, 33, 11, 3)):
        g, gl, q, ql = make_data(port, metric, n, nq, d, c, seed=seed)
        D = ref.all_distances(g, q)
        Dp = np.array([[port.distance(metric, a, b) for b in g] for a in q], np.float32)
        assert np.array_equal(bits(D), bits(Dp))
        ri, rd = ref.bf(g, q, gl)
        pi, pd = port.bf(metric, g, q)
        assert np.array_equal(ri, pi) and np.array_equal(bits(rd), bits(pd))
        ri, rd = ref.bf(g, q, gl, nthreads=4)
        assert np.array_equal(ri, pi)


def test_dem_and_classifiers(port, ref_l2):
    g, gl, q, ql = make_data(port, "l2", 1200, 60, 48, 12, seed=9)
    dem = ref_l2.dem_create(g, gl, seed=3)
    b = port.dem_build("l2", g, gl, int(dem.pivots[0]))
    assert np.array_equal(dem.pivots, b["pivots"][: b["n_pivots"]]) and np.array_equal(bits(dem.P()), bits(b["P"]))
    assert bits(np.float32(dem.threshold)) == bits(np.float32(b["threshold"]))
    low = float(dem.threshold) * 0.02
    dinj = ref_l2.dem_create_injected(g, gl, dem.pivots, dem.P(), low)
    for M in (0, 33, 200):
        for a, c in zip(dinj.search(q, M), port.dem_search("l2", g, dem.pivots, dem.P(), low, M, q)):
            assert np.array_equal(a, c)
    dem.close()
    dinj.close()
    rows = np.concatenate([g, q]).astype(np.float64)
    rows /= np.linalg.norm(rows, axis=1, keepdims=True)
    labels = np.concatenate([gl, ql]).astype(np.int32)
    tr, trl, te, avg = ref_l2.cls_setup(rows, labels, 12, 15, seed=2)
    for K in (1, 3):
        assert np.array_equal(ref_l2.cls_knn(K, 0, len(te)), port.knn(rows[tr], trl, 12, avg, rows[te], K))
    rl, rs = ref_l2.cls_pnn(0, len(te))
    ps, pl = port.pnn(rows[tr], trl, 12, avg, rows[te])
    assert np.array_equal(rl, pl) and np.array_equal(bits(rs), bits(ps))
    assert np.array_equal(ref_l2.cls_pnn_seq(0, len(te)), port.pnn_seq(rows[tr], trl, 12, avg, rows[te]))
    # a misleading first 32-dimension chunk: predict_sequentional prunes the right class and departs from predict_bf
    r = np.random.default_rng(0)
    rows[:, :32] = r.normal(0, 0.04, size=(len(rows), 32))
    rows /= np.linalg.norm(rows, axis=1, keepdims=True)
    tr, trl, te, avg = ref_l2.cls_setup(rows, labels, 12, 15, seed=2)
    seq = ref_l2.cls_pnn_seq(0, len(te))
    assert np.array_equal(seq, port.pnn_seq(rows[tr], trl, 12, avg, rows[te]))
    assert not np.array_equal(seq, ref_l2.cls_pnn(0, len(te))[0])


@pytest.mark.parametrize("metric", ["l2", "chi2", "kl"])
def test_twd_port_matches_reference(port, request, metric):
    """ImageTesting.cpp's TWD classifiers: the C port against the verbatim classes (class + num_of_unreliable per query)."""
    ref = request.getfixturevalue("ref_" + metric)
    g, gl, q, ql = make_data(port, metric, 300, 120, 288, 8, seed=6, sigma=2.0)
    seen = 0
    for fc, th in ((32, 0.7), (64, 0.7), (96, 0.8), (32, 1.5)):
        idx, cls, unrel = port.twd_proposed(metric, g, gl, q, fc, th)
        rc, ru = ref.twd("proposed", g, gl, 8, q, fc, th)
        assert np.array_equal(cls, rc) and np.array_equal(unrel, ru), (fc, th)
        seen += int(unrel.sum())
    for kind, th in (("posteriors", 0.24), ("diff", 0.003), ("ratio", 0.7), ("diff", 0.0003), ("ratio", 0.9)):
        for fc in (64, 17):
            idx, cls, unrel = port.twd_conventional(metric, g, gl, 8, q, kind, th, fc)
            rc, ru = ref.twd("conventional", g, gl, 8, q, fc, th, kind)
            assert np.array_equal(cls, rc) and np.array_equal(unrel, ru), (kind, th, fc)
            seen += int(unrel.sum())
    # BruteForceClassifier(max_feats) = recognize_image_bf over a prefix
    rc, _ = ref.twd("bf", g, gl, 8, q, 64, 0)
    assert np.array_equal(rc, gl[port.bf(metric, g, q, max_features=64)[0]])
    assert seen > 0


def test_fpnn_and_clustering_port_matches_reference(port, ref_l2):
    """FPNNClassifier / PNNwithClusteringClassifier: coefficients bit-identical (same libm), labels and medoids identical."""
    g, gl, q, ql = make_data(port, "l2", 700, 230, 100, 10, seed=800, sigma=2.0)
    rows = np.concatenate([g, q]).astype(np.float64)
    rows /= np.linalg.norm(rows, axis=1, keepdims=True)
    labels = np.concatenate([gl, ql]).astype(np.int32)
    tr, trl, te, avg = ref_l2.cls_setup(rows, labels, 10, 30, seed=3)
    sd = ref_l2.cls_std()
    x = rows[tr]
    assert np.allclose(sd, x.std(axis=0, ddof=1), rtol=1e-9)
    for scale in (1.0, 0.33):
        lab, a, J = ref_l2.cls_fpnn(0, len(te), scale=scale, bf=True, coefficients=True)
        pa, pJ = port.fpnn_train(x, trl, 10, avg, sd, scale)
        assert pJ == J and np.array_equal(bits64(pa), bits64(a))
        assert np.array_equal(port.fpnn_predict(pa, pJ, 10, avg, sd, rows[te], scale), lab)
        for ratio in (0.9, 0.99):
            seq = ref_l2.cls_fpnn(0, len(te), scale=scale, bf=False, output_ratio=ratio)
            assert np.array_equal(port.fpnn_predict(pa, pJ, 10, avg, sd, rows[te], scale, sequential=True, output_ratio=ratio), seq)
    for clusters in (3, 5):
        lab, med = ref_l2.cls_pnn_clustered(clusters, 0, len(te))
        assert np.array_equal(port.kmedoids(x, trl, 10, clusters), med)


def bits64(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.int64)
