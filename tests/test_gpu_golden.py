"""GPU parity against the frozen outputs of the unmodified reference (tests/golden/), through the C-ABI."""
import os

import numpy as np
import pytest

from util import bits

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("metric", ["l2", "chi2", "kl"])
def test_brute_force_golden(fir, metric):
    G = np.load(os.path.join(GOLD, "ref_%s.npz" % metric))
    for d in (8, 32, 100):
        gal = fir.Gallery(G["d%d_gallery" % d], G["d%d_gallery_labels" % d], metric)
        q = G["d%d_queries" % d]
        i, dd = gal.search(q, k=1, path=fir.PATH_EXACT)
        assert np.array_equal(i[:, 0], G["d%d_bf_idx" % d]) and np.array_equal(bits(dd[:, 0]), bits(G["d%d_bf_dist" % d]))
        i, dd = gal.search(q, k=1, max_features=d // 2, path=fir.PATH_EXACT)
        assert np.array_equal(i[:, 0], G["d%d_bf_half_idx" % d]) and np.array_equal(bits(dd[:, 0]), bits(G["d%d_bf_half_dist" % d]))
        n = gal.n
        cand = np.tile(np.arange(n, dtype=np.int32), (len(q), 1))
        assert np.array_equal(bits(gal.distances(q, cand)), bits(G["d%d_all_dist" % d]))
        assert np.array_equal(bits(gal.distances(q, cand, gallery_is_lhs=True)), bits(G["d%d_all_dist_gallery_lhs" % d]))
        if metric == "l2" and d >= 32:
            i, dd = gal.search(q, k=1, path=fir.PATH_TENSOR)
            assert np.array_equal(i[:, 0], G["d%d_bf_idx" % d]) and np.array_equal(bits(dd[:, 0]), bits(G["d%d_bf_dist" % d]))
        gal.close()
    gal = fir.Gallery(G["ties_gallery"], None, metric)
    i, dd = gal.search(G["d32_queries"], k=1, path=fir.PATH_EXACT)
    assert np.array_equal(i[:, 0], G["ties_bf_idx"]) and np.array_equal(bits(dd[:, 0]), bits(G["ties_bf_dist"]))
    gal.close()


@pytest.mark.parametrize("metric", ["l2", "chi2", "kl"])
def test_directed_enumeration_golden(fir, metric):
    G = np.load(os.path.join(GOLD, "ref_%s.npz" % metric))
    gal = fir.Gallery(G["dem_gallery"], G["dem_labels"], metric)
    dem = fir.Dem(gal, pivot0=int(G["dem_pivots"][0]))
    assert np.array_equal(dem.pivots, G["dem_pivots"]) and np.array_equal(bits(dem.P), bits(G["dem_P"]))
    assert bits(np.float32(dem.threshold)) == bits(np.float32(G["dem_threshold"]))
    q = G["dem_queries"]
    for M in (0, 10, 60):
        for name, a in zip(("idx", "dist", "below", "evals"), dem.search(q, M)):
            assert np.array_equal(a, G["dem_M%d_%s" % (M, name)]), (M, name)
    low = fir.Dem(gal, state=(G["dem_pivots"], G["dem_P"], float(G["dem_low_threshold"])))
    for M in (8, 40, 150, 0):
        for name, a in zip(("idx", "dist", "below", "evals"), low.search(q, M)):
            assert np.array_equal(a, G["demlow_M%d_%s" % (M, name)]), (M, name)
    low.close()
    dem.close()
    gal.close()


def test_classification_golden(fir):
    G = np.load(os.path.join(GOLD, "ref_classification.npz"))
    rows, tr, te = G["rows"], G["train_idx"], G["test_idx"]
    clf = fir.Classifier(rows[tr], G["train_labels"], 7, G["avg"])
    for K in (1, 3):
        assert np.array_equal(clf.knn(rows[te], K), G["knn%d" % K])
    lab, sc = clf.pnn(rows[te])
    assert np.array_equal(lab, G["pnn_label"])
    np.testing.assert_allclose(sc, G["pnn_scores"], rtol=1e-5, atol=0)
    assert np.array_equal(clf.pnn_sequential(rows[te]), G["pnn_seq_label"])
    for scale in (1.0, 0.33):
        f = fir.Fpnn(rows[tr], G["train_labels"], 7, G["avg"], G["std"], scale)
        assert f.J == int(G["fpnn_%g_J" % scale])
        a = f.coefficients
        np.testing.assert_allclose(np.array([a.sum(), np.abs(a).sum(), a[::97].sum()]), G["fpnn_%g_a_digest" % scale], rtol=1e-12)
        assert np.array_equal(f.predict(rows[te]), G["fpnn_%g_label" % scale])
        assert np.array_equal(f.predict(rows[te], sequential=True, output_ratio=0.9), G["fpnn_%g_seq_label" % scale])
        f.close()
    sel = fir.kmedoids_select(rows[tr], G["train_labels"], 7, 4)
    assert np.array_equal(sel, G["pnn_clustered_medoids"])
    red = fir.Classifier(rows[tr][sel], G["train_labels"][sel], 7, G["avg"])
    red.set_total(len(tr))
    assert np.array_equal(red.pnn(rows[te], scores=False)[0], G["pnn_clustered_label"])
    red.close()
    rows2, tr2, te2 = G["rows2"], G["train_idx2"], G["test_idx2"]
    clf2 = fir.Classifier(rows2[tr2], G["train_labels2"], 7, G["avg2"])
    assert np.array_equal(clf2.pnn_sequential(rows2[te2]), G["pnn_seq_label2"])
    clf2.close()
    clf.close()


@pytest.mark.parametrize("metric", ["l2", "chi2", "kl"])
def test_twd_golden(fir, metric):
    """Three-way-decision classifiers against the frozen outputs of the reference's own classes (ImageTesting.cpp)."""
    from test_oracle_golden import TWD_CONVENTIONAL, TWD_PROPOSED
    G = np.load(os.path.join(GOLD, "ref_twd.npz"))
    g = G["%s_gallery_f16" % metric].astype(np.float32)
    q = G["%s_queries_f16" % metric].astype(np.float32)
    fir.normalize_rows(g, metric)
    fir.normalize_rows(q, metric)
    gal = fir.Gallery(g, G["%s_gallery_labels" % metric], metric)
    for fc, th in TWD_PROPOSED:
        idx, cls, unrel = gal.twd_proposed(q, fc, th)
        assert np.array_equal(cls, G["%s_prop_%d_%g_class" % (metric, fc, th)])
        assert np.array_equal(unrel, G["%s_prop_%d_%g_unrel" % (metric, fc, th)])
    for kind, th in TWD_CONVENTIONAL:
        idx, cls, unrel = gal.twd_conventional(q, kind, th, 64)
        assert np.array_equal(cls, G["%s_conv_%s_%g_class" % (metric, kind, th)])
        assert np.array_equal(unrel, G["%s_conv_%s_%g_unrel" % (metric, kind, th)])
    gal.close()
