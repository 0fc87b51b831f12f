"""GPU parity: FPNNClassifier (orthogonal-series PNN) and PNNwithClusteringClassifier of qt_cpp/classification.cpp
 against the unmodified reference classes (oracle/_ref) and the C port."""
import numpy as np
import pytest

from util import make_data

pytestmark = pytest.mark.gpu


def _problem(port, n, d, c, seed, sigma):
    g, gl, q, ql = make_data(port, "l2", n, n // 3, d, c, seed=seed, sigma=sigma)
    rows = np.concatenate([g, q]).astype(np.float64)
    rows /= np.linalg.norm(rows, axis=1, keepdims=True)
    return rows, np.concatenate([gl, ql]).astype(np.int32)


@pytest.mark.parametrize("n,d,c,per,sigma", [(900, 96, 12, 20, 0.5), (700, 100, 10, 30, 2.0), (500, 40, 6, 25, 3.0), (640, 50, 5, 90, 2.5)])
def test_fpnn_matches_reference_build(fir, port, ref_l2, n, d, c, per, sigma):
    rows, labels = _problem(port, n, d, c, n + d, sigma)
    tr, trl, te, avg = ref_l2.cls_setup(rows, labels, c, per, seed=3)
    sd = ref_l2.cls_std()
    for scale in (1.0, 0.33):
        f = fir.Fpnn(rows[tr], trl, c, avg, sd, scale)
        want_bf, a, J = ref_l2.cls_fpnn(0, len(te), scale=scale, bf=True, coefficients=True)
        assert f.J == J
        # coefficients: same statements, CUDA's cos/sin instead of glibc's (a few ulp per term)
        np.testing.assert_allclose(f.coefficients, a, rtol=0, atol=1e-13)
        assert np.array_equal(f.predict(rows[te]), want_bf)
        for ratio in (0.9, 0.97):
            want_seq = ref_l2.cls_fpnn(0, len(te), scale=scale, bf=False, output_ratio=ratio)
            assert np.array_equal(f.predict(rows[te], sequential=True, output_ratio=ratio), want_seq)
        pa, pJ = port.fpnn_train(rows[tr], trl, c, avg, sd, scale)
        assert np.array_equal(f.predict(rows[te]), port.fpnn_predict(pa, pJ, c, avg, sd, rows[te], scale))
        f.close()


def test_fpnn_sequential_differs_from_bf_when_pruning_bites(fir, port, ref_l2):
    """A misleading first chunk: predict_sequentional drops the right class early; the GPU follows the same walk."""
    rows, labels = _problem(port, 700, 100, 10, 800, 2.0)
    r = np.random.default_rng(5)
    rows[:, :32] = r.normal(0, 0.06, size=(len(rows), 32))
    rows /= np.linalg.norm(rows, axis=1, keepdims=True)
    tr, trl, te, avg = ref_l2.cls_setup(rows, labels, 10, 30, seed=3)
    sd = ref_l2.cls_std()
    f = fir.Fpnn(rows[tr], trl, 10, avg, sd, 1.0)
    bf = f.predict(rows[te])
    seq = f.predict(rows[te], sequential=True, output_ratio=0.99)
    assert np.array_equal(bf, ref_l2.cls_fpnn(0, len(te), bf=True))
    assert np.array_equal(seq, ref_l2.cls_fpnn(0, len(te), bf=False, output_ratio=0.99))
    assert not np.array_equal(bf, seq)
    f.close()


@pytest.mark.parametrize("n,d,c,per,sigma,clusters", [(900, 96, 12, 20, 0.5, 3), (700, 100, 10, 30, 2.0, 5), (500, 40, 6, 25, 3.0, 4), (640, 50, 5, 90, 2.5, 7)])
def test_pnn_with_clustering_matches_reference_build(fir, port, ref_l2, n, d, c, per, sigma, clusters):
    rows, labels = _problem(port, n, d, c, n + d, sigma)
    tr, trl, te, avg = ref_l2.cls_setup(rows, labels, c, per, seed=3)
    want_lab, want_med = ref_l2.cls_pnn_clustered(clusters, 0, len(te))
    sel = fir.kmedoids_select(rows[tr], trl, c, clusters)
    assert np.array_equal(sel, want_med)
    assert np.array_equal(sel, port.kmedoids(rows[tr], trl, c, clusters))
    clf = fir.Classifier(rows[tr][sel], trl[sel], c, avg)
    clf.set_total(len(tr))                                             # den = total_training_size (:393)
    lab, sc = clf.pnn(rows[te])
    assert np.array_equal(lab, want_lab)
    clf.close()


def test_kmedoids_small_classes_are_kept_whole(fir, port):
    rows, labels = _problem(port, 120, 24, 4, 1, 1.0)
    order = np.argsort(labels, kind="stable")
    rows, labels = rows[order], labels[order]
    sel = fir.kmedoids_select(rows, labels, 4, 1000)
    assert np.array_equal(sel, np.arange(len(rows)))
    with pytest.raises(fir.FirError):
        fir.kmedoids_select(rows[::-1], labels[::-1], 4, 3)            # not class-major
