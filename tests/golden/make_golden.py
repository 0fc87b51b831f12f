"""Generates tests/golden/*.npz from the UNMODIFIED reference code (oracle/_ref, built from /root/reference by
oracle/Makefile).  Run in the build container only:  python tests/golden/make_golden.py
The reference ships no golden vectors of its own; these frozen known-answer files are outputs of
the reference's own translation units on inputs fixed here, and pin both the C restatement (CPU tests) and the
CUDA path (GPU tests) on machines where /root/reference does not exist."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle_py  # noqa: E402
import importlib  # noqa: E402

synth = importlib.import_module("fast-image-recognition_b200.synth")


TWD_PROPOSED = [(32, 0.7), (64, 0.7), (32, 0.9)]
TWD_CONVENTIONAL = [("posteriors", 0.24), ("diff", 0.003), ("ratio", 0.7), ("diff", 0.0003), ("ratio", 0.9)]


def raw(metric, n, nq, d, c, seed):
    return synth.make_split(n, nq, d, c, metric, seed=seed)


def main():
    oracle_py.build()
    for metric in ("l2", "chi2", "kl"):
        ref = oracle_py.Ref(metric)
        out = {}
        for d in (8, 32, 100):
            g, gl, q, ql = raw(metric, 96, 12, d, 6, seed=d)
            # normalise through the reference's own loader: write its text format, loadImages, no split randomisation
            path = os.path.join(HERE, "_tmp_%s_%d.txt" % (metric, d))
            synth.write_features_file(path, np.concatenate([g, q]), ["c%02d" % l for l in np.concatenate([gl, ql])])
            # loader parses '{:f}' text, so feed the parsed values to everything downstream
            gr, grl, gri, tr, trl, tri = ref.load_split(path, d, seed=13, randomize=False)
            os.remove(path)
            rows = np.concatenate([gr, tr])      # under USE_CALTECH the first 30 per class go to the gallery (db_features.cpp:133)
            out["d%d_gallery" % d], out["d%d_gallery_labels" % d] = gr, grl
            out["d%d_test" % d], out["d%d_test_labels" % d] = tr, trl
            out["d%d_raw" % d] = np.concatenate([g, q])
            out["d%d_raw_labels" % d] = np.concatenate([gl, ql])
            if len(tr) == 0:      # every class has <= 30 images here: use the tail of the gallery as queries instead
                tr = gr[-12:]
            out["d%d_all_dist" % d] = ref.all_distances(gr, tr)
            out["d%d_all_dist_gallery_lhs" % d] = ref.all_distances(gr, tr, gallery_is_lhs=True)
            bi, bd = ref.bf(gr, tr, grl)
            out["d%d_bf_idx" % d], out["d%d_bf_dist" % d] = bi, bd
            bi, bd = ref.bf(gr, tr, grl, max_features=d // 2)
            out["d%d_bf_half_idx" % d], out["d%d_bf_half_dist" % d] = bi, bd
            out["d%d_queries" % d] = tr
        # ties: duplicated gallery rows
        g = out["d32_gallery"].copy()
        h = len(g) // 2
        g[h:2 * h] = g[:h]
        out["ties_gallery"] = g
        bi, bd = ref.bf(g, out["d32_queries"], None)
        out["ties_bf_idx"], out["ties_bf_dist"] = bi, bd
        # directed enumeration: verbatim ctor (srand(7)) + recognize at several budgets / an injected low threshold
        g, gl, q, ql = raw(metric, 400, 40, 24, 8, seed=77)
        port = oracle_py.Port()
        g, q = port.normalize_rows(metric, g), port.normalize_rows(metric, q)
        assert np.array_equal(ref.all_distances(g, q), np.array([[port.distance(metric, a, b) for b in g] for a in q], np.float32))
        dem = ref.dem_create(g, gl, seed=7)
        out["dem_gallery"], out["dem_labels"], out["dem_queries"] = g, gl, q
        out["dem_pivots"], out["dem_threshold"], out["dem_P"] = dem.pivots, np.float32(dem.threshold), dem.P()
        for M in (0, 10, 60):
            r = dem.search(q, M)
            for name, a in zip(("idx", "dist", "below", "evals"), r):
                out["dem_M%d_%s" % (M, name)] = a
        low = np.float32(dem.threshold * 0.05)
        dinj = ref.dem_create_injected(g, gl, dem.pivots, dem.P(), low)
        out["dem_low_threshold"] = low
        for M in (8, 40, 150, 0):
            r = dinj.search(q, M)
            for name, a in zip(("idx", "dist", "below", "evals"), r):
                out["demlow_M%d_%s" % (M, name)] = a
        dem.close()
        dinj.close()
        np.savez_compressed(os.path.join(HERE, "ref_%s.npz" % metric), **out)
        print(metric, "->", len(out), "arrays")
    # fp64 kNN / PNN (classification.cpp), verbatim split_train_test with srand(5)
    ref = oracle_py.Ref("l2")
    g, gl, q, ql = raw("l2", 300, 100, 40, 7, seed=5)
    rows = np.concatenate([g, q]).astype(np.float64)
    rows /= np.linalg.norm(rows, axis=1, keepdims=True)
    labels = np.concatenate([gl, ql]).astype(np.int32)
    tr, trl, te, avg = ref.cls_setup(rows, labels, 7, 12, seed=5)
    out = dict(rows=rows, labels=labels, train_idx=tr, train_labels=trl, test_idx=te, avg=avg)
    for K in (1, 3):
        out["knn%d" % K] = ref.cls_knn(K, 0, len(te))
    out["pnn_label"], out["pnn_scores"] = ref.cls_pnn(0, len(te))
    out["pnn_seq_label"] = ref.cls_pnn_seq(0, len(te))
    # FPNN (orthogonal-series PNN) and PNN with clustering on the same split
    out["std"] = ref.cls_std()
    for scale in (1.0, 0.33):
        lab, a, J = ref.cls_fpnn(0, len(te), scale=scale, bf=True, coefficients=True)
        out["fpnn_%g_label" % scale], out["fpnn_%g_J" % scale] = lab, np.int32(J)
        out["fpnn_%g_a_digest" % scale] = np.array([a.sum(), np.abs(a).sum(), a[::97].sum()])
        out["fpnn_%g_seq_label" % scale] = ref.cls_fpnn(0, len(te), scale=scale, bf=False, output_ratio=0.9)
    out["pnn_clustered_label"], out["pnn_clustered_medoids"] = ref.cls_pnn_clustered(4, 0, len(te))
    # second split with a class-independent first chunk so predict_sequentional's pruning departs from predict_bf
    rows2 = rows.copy()
    rows2[:, :32] = np.random.default_rng(11).normal(0, 0.04, size=(len(rows2), 32))
    rows2 /= np.linalg.norm(rows2, axis=1, keepdims=True)
    tr2, trl2, te2, avg2 = ref.cls_setup(rows2, labels, 7, 12, seed=5)
    out.update(rows2=rows2, train_idx2=tr2, train_labels2=trl2, test_idx2=te2, avg2=avg2)
    out["pnn_seq_label2"] = ref.cls_pnn_seq(0, len(te2))
    out["pnn_label2"] = ref.cls_pnn(0, len(te2))[0]
    np.savez_compressed(os.path.join(HERE, "ref_classification.npz"), **out)
    print("classification ->", len(out), "arrays")
    # three-way-decision classifiers (ImageTesting.cpp), the verbatim classes on one fixed 256-d split per metric
    port = oracle_py.Port()
    out = {}
    for metric in ("l2", "chi2", "kl"):
        ref = oracle_py.Ref(metric)
        g, gl, q, ql = synth.make_split(180, 60, 256, 6, metric, sigma=2.0, seed=77)
        # inputs are stored as float16 (exactly representable, small file) and normalised with the loader restatement
        g, q = g.astype(np.float16), q.astype(np.float16)
        out["%s_gallery_f16" % metric], out["%s_queries_f16" % metric] = g, q
        out["%s_gallery_labels" % metric] = gl
        gn, qn = port.normalize_rows(metric, g.astype(np.float32)), port.normalize_rows(metric, q.astype(np.float32))
        for fc, th in TWD_PROPOSED:
            out["%s_prop_%d_%g_class" % (metric, fc, th)], out["%s_prop_%d_%g_unrel" % (metric, fc, th)] = ref.twd("proposed", gn, gl, 6, qn, fc, th)
        for kind, th in TWD_CONVENTIONAL:
            out["%s_conv_%s_%g_class" % (metric, kind, th)], out["%s_conv_%s_%g_unrel" % (metric, kind, th)] = \
                ref.twd("conventional", gn, gl, 6, qn, 64, th, kind)
    np.savez_compressed(os.path.join(HERE, "ref_twd.npz"), **out)
    print("twd ->", len(out), "arrays")


if __name__ == "__main__":
    main()
