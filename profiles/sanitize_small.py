"""Smallest invocation of every kernel family, for compute-sanitizer (one tool per gpurun call)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fir_b200
synth = importlib.import_module("fast-image-recognition_b200.synth")
for metric in ("l2", "chi2", "kl"):
    g, gl, q, ql = synth.make_split(700, 70, 96, 6, metric, seed=1)
    g, q = fir_b200.normalize_rows(g, metric), fir_b200.normalize_rows(q, metric)
    gal = fir_b200.Gallery(g, gl, metric)
    gal.search(q, k=5, path=fir_b200.PATH_EXACT)
    gal.class_min(q); gal.pnn_scores(q, 1e-3)
    gal.distances(q, np.tile(np.arange(40, dtype=np.int32), (70, 1)))
    gal.close()
g, gl, q, ql = synth.make_split(4500, 8, 128, 6, "chi2", seed=2)
g, q = fir_b200.normalize_rows(g, "chi2"), fir_b200.normalize_rows(q, "chi2")
gal = fir_b200.Gallery(g, gl, "chi2"); gal.search(q[:3], k=4, path=fir_b200.PATH_EXACT); gal.pnn_scores(q[:3], 1e-3); gal.class_min(q[:5]); gal.close()
g, gl, q, ql = synth.make_split(900, 300, 128, 6, "l2", seed=3)
g, q = fir_b200.normalize_rows(g, "l2"), fir_b200.normalize_rows(q, "l2")
gal = fir_b200.Gallery(g, gl, "l2")
for k in (1, 10, 20):
    gal.search(q, k=k, path=fir_b200.PATH_TENSOR)
dem = fir_b200.Dem(gal, pivot0=5)
dem.search(q, 200)
low = fir_b200.Dem(gal, state=(dem.pivots, dem.P, 1e-9)); low.search(q[:40], 300); low.close(); dem.close(); gal.close()
rows = np.concatenate([g, q]).astype(np.float64); lab = np.concatenate([gl, ql]).astype(np.int32)
order = np.argsort(lab[:900], kind="stable")
clf = fir_b200.Classifier(rows[order], lab[order], 6, rows[order].mean(0)); clf.knn(rows[900:], 3); clf.pnn(rows[900:]); clf.close()
print("sanitize_small ok")
