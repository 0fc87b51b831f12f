import sys, time, numpy as np
sys.path.insert(0, "/root/repo")
import fir_b200
r = np.random.default_rng(0)
C, per, d, nq = 101, 5, 1536, 5647
tr = r.normal(size=(C * per, d)); tr /= np.linalg.norm(tr, axis=1, keepdims=True)
lab = np.repeat(np.arange(C), per).astype(np.int32)
te = r.normal(size=(nq, d)); te /= np.linalg.norm(te, axis=1, keepdims=True)
avg = tr.mean(axis=0)
clf = fir_b200.Classifier(tr, lab, C, avg)
for i in range(3):
    t0 = time.perf_counter(); clf.pnn(te, scores=False); print("pnn 505 rows: %.2f ms" % (1e3 * (time.perf_counter() - t0)))
for i in range(2):
    t0 = time.perf_counter(); clf.knn(te, 1); print("knn: %.2f ms" % (1e3 * (time.perf_counter() - t0)))
