"""Robustness probe: queries unrelated to the gallery's classes (all distances ~ 2, nearest neighbours in the extreme tail)."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, fir_b200
synth = importlib.import_module("fast-image-recognition_b200.synth")
dev = torch.device("cuda", 0)
g, gl, _, _ = synth.make_split(100000, 8, 512, 1000, "l2", seed=0)
_, _, q, ql = synth.make_split(8, 10000, 512, 1000, "l2", seed=1)      # other centroids
gd, qd = torch.from_numpy(g).to(dev), torch.from_numpy(q).to(dev)
fir_b200.normalize_rows(gd); fir_b200.normalize_rows(qd)
gal = fir_b200.Gallery(gd, torch.from_numpy(gl).to(dev), "l2")
gal.profile(True)
for k in (1, 10):
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        idx, dd = gal.search(qd, k=k); torch.cuda.synchronize()
        t = time.perf_counter() - t0
    print("k=%d: %.3f ms  stats %s" % (k, 1e3 * t, gal.stats()))
print("candidates pass1 ms/launches", gal.profile_read(0), "pass2", gal.profile_read(3), "exact", gal.profile_read(1))
