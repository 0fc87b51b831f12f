"""Per-class nearest neighbour (fir_class_min, L2): tensor-core passes + exact rerank against the exact CUDA-core tiles.
usage: python profiles/prof_classmin.py [n] [nq] [d] [classes]"""
import importlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fir_b200
synth = importlib.import_module("fast-image-recognition_b200.synth")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
d = int(sys.argv[3]) if len(sys.argv) > 3 else 512
c = int(sys.argv[4]) if len(sys.argv) > 4 else 1000
g, gl = synth.synth_rows_device(0, 0, n, n, d, c, seed=0x5EED0000)
q, ql = synth.synth_rows_device(1, 0, nq, nq, d, c, seed=0x5EED0000)
fir_b200.normalize_rows(g, "l2"); fir_b200.normalize_rows(q, "l2")
gal = fir_b200.Gallery(g, gl, "l2", stream=torch.cuda.current_stream().cuda_stream)
gal.set_num_classes(c)
if True:                                                # FIR_CLASSMIN_TENSOR=0 in the environment times the exact CUDA-core tiles instead
    gal.profile(True)
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        mn, arg = gal.class_min(q)
        torch.cuda.synchronize(); t = time.perf_counter() - t0
    k1 = gal.profile_read(0); k2 = gal.profile_read(3); k3 = gal.profile_read(8)
    res = {"n": n, "queries": nq, "d": d, "classes": c, "ms": 1e3 * t, "evals_per_s": nq * n / t, "path": gal.stats()["path_used"],
           "pass1_ms": k1[0] / max(k1[1], 1), "pass2_ms": k2[0] / max(k2[1], 1), "rerank_ms": k3[0] / max(k3[1], 1), "overflow_cells": gal.stats()["n_fallback"]}
print(json.dumps(res))
