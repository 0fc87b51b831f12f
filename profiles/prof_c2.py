"""Profiling driver: the C2 workload (100k x 512 gallery, 10k queries, L2 top-k), a few searches, nothing else.
Used under `ncu` (launch list and --set full captures) — see profiles/README.md.  Numbers printed by a run under
ncu are not bench values."""
import importlib
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fir_b200

synth = importlib.import_module("fast-image-recognition_b200.synth")
k = int(sys.argv[1]) if len(sys.argv) > 1 else 10
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n = int(sys.argv[3]) if len(sys.argv) > 3 else 100_000
nq = int(sys.argv[4]) if len(sys.argv) > 4 else 10_000
g, gl, q, ql = synth.make_split(n, nq, 512, 1000, "l2")
dev = torch.device("cuda", 0)
g_dev, q_dev = torch.from_numpy(g).to(dev), torch.from_numpy(q).to(dev)
fir_b200.normalize_rows(g_dev, "l2")
fir_b200.normalize_rows(q_dev, "l2")
gal = fir_b200.Gallery(g_dev, torch.from_numpy(gl).to(dev), "l2", stream=torch.cuda.current_stream().cuda_stream)
gal.profile(True)
for it in range(iters):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    idx, dist = gal.search(q_dev, k=k)
    torch.cuda.synchronize()
    print("iter %d: %.3f ms, stats %s" % (it, 1e3 * (time.perf_counter() - t0), gal.stats()))
print("candidates kernel: total %.3f ms over %d launches" % gal.profile_read(0))
print("exact tiles kernel: total %.3f ms over %d launches" % gal.profile_read(1))
acc = (torch.from_numpy(gl).to(dev)[idx[:, 0].long()] == torch.from_numpy(ql).to(dev)).float().mean().item()
print("label accuracy %.4f" % acc)
