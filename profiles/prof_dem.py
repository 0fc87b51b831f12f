"""Profiling driver for directed enumeration: 1M x 512 gallery, one search of nq queries (used under ncu's launch list)."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fir_b200
synth = importlib.import_module("fast-image-recognition_b200.synth")
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
ratio = float(sys.argv[3]) if len(sys.argv) > 3 else 0.05
dev = torch.device("cuda", 0)
g, gl, q, ql = synth.make_split_device(n, nq, 512, 10_000, "l2", device=dev)
fir_b200.normalize_rows(g, "l2"); fir_b200.normalize_rows(q, "l2")
gal = fir_b200.Gallery(g, gl, "l2", stream=torch.cuda.current_stream().cuda_stream)
dem = fir_b200.Dem(gal, pivot0=12345, max_chain=64)
for it in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    idx, dist, below, evals = dem.search(q, int(ratio * n))
    torch.cuda.synchronize(); print("search %d: %.2f ms  (%.1f us/query), mean evals %.1f" % (it, 1e3 * (time.perf_counter() - t0), 1e6 * (time.perf_counter() - t0) / nq, evals.float().mean().item()))
