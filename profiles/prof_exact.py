"""Profiling driver for the exact CUDA-core tile kernel: metric in {l2,chi2,kl}, N x D gallery, Q queries, top-1 + PNN."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fir_b200
synth = importlib.import_module("fast-image-recognition_b200.synth")
metric = sys.argv[1] if len(sys.argv) > 1 else "chi2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 512
d = int(sys.argv[4]) if len(sys.argv) > 4 else 1280
dev = torch.device("cuda", 0)
g, gl, q, ql = synth.make_split_device(n, nq, d, 1000, metric, device=dev)
fir_b200.normalize_rows(g, metric); fir_b200.normalize_rows(q, metric)
gal = fir_b200.Gallery(g, gl, metric, stream=torch.cuda.current_stream().cuda_stream)
gal.profile(True)
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    idx, dd = gal.search(q, k=1, path=fir_b200.PATH_EXACT)
    torch.cuda.synchronize(); t = time.perf_counter() - t0
    print("search %.3f ms  %.3e evals/s  %.3e elem/s" % (1e3 * t, nq * n / t, nq * n * d / t))
print("exact tile kernel: %.3f ms over %d launches" % gal.profile_read(1))
