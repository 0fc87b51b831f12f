"""chi2/KL batched top-k: approximate tiles + exact rerank (PATH_APPROX) against the exact tile kernel (PATH_EXACT)."""
import importlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fir_b200
synth = importlib.import_module("fast-image-recognition_b200.synth")
metric = sys.argv[1] if len(sys.argv) > 1 else "chi2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
d = int(sys.argv[4]) if len(sys.argv) > 4 else 1280
k = int(sys.argv[5]) if len(sys.argv) > 5 else 10
dev = torch.device("cuda", 0)
g, gl, q, ql = synth.make_split_device(n, nq, d, 1000, metric, device=dev)
fir_b200.normalize_rows(g, metric); fir_b200.normalize_rows(q, metric)
gal = fir_b200.Gallery(g, gl, metric, stream=torch.cuda.current_stream().cuda_stream)
res = {}
for name, path in (("approx", fir_b200.PATH_APPROX), ("exact", fir_b200.PATH_EXACT)):
    for it in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        idx, dd = gal.search(q, k=k, path=path)
        torch.cuda.synchronize(); t = time.perf_counter() - t0
    res[name] = (idx, dd, t, gal.stats())
same = bool(torch.equal(res["approx"][0], res["exact"][0]) and torch.equal(res["approx"][1].view(torch.int32), res["exact"][1].view(torch.int32)))
print(json.dumps({"metric": metric, "n": n, "d": d, "queries": nq, "k": k, "approx_ms": 1e3 * res["approx"][2], "exact_ms": 1e3 * res["exact"][2],
                  "approx_evals_per_s": nq * n / res["approx"][2], "exact_evals_per_s": nq * n / res["exact"][2],
                  "approx_elem_per_s": nq * n * d / res["approx"][2], "identical_results": same, "approx_stats": res["approx"][3]}))
