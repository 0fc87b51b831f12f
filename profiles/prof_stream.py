"""Latency mode: Q (<= 8) queries against an N x D gallery, one streaming pass; reports achieved HBM GB/s of the
streaming kernel (algorithmic bytes = N*D*4 per launch) from CUDA events around the launch."""
import importlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fir_b200
synth = importlib.import_module("fast-image-recognition_b200.synth")
metric = sys.argv[1] if len(sys.argv) > 1 else "chi2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
d = int(sys.argv[3]) if len(sys.argv) > 3 else 1280
dev = torch.device("cuda", 0)
g, gl, q, ql = synth.make_split_device(n, 8, d, 1000, metric, device=dev)
fir_b200.normalize_rows(g, metric); fir_b200.normalize_rows(q, metric)
gal = fir_b200.Gallery(g, gl, metric, stream=torch.cuda.current_stream().cuda_stream)
del g
peak = 6545.9
if os.path.exists("MEASURED_PEAKS.json"):
    peak = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", peak)
for nq in (1, 2, 4, 8):
    gal.profile(True)
    for it in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        sc, lab = gal.pnn_scores(q[:nq], 2e-5)
        idx, dd = gal.search(q[:nq], k=1, path=fir_b200.PATH_EXACT)
        torch.cuda.synchronize(); t = time.perf_counter() - t0
    ms, cnt = gal.profile_read(4)
    per = ms / cnt
    gbs = n * d * 4 / (per * 1e-3) / 1e9
    print(json.dumps({"metric": metric, "n": n, "d": d, "queries": nq, "stream_kernel_ms": per, "achieved_GBps": gbs, "hbm_peak_GBps": peak,
                      "frac": gbs / peak, "pnn_plus_top1_ms": 1e3 * t, "evals_per_s": 2 * nq * n / t}))
