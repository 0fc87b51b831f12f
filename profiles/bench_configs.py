"""Secondary workloads of BASELINE.json (configs[2..4]) — one JSON line each.  bench.py stays the headline (configs[1]).
usage: python profiles/bench_configs.py c3 [chi2|kl] [nq]   |   c4 [nq]   |   c5 [n] [nq] [k]
All inputs are generated on the GPU (synth.make_split_device), normalised by the loader kernel; timings are CUDA
events around the C-ABI call with device pointers."""
import importlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import fir_b200

synth = importlib.import_module("fast-image-recognition_b200.synth")
dev = torch.device("cuda", 0)


def timed(fn, iters=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / 1e3)
    return min(ts), out


def c3(metric="chi2", nq=1024, n=1_000_000, d=1280, classes=1000):
    g, gl, q, ql = synth.make_split_device(n, nq, d, classes, metric, device=dev)
    fir_b200.normalize_rows(g, metric)
    fir_b200.normalize_rows(q, metric)
    gal = fir_b200.Gallery(g, gl, metric, stream=torch.cuda.current_stream().cuda_stream)
    del g
    var = 2e-5
    t_pnn, (sc, lab) = timed(lambda: gal.pnn_scores(q, var), iters=2, warm=1)
    t_top, (idx, dd) = timed(lambda: gal.search(q, k=1), iters=2, warm=0)
    acc = float((lab == ql).float().mean().item())
    acc1 = float((gl[idx[:, 0].long()] == ql).float().mean().item())
    evals = float(nq) * n
    print(json.dumps({"workload": "C3 %s: %d x %d gallery, %d queries, PNN class scores + 1-NN" % (metric, n, d, nq), "metric": "distance_evals_per_s",
                      "pnn": {"value": evals / t_pnn, "ms": 1e3 * t_pnn, "label_accuracy": acc},
                      "top1": {"value": evals / t_top, "ms": 1e3 * t_top, "label_accuracy": acc1},
                      "element_ops_per_s": evals * d / t_top, "dtype": "f32 exact (CUDA cores)"}))


def c4(nq=10_000, n=1_000_000, d=512, classes=10_000, max_chain=64):
    g, gl, q, ql = synth.make_split_device(n, nq, d, classes, "l2", device=dev)
    fir_b200.normalize_rows(g, "l2")
    fir_b200.normalize_rows(q, "l2")
    gal = fir_b200.Gallery(g, gl, "l2", stream=torch.cuda.current_stream().cuda_stream)
    del g
    t0 = time.perf_counter()
    dem = fir_b200.Dem(gal, pivot0=12345, max_chain=max_chain)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0
    t_bf, (bi, bd) = timed(lambda: gal.search(q, k=1), iters=2, warm=1)
    rows = []
    for ratio in (0.025, 0.05, 0.1, 0.2):
        M = int(ratio * n)
        t, (idx, dist, below, evals) = timed(lambda: dem.search(q, M), iters=1, warm=0)
        rows.append({"ratio": ratio, "ms": 1e3 * t, "queries_per_s": nq / t, "recall_at_1_vs_bf": float((idx == bi[:, 0]).float().mean().item()),
                     "below_threshold_frac": float(below.float().mean().item()), "checked_percent": float(100.0 * evals.float().mean().item() / n)})
    print(json.dumps({"workload": "C4 DEM: %d x %d gallery, %d queries, %d pivots, chain %d rows" % (n, d, nq, dem.n_pivots, dem.chain_rows),
                      "build_s": t_build, "threshold": float(dem.threshold), "bf_ms": 1e3 * t_bf, "bf_queries_per_s": nq / t_bf, "dem": rows}))


def c5(n=10_000_000, nq=100_000, k=10, d=512, classes=1000):
    g, gl, q, ql = synth.make_split_device(n, nq, d, classes, "l2", device=dev)
    fir_b200.normalize_rows(g, "l2")
    fir_b200.normalize_rows(q, "l2")
    gal = fir_b200.Gallery(g, gl, "l2", stream=torch.cuda.current_stream().cuda_stream)
    del g
    torch.cuda.empty_cache()
    gal.profile(True)
    t, (idx, dd) = timed(lambda: gal.search(q, k=k), iters=2, warm=1)
    kms, kn = gal.profile_read(0)
    st = gal.stats()
    acc = float((gl[idx[:, 0].long()] == ql).float().mean().item())
    evals = float(nq) * n
    # the candidates kernel runs twice per search when a second pass exists; attribute the big launches only
    print(json.dumps({"workload": "C5: %d x %d gallery, %d queries, L2 top-%d, 1 GPU" % (n, d, nq, k), "metric": "distance_evals_per_s",
                      "value": evals / t, "ms": 1e3 * t, "queries_per_s": nq / t, "label_accuracy": acc, "stats": st,
                      "candidates_kernels_ms_total": kms, "candidates_launches": kn,
                      "tensor_tflops_whole_search": 2.0 * d * evals / t / 1e12}))


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "c3"
    args = sys.argv[2:]
    if which == "c3":
        c3(args[0] if args else "chi2", int(args[1]) if len(args) > 1 else 1024, int(args[2]) if len(args) > 2 else 1_000_000)
    elif which == "c4":
        c4(int(args[0]) if args else 10_000, int(args[1]) if len(args) > 1 else 1_000_000)
    else:
        c5(int(args[0]) if args else 10_000_000, int(args[1]) if len(args) > 1 else 100_000, int(args[2]) if len(args) > 2 else 10)
