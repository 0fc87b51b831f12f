"""Large-shape soak (GPU only, no oracle): random BASELINE-scale shapes through the fast paths — tensor-core top-k (full rounds,
balanced and phased remainders, prefixes), tensor class minima, approximate chi2 / KL top-k, latency mode — each compared bit
for bit with the exact CUDA-core kernels on a query subset (the exact kernels are pinned to the reference by the test suite).
usage: python profiles/soak_large.py [n_cases] [seed]"""
import importlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, fir_b200 as fir
synth = importlib.import_module("fast-image-recognition_b200.synth")
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 12
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
dev = torch.device("cuda", 0)


def same(a, b):
    return bool(torch.equal(a[0], b[0]) and torch.equal(a[1].view(torch.int32), b[1].view(torch.int32)))


fails = 0
for case in range(n_cases):
    metric = ["l2", "l2", "l2", "chi2", "kl"][int(rng.integers(0, 5))]
    if metric == "l2":
        n = int(rng.choice([60_000, 250_001, 1_000_003, 3_000_000]))
        nq = int(rng.choice([300, 5_000, 19_200, 24_333, 60_001]))
        d = int(rng.choice([64, 200, 512, 513, 1024]))
    else:
        n = int(rng.choice([50_000, 200_003, 600_000]))
        nq = int(rng.choice([40, 700, 2_000]))
        d = int(rng.choice([96, 500, 1280]))
    c = int(rng.choice([7, 300, 1000]))
    k = int(rng.choice([1, 5, 10, 28]))
    t0 = time.perf_counter()
    g, gl = synth.synth_rows_device(synth.ROLE_GALLERY, 0, n, n, d, c, 1000 + case, relu=metric != "l2", device=dev)
    q, ql = synth.synth_rows_device(synth.ROLE_QUERY, 0, nq, nq, d, c, 1000 + case, relu=metric != "l2", device=dev)
    fir.normalize_rows(g, metric); fir.normalize_rows(q, metric)
    if rng.random() < 0.5:
        g[n // 2: n // 2 + 3] = g[5]                                       # a few exact ties
    gal = fir.Gallery(g, gl, metric)
    del g
    sel = torch.from_numpy(np.sort(rng.choice(nq, min(nq, 192), replace=False))).to(dev)
    qs = q[sel].contiguous()
    res = {"case": case, "metric": metric, "n": n, "nq": nq, "d": d, "classes": c, "k": k, "checks": {}}
    ex = gal.search(qs, k=k, path=fir.PATH_EXACT)
    fast = gal.search(q, k=k)                                             # PATH_AUTO: tensor (L2) / approximate (chi2, KL) at these sizes
    res["path"] = gal.stats()["path_used"]
    res["checks"]["topk"] = same((fast[0][sel], fast[1][sel]), ex)
    if metric == "l2":
        mf = int(rng.integers(8, d))
        exp = gal.search(qs, k=k, max_features=mf, path=fir.PATH_EXACT)
        fp = gal.search(q, k=k, max_features=mf, path=fir.PATH_TENSOR)
        res["checks"]["prefix_%d" % mf] = same((fp[0][sel], fp[1][sel]), exp)
        if nq * n >= 1 << 24 and nq * c < 2 ** 31:
            mn, arg = gal.class_min(q)                                    # tensor path
            used = gal.stats()["path_used"]
            mne, arge = gal.class_min(qs[:8])                             # <= 8 queries: latency mode + exact arithmetic
            res["checks"]["class_min(path %d)" % used] = bool(torch.equal(mn[sel[:8]].view(torch.int32), mne.view(torch.int32)) and torch.equal(arg[sel[:8]], arge))
    for m in (3, 6):                                                      # latency mode: query counts that are not powers of two
        lm = gal.search(qs[:m], k=min(k, 16), path=fir.PATH_EXACT)
        exm = (ex[0][:m, :min(k, 16)], ex[1][:m, :min(k, 16)])
        res["checks"]["latency_%d" % m] = same(lm, exm)
    torch.cuda.synchronize()
    res["seconds"] = time.perf_counter() - t0
    ok = all(res["checks"].values())
    fails += not ok
    print(json.dumps(res), flush=True)
    gal.close()
    del q, fast, ex
    torch.cuda.empty_cache()
print(json.dumps({"cases": n_cases, "failed": fails}))
sys.exit(1 if fails else 0)
