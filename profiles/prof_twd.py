"""Sequential three-way-decision classifiers (fir_twd_*) and the prefix brute force they are compared with in
ImageTesting.cpp:525-534, timed on the GPU (device-resident queries, CUDA events) with the reference's own classes timed
beside them on a bounded sample of the same queries (oracle/_ref, one host thread — the reference is single-threaded)."""
import importlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, fir_b200
synth = importlib.import_module("fast-image-recognition_b200.synth")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3030          # Caltech-101 split of the reference: 101 classes x 30 (db.h:11, db_features.cpp:133)
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 5647
d = int(sys.argv[3]) if len(sys.argv) > 3 else 1536          # FEATURES_COUNT, db.h:86
n_classes = int(sys.argv[4]) if len(sys.argv) > 4 else 101
cpu_sample = int(sys.argv[5]) if len(sys.argv) > 5 else 64
sigma = float(sys.argv[6]) if len(sys.argv) > 6 else 2.0
dev = torch.device("cuda", 0)
g, gl, q, ql = synth.make_split_device(n, nq, d, n_classes, "l2", device=dev, sigma=sigma)
fir_b200.normalize_rows(g, "l2"); fir_b200.normalize_rows(q, "l2")
gal = fir_b200.Gallery(g, gl, "l2", stream=torch.cuda.current_stream().cuda_stream)
gal.set_num_classes(n_classes)

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): out = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out

cases = [("BF 256", lambda: gal.search(q, k=1, max_features=256)),
         ("BF 64", lambda: gal.search(q, k=1, max_features=64)),
         ("TWD posteriors 0.24", lambda: gal.twd_conventional(q, "posteriors", 0.24)),
         ("TWD diff 0.003", lambda: gal.twd_conventional(q, "diff", 0.003)),
         ("TWD ratio 0.7", lambda: gal.twd_conventional(q, "ratio", 0.7)),
         ("Proposed TWD 32 0.7", lambda: gal.twd_proposed(q, 32, 0.7)),
         ("Proposed TWD 64 0.7", lambda: gal.twd_proposed(q, 64, 0.7))]
rows = []
ref = None
try:
    from oracle import oracle_py
    ref = oracle_py.Ref("l2") if oracle_py.Ref.available("l2") else None
except Exception:
    ref = None
gh, qh, glh = g.cpu().numpy(), q[:cpu_sample].cpu().numpy(), gl.cpu().numpy()
refcalls = {"BF 256": ("bf", 256, 0, "diff"), "BF 64": ("bf", 64, 0, "diff"), "TWD posteriors 0.24": ("conventional", 64, 0.24, "posteriors"),
            "TWD diff 0.003": ("conventional", 64, 0.003, "diff"), "TWD ratio 0.7": ("conventional", 64, 0.7, "ratio"),
            "Proposed TWD 32 0.7": ("proposed", 32, 0.7, "diff"), "Proposed TWD 64 0.7": ("proposed", 64, 0.7, "diff")}
for name, fn in cases:
    ms, out = timed(fn)
    row = {"classifier": name, "gpu_ms": ms, "gpu_us_per_query": 1e3 * ms / nq, "queries_per_s": nq / (ms * 1e-3)}
    if name.startswith("BF"):
        lab = gl[out[0][:, 0].long()]
    else:
        lab = out[1]; row["unreliable_pct"] = 100.0 * float(out[2].float().mean())
    row["accuracy_pct"] = 100.0 * float((lab == ql).float().mean())
    if ref is not None:
        kind, fc, th, tt = refcalls[name]
        t0 = time.perf_counter(); rc, ru = ref.twd(kind, gh, glh, n_classes, qh, fc, th, tt); t = time.perf_counter() - t0
        row["cpu_ref_us_per_query"] = 1e6 * t / cpu_sample
        row["cpu_ref_sample_queries"] = cpu_sample
        row["matches_reference_on_sample"] = bool(np.array_equal(rc, lab[:cpu_sample].cpu().numpy()))
        row["speedup_vs_cpu_ref"] = row["cpu_ref_us_per_query"] / row["gpu_us_per_query"]
    rows.append(row)
print(json.dumps({"workload": "twd", "gallery": n, "queries": nq, "d": d, "classes": n_classes, "sigma": sigma, "rows": rows}))

# fp64 PNN: predict_bf against predict_sequentional (classification.cpp:188-295) on the same split, C port timed beside it
if n <= 20000:
    from oracle import oracle_py
    port = oracle_py.Port()
    order = np.argsort(glh, kind="stable")
    tr = gh[order].astype(np.float64); trl = glh[order].astype(np.int32)
    te = q.cpu().numpy().astype(np.float64)
    avg = tr.mean(axis=0)
    clf = fir_b200.Classifier(tr, trl, n_classes, avg)
    res = []
    for name, fn, pf in (("PNN", lambda: clf.pnn(te, scores=False)[0], lambda x: port.pnn(tr, trl, n_classes, avg, x)[1]),
                         ("PNN (seq)", lambda: clf.pnn_sequential(te), lambda x: port.pnn_seq(tr, trl, n_classes, avg, x))):
        fn(); t0 = time.perf_counter(); lab = fn(); t = time.perf_counter() - t0          # host fp64 queries in, labels out (second call: workspace in place)
        t1 = time.perf_counter(); pl = pf(te[:cpu_sample]); tc = time.perf_counter() - t1
        res.append({"classifier": name, "gpu_e2e_ms": 1e3 * t, "gpu_us_per_query": 1e6 * t / nq, "cpu_port_us_per_query": 1e6 * tc / cpu_sample,
                    "matches_port_on_sample": bool(np.array_equal(pl, lab[:cpu_sample])), "accuracy_pct": 100.0 * float((lab == ql.cpu().numpy()).mean())})
    print(json.dumps({"workload": "pnn_fp64", "train": n, "queries": nq, "d": d, "classes": n_classes, "rows": res}))
    # FPNN (orthogonal-series PNN) and PNN with clustering (classification.cpp:618-791, 311-428)
    sd = tr.std(axis=0, ddof=1)
    res = []
    fir_b200.Fpnn(tr, trl, n_classes, avg, sd, 1.0).close()           # first call loads the kernels (lazy module loading); time the second
    t0 = time.perf_counter(); f = fir_b200.Fpnn(tr, trl, n_classes, avg, sd, 1.0); t_train = time.perf_counter() - t0
    t0 = time.perf_counter(); pa, pJ = port.fpnn_train(tr, trl, n_classes, avg, sd, 1.0); t_ptrain = time.perf_counter() - t0
    for name, seq in (("FPNN", False), ("FPNN (seq)", True)):
        f.predict(te, sequential=seq); t0 = time.perf_counter(); lab = f.predict(te, sequential=seq); t = time.perf_counter() - t0
        t1 = time.perf_counter(); pl = port.fpnn_predict(pa, pJ, n_classes, avg, sd, te[:cpu_sample], 1.0, sequential=seq); tc = time.perf_counter() - t1
        res.append({"classifier": name, "J": f.J, "gpu_train_ms": 1e3 * t_train, "cpu_port_train_ms": 1e3 * t_ptrain, "gpu_e2e_ms": 1e3 * t,
                    "gpu_us_per_query": 1e6 * t / nq, "cpu_port_us_per_query": 1e6 * tc / cpu_sample,
                    "matches_port_on_sample": bool(np.array_equal(pl, lab[:cpu_sample])), "accuracy_pct": 100.0 * float((lab == ql.cpu().numpy()).mean())})
    clusters = 5
    fir_b200.kmedoids_select(tr, trl, n_classes, clusters)
    t0 = time.perf_counter(); sel = fir_b200.kmedoids_select(tr, trl, n_classes, clusters); t_sel = time.perf_counter() - t0
    t0 = time.perf_counter(); psel = port.kmedoids(tr, trl, n_classes, clusters); t_psel = time.perf_counter() - t0
    red = fir_b200.Classifier(tr[sel], trl[sel], n_classes, avg); red.set_total(len(tr))
    red.pnn(te, scores=False); t0 = time.perf_counter(); lab = red.pnn(te, scores=False)[0]; t = time.perf_counter() - t0
    res.append({"classifier": "PNN with clustering, %d" % clusters, "kept_rows": int(len(sel)), "gpu_kmedoids_ms": 1e3 * t_sel, "cpu_port_kmedoids_ms": 1e3 * t_psel,
                "medoids_match_port": bool(np.array_equal(sel, psel)), "gpu_e2e_ms": 1e3 * t, "gpu_us_per_query": 1e6 * t / nq,
                "accuracy_pct": 100.0 * float((lab == ql.cpu().numpy()).mean())})
    print(json.dumps({"workload": "fpnn_and_clustering", "train": n, "queries": nq, "d": d, "classes": n_classes, "rows": res}))
