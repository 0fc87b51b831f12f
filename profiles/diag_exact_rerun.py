"""Cost of the exact CUDA-core re-run when ONE query cannot be certified by either tensor pass (here: 5000 gallery rows are
exact copies of query 0, so its top-10 is a mass tie): large gallery, device-resident queries."""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fir_b200
synth = importlib.import_module("fast-image-recognition_b200.synth")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
dev = torch.device("cuda", 0)
g, gl, q, ql = synth.make_split_device(n, nq, 512, 1000, "l2", device=dev)
fir_b200.normalize_rows(g, "l2"); fir_b200.normalize_rows(q, "l2")
out = {"gallery": n, "queries": nq}
for tag, dup in (("no_ties", 0), ("mass_tie", 5000)):
    if dup:
        g[1000:1000 + dup] = q[0]
    gal = fir_b200.Gallery(g, gl, "l2", stream=torch.cuda.current_stream().cuda_stream)
    for _ in range(2): idx, dist = gal.search(q, k=10)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): idx, dist = gal.search(q, k=10)
    b.record(); torch.cuda.synchronize()
    st = gal.stats()
    out[tag] = {"ms": a.elapsed_time(b) / 3, "flagged_pass1": st["n_fallback"], "exact_reruns": st["reserved"],
                "query0_top": idx[0, :3].tolist(), "query0_dist": dist[0, :3].tolist()}
    gal.close()
print(json.dumps(out))
