"""Live phase breakdown of the tensor path (CUDA events around each phase on the launching stream, no profiler attached):
C2 workload by default.  Event pairs add a little launch gap, so the sum is slightly above an unprofiled step."""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fir_b200
synth = importlib.import_module("fast-image-recognition_b200.synth")
k = int(sys.argv[1]) if len(sys.argv) > 1 else 10
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 10_000
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 20
dev = torch.device("cuda", 0)
if n > 2_000_000:      # the bench's generator (counter-based, on the device): a 10M-row gallery takes minutes to build on the host
    g_dev, gl_dev = synth.synth_rows_device(synth.ROLE_GALLERY, 0, n, n, 512, 1000, 0x5EED0000, device=dev)
    q_dev, _ = synth.synth_rows_device(synth.ROLE_QUERY, 0, nq, nq, 512, 1000, 0x5EED0000, device=dev)
else:
    g, gl, q, ql = synth.make_split(n, nq, 512, 1000, "l2")
    g_dev, q_dev, gl_dev = torch.from_numpy(g).to(dev), torch.from_numpy(q).to(dev), torch.from_numpy(gl).to(dev)
fir_b200.normalize_rows(g_dev, "l2"); fir_b200.normalize_rows(q_dev, "l2")
gal = fir_b200.Gallery(g_dev, gl_dev, "l2", stream=torch.cuda.current_stream().cuda_stream)
del g_dev
for _ in range(3): gal.search(q_dev, k=k)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
# unprofiled step time first
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tot = 0.0
for _ in range(iters):
    flush.zero_(); a.record(); gal.search(q_dev, k=k); b.record(); torch.cuda.synchronize(); tot += a.elapsed_time(b)
plain = tot / iters
gal.profile(True)
tot = 0.0
for _ in range(iters):
    flush.zero_(); a.record(); gal.search(q_dev, k=k); b.record(); torch.cuda.synchronize(); tot += a.elapsed_time(b)
names = {6: "pack_queries", 12: "seed_sample_pass", 0: "candidates_pass1", 7: "prune", 8: "rerank", 9: "select", 10: "second_pass_total", 3: "candidates_pass2", 11: "exact_rerun_windows"}
out = {"k": k, "gallery": n, "queries": nq, "step_ms_unprofiled": plain, "step_ms_profiled": tot / iters, "phases_ms": {}}
for kind, name in names.items():
    ms, cnt = gal.profile_read(kind)
    out["phases_ms"][name] = ms / iters
out["sum_top_level_ms"] = sum(v for kname, v in out["phases_ms"].items() if kname != "candidates_pass2")
out["stats"] = gal.stats()
print(json.dumps(out))
