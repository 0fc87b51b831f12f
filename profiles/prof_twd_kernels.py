"""Profiling driver: ProposedTWD (32-dim chunks) and ConventionalTWD on a 100k x 512 gallery, 10k device-resident queries."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fir_b200
synth = importlib.import_module("fast-image-recognition_b200.synth")
dev = torch.device("cuda", 0)
g, gl, q, ql = synth.make_split_device(100_000, 10_000, 512, 1000, "l2", device=dev, sigma=2.0)
fir_b200.normalize_rows(g, "l2"); fir_b200.normalize_rows(q, "l2")
gal = fir_b200.Gallery(g, gl, "l2", stream=torch.cuda.current_stream().cuda_stream)
for name, fn in (("proposed 32", lambda: gal.twd_proposed(q, 32, 0.7)), ("conventional ratio", lambda: gal.twd_conventional(q, "ratio", 0.7))):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize()
    print("%s: %.2f ms, unreliable %.1f%%" % (name, 1e3 * (time.perf_counter() - t0), 100 * out[2].float().mean().item()))
