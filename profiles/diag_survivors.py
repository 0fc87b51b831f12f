"""How many candidates per query survive the prune (and so get an exact rerank) on the C2 workload."""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, fir_b200
synth = importlib.import_module("fast-image-recognition_b200.synth")
k = int(sys.argv[1]) if len(sys.argv) > 1 else 10
g, gl, q, ql = synth.make_split(100_000, 10_000, 512, 1000, "l2")
dev = torch.device("cuda", 0)
g_dev, q_dev = torch.from_numpy(g).to(dev), torch.from_numpy(q).to(dev)
fir_b200.normalize_rows(g_dev, "l2"); fir_b200.normalize_rows(q_dev, "l2")
gal = fir_b200.Gallery(g_dev, torch.from_numpy(gl).to(dev), "l2", stream=torch.cuda.current_stream().cuda_stream)
idx, dist = gal.search(q_dev, k=k)
ci, ca, ce = gal.debug_candidates(10_000)
alive = (ci >= 0).sum(axis=1)
st = gal.stats()
kth = dist[:, k - 1].cpu().numpy() * 512
print(json.dumps({"k": k, "slots_x_R": int(ci.shape[1]), "survivors_mean": float(alive.mean()), "survivors_p50": float(np.median(alive)),
                  "survivors_p99": float(np.percentile(alive, 99)), "err_bound_E": st["approx_err_bound"], "kth_sqdist_mean": float(kth.mean()),
                  "kth_sqdist_p10": float(np.percentile(kth, 10))}))
