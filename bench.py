#!/usr/bin/env python
"""bench.py — benchmark of the matching hot path on the configurations BASELINE.json names.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config c5|c2|c1|c3-chi2|c3-kl|c4] [--k 10]

  c5 (default)  10M x 512 gallery, 100k queries, Euclidean top-k — BASELINE.json configs[4]; at 1 GPU it is the largest
                single-GPU configuration, at --gpus N the SAME gallery is strong-sharded row-wise over the N GPUs
                (fir_shard_search_topk: local exact top-k, NCCL all-gather of packed (dist, idx) keys, k-way merge).
  c2            100k x 512 gallery, 10k queries (configs[1])          c1   Caltech-101-shaped split, D = 1536 (configs[0])
  c3-chi2/-kl   1M x 1280 gallery, chi-square / KL distances + PNN class scores (configs[2])
  c4            directed-enumeration ANN on a 1M x 512 gallery with the reference's full pivot chain (configs[3])
  cls           the fp64 kNN / PNN classifiers of classification.cpp at the Caltech split shape

A "step" is one pass of the hot path over one query batch.  Every line carries
  value      the metric with inputs resident in HBM, CUDA-event timed on the launching stream, max over ranks;
  e2e        the same through the C-ABI with HOST (pinned) buffers: H2D of the queries and D2H of the results inside the timed region;
  roofline   the dominant kernel, timed live with CUDA events around each of its launches, against MEASURED_PEAKS.json;
  parity     the GPU results of a query sample against the reference's own code (oracle/_ref, the unmodified reference compiled
             here) run over the FULL gallery on the host after the timed regions — at N > 1 it is the merged result that is checked;
  cpu_baseline (N = 1)  the same reference run, timed: the CPU number reported beside the GPU one;
  clocks     SM clocks / throttle reasons sampled during the timed regions.
Synthetic inputs come from a counter-based generator keyed by (seed, row, column) (csrc/synth_common.h): rank r materialises
rows [lo, hi) of the same gallery on its GPU, the host regenerates identical bits (tests/test_synth.py).
`--impl reference` times the reference's own CPU loop (oracle/_ref) on the host cores for the same config; no GPU is touched.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 0x5EED0000
CONFIGS = {
    "c5": dict(kind="bf", metric="l2", n=10_000_000, nq=100_000, d=512, classes=1000, parity_q=64, parity_topk_q=16,
               name="C5: synthetic 512-d L2-normalised embeddings, 10M gallery x 100k queries, Euclidean top-%(k)d (BASELINE.json configs[4])"),
    "c2": dict(kind="bf", metric="l2", n=100_000, nq=10_000, d=512, classes=1000, parity_q=1024, parity_topk_q=1024,
               name="C2: synthetic 512-d L2-normalised embeddings, 100k gallery x 10k queries, Euclidean top-%(k)d (BASELINE.json configs[1])"),
    "c1": dict(kind="bf", metric="l2", n=3030, nq=5647, d=1536, classes=101, parity_q=5647, parity_topk_q=512,
               name="C1: Caltech-101-shaped split (101 classes, 30 gallery rows per class = 3030, 5647 queries, D = 1536), Euclidean top-%(k)d "
                    "(BASELINE.json configs[0], synthetic features of the reference's shape)"),
    "c3-chi2": dict(kind="pnn", metric="chi2", n=1_000_000, nq=1024, d=1280, classes=1000, parity_q=16, var=2e-5,
                    name="C3 chi-square: synthetic 1280-d EfficientNet-style (ReLU, L1-normalised) features, 1M gallery x 1024 queries, "
                         "PNN class scores + labels (BASELINE.json configs[2])"),
    "c3-kl": dict(kind="pnn", metric="kl", n=1_000_000, nq=256, d=1280, classes=1000, parity_q=8, var=2e-5,
                  name="C3 KL: synthetic 1280-d EfficientNet-style (ReLU, L1-normalised) features, 1M gallery x 256 queries, "
                       "PNN class scores + labels (BASELINE.json configs[2])"),
    "cls": dict(kind="cls", metric="l2", n=3030, nq=5647, d=1536, classes=101, parity_q=256, K=3,
                name="fp64 classifiers of classification.cpp at the Caltech-101 split shape (3030 training rows, 5647 queries, D = 1536, 101 classes): "
                     "PNNClassifier::predict_bf (Parzen scores + labels), kNN (K = 3) reported beside it (BASELINE.json configs[0] data, SURVEY 8(a) a8/a9)"),
    "c4": dict(kind="dem", metric="l2", n=1_000_000, nq=10_000, d=512, classes=10_000, parity_q=64, ratio=0.05,
               name="C4: directed-enumeration ANN (32 pivots, FAR 0.01, the reference's full 0.015 N pivot chain) on a 1M x 512 gallery, "
                    "10k queries, imageCountToCheck = 0.05 N (BASELINE.json configs[3])"),
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    out = {"tensor_tflops_sustained": 1400.0, "tensor_tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}
    if os.path.exists(p):
        with open(p) as f:
            j = json.load(f)
        out = {"tensor_tflops_sustained": float(j.get("bf16_tflops_sustained", j.get("bf16_tflops", 1418.0))),
               "tensor_tflops_burst": float(j.get("bf16_tflops", 1687.2)), "hbm_gbs": float(j.get("hbm_gbs", 6545.9)),
               "source": "MEASURED_PEAKS.json"}
    p = os.path.join(ROOT, "profiles", "r2_peak_pipes.json")
    if os.path.exists(p):
        with open(p) as f:
            out["pipes"] = json.load(f)
    return out


def traffic_from_profiles(config, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel per launch from the committed `ncu --set full`
    capture of THIS config (profiles/r2_traffic.json); None when no capture of the config is recorded."""
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get(config, {}).get(kernel)
    return None


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons while the timed regions run: NVML every 2 ms when importable, else nvidia-smi."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index=0):
        super().__init__(daemon=True)
        self.gpu, self.sm, self.max_sm, self.reasons, self._stop_evt, self.source = gpu_index, [], 0.0, set(), threading.Event(), "nvidia-smi"
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].strip().isdigit() else gpu_index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml, self.source = pynvml, "nvml"
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        for name, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40)):
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        if not out:
            return
        r = [c.strip() for c in out.split(",")]
        self.sm.append(float(r[0]))
        self.max_sm = max(self.max_sm, float(r[1]))
        for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
            if v.lower().startswith("active"):
                self.reasons.add(name)

    def run(self):
        while not self._stop_evt.is_set():
            try:
                if self.nvml:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop_evt.wait(0.002 if self.nvml else 0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        busy = sorted(s for s in self.sm if s > 0)
        med = busy[len(busy) // 2] if busy else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_sm or None, "reasons": sorted(self.reasons), "samples": len(self.sm), "source": self.source}


# ------------------------------------------------------------------------------------------------------------------
# host side (reference arm, parity, cpu_baseline): test infrastructure under oracle/ — never on the product path
# ------------------------------------------------------------------------------------------------------------------
def host_gallery(cfg, threads):
    """The config's gallery + queries in host memory, normalised "as loaded by db_features" by the oracle's restatement of
    the loader (db_features.cpp:79-101) — what `--impl reference` matches against.  No GPU."""
    import importlib
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle_py
    synth = importlib.import_module("fast-image-recognition_b200.synth")
    port = oracle_py.Port()
    relu = cfg["metric"] != "l2"
    g = np.empty((cfg["n"], cfg["d"]), np.float32)
    step = 1 << 18
    gl = np.empty(cfg["n"], np.int32)
    for lo in range(0, cfg["n"], step):
        hi = min(cfg["n"], lo + step)
        _, gl[lo:hi] = synth.synth_rows_host(synth.ROLE_GALLERY, lo, hi - lo, cfg["n"], cfg["d"], cfg["classes"], SEED, relu=relu, out=g[lo:hi], threads=threads)
    q, ql = synth.synth_rows_host(synth.ROLE_QUERY, 0, cfg["nq"], cfg["nq"], cfg["d"], cfg["classes"], SEED, relu=relu, threads=threads)
    chunks = [(lo, min(cfg["n"], lo + (1 << 16))) for lo in range(0, cfg["n"], 1 << 16)]
    with ThreadPoolExecutor(max_workers=threads) as ex:                      # ctypes releases the GIL: row chunks in parallel
        list(ex.map(lambda c: port.normalize_rows(cfg["metric"], g[c[0]:c[1]]), chunks))
    port.normalize_rows(cfg["metric"], q)
    return g, gl, q, ql


def host_dem_state(g, gl, pivot0, rows, nthreads, far=0.01):
    """First `rows` steps of the DEM pivot chain on the host (ann.cpp:302-331): P[ii][j] = feature_distance(db[j], pivot),
    fp64 far-sums, arg-max with strict '>' (lowest j), per-row minimum other-class distance; threshold = getThreshold."""
    import numpy as np
    from oracle import oracle_py
    port = oracle_py.Port()
    n = g.shape[0]
    piv, P, mo = [int(pivot0)], np.empty((rows, n), np.float32), []
    far_sum = np.zeros(n, np.float64)
    for ii in range(rows):
        p = piv[ii]
        _, dd = port.topk("l2", g[p:p + 1], g, 1, nthreads=nthreads)          # lhs = gallery row j, rhs = the pivot (ann.cpp:309)
        P[ii] = dd[:, 0]
        other = gl != gl[p]
        mo.append(P[ii][other].min() if other.any() else np.float32(3.4e38))
        far_sum += P[ii].astype(np.float64)
        far_sum[p] = -1000000.0
        piv.append(int(np.argmax(far_sum)))
    o = np.sort(np.asarray(mo, np.float32))
    return np.asarray(piv[:rows], np.int32), P, float(o[max(0, min(int(rows * far), rows - 1))])


def reference_matcher(cfg):
    """→ (callable(gallery, labels, queries, nthreads) → (idx, dist, seconds), kind)."""
    from oracle import oracle_py
    m = cfg["metric"]
    if oracle_py.Ref.available(m):
        ref = oracle_py.Ref(m)
        return (lambda g, gl, q, nt: ref.bf(g, q, None, nthreads=nt, timing=True)), "reference"
    port = oracle_py.Port()
    return (lambda g, gl, q, nt: port.bf(m, g, q, nthreads=nt, timing=True)), "port"


def run_reference(args):
    """--impl reference: the reference's own CPU matching loop (BruteForce::recognize, ann.cpp:113-126; for c4
    DirectedEnumeration::recognize, ann.cpp:416-507) on all host threads, same config and metric; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cfg = dict(CONFIGS[args.config])
    nthreads = os.cpu_count() or 1
    steps = max(1, args.steps)
    per_step_s = max(2.0, min(20.0, 150.0 / (steps + max(0, min(args.warmup, 1)))))
    g, gl, q, ql = host_gallery(cfg, nthreads)
    unit, metric = ("queries/s", "queries_per_s") if cfg["kind"] == "dem" else ("evals/s", "distance_evals_per_s")
    if cfg["kind"] == "dem":
        from oracle import oracle_py
        ref = oracle_py.Ref("l2")
        # the verbatim constructor is cubic in N and cannot run at 1M rows: the 32 search pivots come from the same
        # farthest-point chain (ann.cpp:302-331) walked for 32 rows with the oracle's distances, the threshold is the FAR
        # quantile over those rows, and the state is injected into the verbatim class whose recognize() is what is timed
        piv, P, thr = host_dem_state(g, gl, 12345, 32, nthreads)
        dem = ref.dem_create_injected(g, gl, piv, P, thr)
        M = int(cfg["ratio"] * cfg["n"])
        run = lambda qq: dem.search(qq, M, nthreads=nthreads, timing=True)[-1]
        work = lambda nq: float(nq)
        kind = "reference"
        what = "DirectedEnumeration::recognize (ann.cpp:416-507)"
    else:
        match, kind = reference_matcher(cfg)
        run = lambda qq: match(g, gl, qq, nthreads)[2]
        work = lambda nq: float(nq) * cfg["n"]
        what = "BruteForce::recognize (ann.cpp:113-126)"
    t = run(q[:nthreads])                                                     # calibration
    rate = work(nthreads) / max(t, 1e-9)
    nq = int(max(nthreads, min(cfg["nq"], per_step_s * rate / work(1))))
    nq = max(nthreads, nq // nthreads * nthreads)
    for _ in range(max(0, min(args.warmup, 1))):
        run(q[:nq])
    tot_w = tot_t = 0.0
    for s in range(steps):
        lo = (s * nq) % max(1, cfg["nq"] - nq + 1)
        t = run(q[lo:lo + nq])
        tot_w += work(nq)
        tot_t += t
    value = tot_w / tot_t
    sample = "%d of %d queries per step x full %d-row gallery, %s, %d threads" % (nq, cfg["nq"], cfg["n"], what, nthreads)
    line = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * tot_t / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["name"] % {"k": 1}, "gallery": cfg["n"], "queries": cfg["nq"], "dim": cfg["d"], "k": 1,
                       "note": "the reference loop is a 1-NN argmin (no top-k); each step is a bounded query sample against the full gallery"},
            "cpu_baseline": {"value": value, "unit": unit, "cores": nthreads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------
class Env:
    def __init__(self, args):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.numa = bind_to_gpu_numa_node(torch, self.local_rank) if self.world > 1 else None
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
        import fir_b200
        self.fir = fir_b200
        # torch.distributed is the launcher's plumbing (rendezvous, barrier, max over ranks); the data-path collectives are the
        # library's own NCCL communicator (csrc/sharded.cu), whose 128-byte id travels through the process group
        self.comm = fir_b200.Comm.from_torch_distributed(self.dist)
        self.stream = torch.cuda.current_stream().cuda_stream
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)          # > 126 MB L2

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps, host_clock=False):
        """K steps, L2 flushed before each; per-step CUDA events on the launching stream (device path) or the host clock (e2e:
        the call synchronises itself); barrier + synchronize on both sides; total seconds, max over ranks."""
        torch = self.torch
        self.barrier()
        tot, evs = 0.0, []
        for _ in range(steps):
            self.flush.fill_(1)
            if host_clock:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                fn()
                tot += time.perf_counter() - t0
            else:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                evs.append((a, b))
        self.barrier()
        if not host_clock:
            tot = sum(a.elapsed_time(b) for a, b in evs) / 1e3
        return self.max_over_ranks(tot)

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.comm.close()
            self.dist.destroy_process_group()


def device_rows(env, synth, role, lo, cnt, total, cfg, out=None):
    """Rows [lo, lo+cnt) of the config's gallery / query matrix, generated and normalised on this rank's GPU."""
    rows, lab = synth.synth_rows_device(role, lo, cnt, total, cfg["d"], cfg["classes"], SEED, relu=cfg["metric"] != "l2", device=env.dev, out=out)
    env.fir.normalize_rows(rows, cfg["metric"])
    return rows, lab


def full_gallery_to_host(env, synth, cfg):
    """Rank 0 rebuilds the WHOLE gallery chunk by chunk on its GPU (the generator is keyed by (seed, row, column), so these are
    the bits every shard holds) and copies it to host memory for the reference run."""
    import numpy as np
    torch = env.torch
    g = np.empty((cfg["n"], cfg["d"]), np.float32)
    gl = np.empty(cfg["n"], np.int32)
    step = 1 << 20
    for lo in range(0, cfg["n"], step):
        cnt = min(step, cfg["n"] - lo)
        rows, lab = device_rows(env, synth, synth.ROLE_GALLERY, lo, cnt, cfg["n"], cfg)
        torch.from_numpy(g[lo:lo + cnt]).copy_(rows)
        gl[lo:lo + cnt] = lab.cpu().numpy()
        del rows, lab
    return g, gl


def run_gpu(args):
    import importlib
    import numpy as np
    cfg = dict(CONFIGS[args.config])
    env = Env(args)
    torch, fir = env.torch, env.fir
    synth = importlib.import_module("fast-image-recognition_b200.synth")
    peaks = measured_peaks()
    world, rank, dev = env.world, env.rank, env.dev
    n, nq, d, k = cfg["n"], cfg["nq"], cfg["d"], args.k
    if cfg["kind"] == "dem" and world > 1:
        if rank == 0:
            print(json.dumps({"metric": "queries_per_s", "n_gpus": world, "unavailable": "c4 (directed enumeration) runs on one GPU per index; use --gpus 1"}))
        env.close()
        return
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    if cfg["kind"] == "cls":
        run_cls(env, synth, cfg, args, peaks, steps, warmup)
        env.close()
        return

    # ---- this rank's shard of the SAME gallery (strong sharding), queries replicated -----------------------------------
    lo, hi = (n * rank) // world, (n * (rank + 1)) // world
    g_dev, gl_dev = device_rows(env, synth, synth.ROLE_GALLERY, lo, hi - lo, n, cfg)
    torch.cuda.synchronize()
    gal = fir.Gallery(g_dev, gl_dev, cfg["metric"], index_offset=lo, stream=env.stream)
    gal.set_num_classes(cfg["classes"])
    del g_dev
    torch.cuda.empty_cache()
    q_dev, ql_dev = device_rows(env, synth, synth.ROLE_QUERY, 0, nq, nq, cfg)
    q_host = torch.empty((nq, d), dtype=torch.float32).pin_memory()
    q_host.copy_(q_dev)
    torch.cuda.synchronize()
    shard = fir.RankShard(gal, env.comm, n) if world > 1 else None
    line_extra, roof, launches, dem = {}, None, None, None

    if cfg["kind"] == "bf":
        idx_dev = torch.empty((nq, k), dtype=torch.int32, device=dev)
        dist_dev = torch.empty((nq, k), dtype=torch.float32, device=dev)
        idx_host = torch.empty((nq, k), dtype=torch.int32).pin_memory()
        dist_host = torch.empty((nq, k), dtype=torch.float32).pin_memory()
        who = shard if world > 1 else gal
        step_device = lambda: who.search(q_dev, k=k, out=(idx_dev, dist_dev))
        step_host = lambda: who.search(q_host.numpy(), k=k, out=(idx_host.numpy(), dist_host.numpy()))
        work_per_step, unit, metric = float(nq) * n, "evals/s", "distance_evals_per_s"
        h2d, d2h = nq * d * 4, nq * k * 8
    elif cfg["kind"] == "pnn":
        C = cfg["classes"]
        sc_host = torch.empty((nq, C), dtype=torch.float64).pin_memory()
        lab_host = torch.empty((nq,), dtype=torch.int32).pin_memory()
        var = cfg["var"]
        if world > 1:
            step_device = lambda: shard.pnn_scores(q_dev, var)
            step_host = lambda: fir._check(fir.lib().fir_shard_pnn_scores(gal._h, env.comm._h, fir._ptr(q_host.numpy()), nq, var, n, fir.HOST,
                                                                          fir._ptr(sc_host.numpy()), fir._ptr(lab_host.numpy())))
        else:
            step_device = lambda: gal.pnn_scores(q_dev, var)
            step_host = lambda: fir._check(fir.lib().fir_pnn_scores(gal._h, fir._ptr(q_host.numpy()), nq, var, n, fir.HOST,
                                                                    fir._ptr(sc_host.numpy()), fir._ptr(lab_host.numpy())))
        work_per_step, unit, metric = float(nq) * n, "evals/s", "distance_evals_per_s"
        h2d, d2h = nq * d * 4, nq * C * 8 + nq * 4
    else:  # dem
        t0 = time.perf_counter()
        dem = fir.Dem(gal, pivot0=12345, max_chain=args.dem_chain)               # 0 = the reference's max(5, 0.015 N) rows (ann.cpp:373-375)
        torch.cuda.synchronize()
        build_s = time.perf_counter() - t0
        M = int(cfg["ratio"] * n)
        step_device = lambda: dem.search(q_dev, M)
        out_h = [torch.empty((nq,), dtype=t).pin_memory() for t in (torch.int32, torch.float32, torch.uint8, torch.int32)]
        step_host = lambda: fir._check(fir.lib().fir_dem_search(dem._h, fir._ptr(q_host.numpy()), nq, M, fir.HOST, *[fir._ptr(o.numpy()) for o in out_h]))
        work_per_step, unit, metric = float(nq), "queries/s", "queries_per_s"
        h2d, d2h = nq * d * 4, nq * 13
        line_extra["dem_build"] = {"seconds": build_s, "chain_rows": dem.chain_rows, "pivots": dem.n_pivots, "threshold": float(dem.threshold)}

    # ---- warm-up, then the timed regions ----------------------------------------------------------------------------
    for _ in range(warmup):
        step_device()
    step_host()
    env.barrier()
    sampler = ClockSampler(env.local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    gal.profile(True)
    t_dev = env.timed(step_device, steps)
    prof = {kk: gal.profile_read(kk) for kk in (0, 1, 2, 5)}
    dem_stats = dem.search_stats() if dem is not None else None
    gal.profile(False)
    st = gal.stats() if cfg["kind"] == "bf" else None
    t_e2e = env.timed(step_host, steps, host_clock=True)
    clocks = sampler.stop() if sampler else None

    value = work_per_step * steps / t_dev
    e2e = work_per_step * steps / t_e2e
    res_host = None
    if cfg["kind"] == "bf":
        launches = (st["gpu_launches"] + (2 if world > 1 else 0)) * steps                 # + pack / merge kernels of the sharded call
        k_ms, k_n = prof[0]
        flops = 2.0 * d * nq * (hi - lo)
        if k_n:
            ach = flops / (k_ms / k_n * 1e-3) / 1e12
            long_steps = t_dev / steps > 0.1                                              # seconds-long steps run at sustained clocks
            peak = peaks["tensor_tflops_sustained"] if long_steps else peaks["tensor_tflops_burst"]
            roof = {"bound": "tensor", "kernel": "l2_candidates_kernel_2cta, first pass (tcgen05.mma.cta_group::2 kind::f16, TMA, TMEM)",
                    "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "peak_source": "%s %s fp16/bf16 dense" % (peaks["source"], "sustained (steps of %.0f ms)" % (1e3 * t_dev / steps) if long_steps else "burst (steps of a few ms)"),
                    "kernel_ms": k_ms / k_n, "kernel_share_of_step": (k_ms / 1e3) / t_dev, "flops_per_launch": flops,
                    "algorithmic": "2*D flop per distance evaluation x queries x this rank's gallery rows",
                    "traffic": traffic_from_profiles(args.config if world == 1 else "", "l2_candidates_kernel_2cta")}
            try:        # DRAM floor of the launch: the fp16 shadow once per sweep of the work partition (full rounds + remainder phases / blocks)
                import ctypes as C
                n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
                dph = -(-d // 64) * 64
                ph, sl = C.c_int32(0), C.c_int32(0)
                fir.lib().fir_debug_partition_check(nq, min(hi - lo, 40_000), n_sm, 2, 1 << 40 if (hi - lo) * dph * 2 > (48 << 20) else 0, C.byref(ph), C.byref(sl))
                units = max(1, n_sm // 2)
                nqb = -(-nq // 256)
                full, rem = divmod(nqb, units)
                sweeps = (full + (ph.value if ph.value > 0 else rem)) if full > 0 else 1        # no full round: the shadow is (re)used out of L2
                roof["traffic_floor"] = float(sweeps) * (hi - lo) * dph * 2
                roof["traffic_floor_is"] = "%d sweeps (%d full rounds + %d remainder %s) x %.2f GB fp16 shadow" % (
                    sweeps, full, sweeps - full, "phases" if ph.value > 0 else "blocks", (hi - lo) * dph * 2 / 1e9)
            except Exception:
                pass
        # k = 1 on the same data (the reference's own operation)
        idx1 = torch.empty((nq, 1), dtype=torch.int32, device=dev)
        d1 = torch.empty((nq, 1), dtype=torch.float32, device=dev)
        who.search(q_dev, k=1, out=(idx1, d1))
        t_k1 = env.timed(lambda: who.search(q_dev, k=1, out=(idx1, d1)), 1)
        line_extra["k1"] = {"value": float(nq) * n / t_k1, "unit": unit, "ms": 1e3 * t_k1}
        line_extra["certificate_fallback_queries"] = st["n_fallback"]
        line_extra["reranked_candidates_per_query"] = st["n_candidates"]
        res_host = (idx_host.numpy(), dist_host.numpy())
    elif cfg["kind"] == "pnn":
        k_ms, k_n = prof[1]
        launches = 3 * steps
        if k_n:
            t_k = k_ms / k_n * 1e-3
            rows_here = hi - lo
            passes = -(-nq // 64)
            hbm = rows_here * d * 4.0 / t_k / 1e9
            pipes = peaks.get("pipes", {})
            issue_peak = pipes.get("fdiv_ieee") if cfg["metric"] == "chi2" else None
            elem = float(nq) * rows_here * d / t_k
            roof = {"bound": "hbm", "kernel": "exact_tile_kernel<%s> with the fused Parzen class-sum epilogue (CUDA cores, bit-exact feature_distance)" % cfg["metric"],
                    "achieved": hbm, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm / peaks["hbm_gbs"],
                    "algorithmic": "D*4 bytes per gallery row, the gallery read once per launch (each 64-row tile is reused by every 64-query block "
                                   "through L2: %d query blocks)" % passes,
                    "kernel_ms": 1e3 * t_k, "kernel_share_of_step": (k_ms / 1e3) / t_dev,
                    "binding": "fp32 issue (batched mode, SURVEY 8(d)): HBM binds only below ~8 queries",
                    "issue": {"achieved_elem_per_s": elem, "peak_elem_per_s": issue_peak, "frac": (elem / issue_peak) if issue_peak else None,
                              "peak_source": "profiles/r2_peak_pipes.json: IEEE fp32 divisions/s (one per chi-square element)" if issue_peak else
                                             "two glibc-exact logf per element (fp64 polynomial) — no single-pipe peak applies"},
                    "traffic": traffic_from_profiles(args.config if world == 1 else "", "exact_tile_kernel")}
        res_host = (sc_host.numpy(), lab_host.numpy())
    else:
        launches = dem_stats["gpu_launches"] * steps
        k_ms, k_n = dem_stats["candidates_kernel_ms"], dem_stats["candidates_kernel_launches"]
        if dem_stats["tensor_round"] and k_n:
            t_k = k_ms / k_n * 1e-3
            flops = 2.0 * 32 * nq * n
            ach = flops / t_k / 1e12
            peak = peaks["tensor_tflops_burst"]
            roof = {"bound": "tensor", "kernel": "l2_candidates_kernel_2cta over the pivot-space gallery P^T [N][32] (first round of the candidate walk: "
                                                 "likelihood = squared Euclidean distance between pivot-distance vectors)",
                    "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "kernel_ms": 1e3 * t_k, "kernel_share_of_step": (k_ms / 1e3) / t_dev,
                    "algorithmic": "2*32 flop per (query, gallery row) likelihood", "flops_per_launch": flops,
                    "binding": "the TMEM->register epilogue (one accumulator element per (query, row), ~3 instructions each): a K = 32 contraction cannot "
                               "load the tensor pipe; the CUDA-core form of the same pass costs 96 rounded FP32 instructions per element",
                    "epilogue_elements_per_s": float(nq) * n / t_k,
                    "peak_source": "%s burst fp16/bf16 dense" % peaks["source"], "traffic": traffic_from_profiles(args.config, "l2_candidates_kernel_2cta")}
        elif prof[2][1]:
            roof = {"bound": "hbm", "kernel": "dem_likelihood_kernel (CUDA cores)", "achieved": None, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": None,
                    "kernel_ms": prof[2][0] / prof[2][1], "traffic": None}
        res_host = tuple(o.numpy() for o in out_h)

    if rank == 0:
        line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": 1e3 * t_dev / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": {"bf": "f16 tensor-core candidates, f32 reference-exact results", "pnn": "f32 (bit-exact feature_distance) + f64 Parzen sums",
                          "dem": "f32"}[cfg["kind"]], "data": "synthetic",
                "queries_per_s": nq * steps / t_dev,
                "config": {"workload": cfg["name"] % {"k": k}, "gallery": n, "gallery_rows_per_gpu": hi - lo, "queries": nq, "dim": d, "k": k if cfg["kind"] == "bf" else None,
                           "classes": cfg["classes"],
                           "parallelism": ("gallery row shards x%d (strong: the same %d-row gallery), queries replicated, NCCL all-gather of packed (dist, idx) keys + k-way merge "
                                           "inside libfir_b200.so (fir_shard_*)" % (world, n)) if world > 1 else "single GPU",
                           "l2": "256 MiB fill before every timed step", "timing": "per-step CUDA events on the launching stream, max over ranks",
                           "generator": "Philox4x32-10 keyed by (seed, row, column), csrc/synth_common.h",
                           "e2e_path": ("fir_shard_* with the same pinned host batch on every rank: each rank uploads 1/N of the rows, NVLink all-gather assembles it; results to pinned host"
                                        if world > 1 else "C-ABI call with host buffers (H2D of the queries, D2H of the results inside the call)"),
                           "numa_node": env.numa},
                "e2e": {"value": e2e, "unit": unit, "ms_per_step": 1e3 * t_e2e / steps, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": launches, "roofline": roof, "clocks": clocks}
        line.update(line_extra)
        if not args.skip_parity:
            par, cpu = parity_and_cpu(env, synth, cfg, args, q_host.numpy(), res_host, gal, k, dem)
            line["parity"] = par
            if cpu and world == 1 and not args.skip_cpu:
                line["cpu_baseline"] = cpu
        print(json.dumps(line))
    env.barrier()
    if cfg["kind"] == "dem":
        dem.close()
    gal.close()
    env.close()


def run_cls(env, synth, cfg, args, peaks, steps, warmup):
    """fp64 PNN / kNN (classification.cpp:116-226) — single GPU; with --gpus N every rank runs a replica (the path does not shard)."""
    import numpy as np
    from oracle import oracle_py
    torch, fir = env.torch, env.fir
    n, nq, d, C, K = cfg["n"], cfg["nq"], cfg["d"], cfg["classes"], cfg["K"]
    g, gl = device_rows(env, synth, synth.ROLE_GALLERY, 0, n, n, cfg)
    q, _ = device_rows(env, synth, synth.ROLE_QUERY, 0, nq, nq, cfg)
    tr = g.cpu().numpy().astype(np.float64)
    tl = gl.cpu().numpy()
    avg = np.array([sum(tr[:, f].tolist()) for f in range(d)]) / n             # sequential column sums (split_train_test, classification.cpp:969-976)
    clf = fir.Classifier(tr, tl, C, avg)
    clf.set_stream(env.stream)
    q_dev = q.double()
    q_host = torch.empty((nq, d), dtype=torch.float64).pin_memory()
    q_host.copy_(q_dev)
    torch.cuda.synchronize()
    step_device = lambda: clf.pnn(q_dev)
    step_host = lambda: clf.pnn(q_host.numpy())
    for _ in range(warmup):
        step_device()
    lab_h, sc_h = step_host()
    env.barrier()
    sampler = ClockSampler(env.local_rank) if env.rank == 0 else None
    if sampler:
        sampler.start()
    clf.profile(True)
    t_dev = env.timed(step_device, steps)
    k_ms, k_n = clf.profile(False)
    t_e2e = env.timed(step_host, steps, host_clock=True)
    knn_lab = clf.knn(q_dev, K)
    t_knn = env.timed(lambda: clf.knn(q_dev, K), max(1, steps // 2))
    clocks = sampler.stop() if sampler else None
    if env.rank != 0:
        return
    work = float(nq) * n
    pipes = peaks.get("pipes", {})
    t_k = k_ms / max(k_n, 1) * 1e-3
    ops = 3.0 * work * d                                                        # DSUB, DMUL, DADD per element, separately rounded
    peak = (pipes.get("dadd", 0) * 2 + pipes.get("dfma", 0)) / 3.0 if pipes else None
    sample = np.unique(np.linspace(0, nq - 1, cfg["parity_q"]).astype(np.int64))
    port = oracle_py.Port()
    t0 = time.perf_counter()
    psc, plab = port.pnn(tr, tl, C, avg, q_host.numpy()[sample])
    t_cpu = time.perf_counter() - t0
    pk = port.knn(tr, tl, C, avg, q_host.numpy()[sample], K)
    rel = np.abs(sc_h[sample] - psc) / np.maximum(np.abs(psc), 1e-300)
    big = psc > psc.max(axis=1, keepdims=True) * 1e-30
    line = {"metric": "distance_evals_per_s", "value": work * steps / t_dev, "unit": "evals/s", "n_gpus": env.world, "steps": steps, "warmup": warmup,
            "ms_per_step": 1e3 * t_dev / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "queries_per_s": nq * steps / t_dev,
            "config": {"workload": cfg["name"], "gallery": n, "queries": nq, "dim": d, "classes": C, "parallelism": "single GPU (replicas only: the classifier does not shard)",
                       "l2": "256 MiB fill before every timed step", "timing": "per-step CUDA events on the launching stream"},
            "e2e": {"value": work * steps / t_e2e, "unit": "evals/s", "ms_per_step": 1e3 * t_e2e / steps, "h2d_bytes_per_step": nq * d * 8, "d2h_bytes_per_step": nq * C * 8 + nq * 4},
            "gpu_launches": 5 * steps,
            "knn": {"K": K, "value": work * max(1, steps // 2) / t_knn, "unit": "evals/s", "ms": 1e3 * t_knn / max(1, steps // 2)},
            "roofline": {"bound": "fp64", "kernel": "cls_dist_kernel<FUSE_PNN> (fp64 tiles, fused Parzen class sums)", "achieved": ops / t_k / 1e12 if k_n else None,
                         "peak": peak / 1e12 if peak else None, "unit": "Tlane-op/s", "frac": (ops / t_k / peak) if (k_n and peak) else None,
                         "algorithmic": "3 separately rounded fp64 operations per (query, row, dimension)", "kernel_ms": 1e3 * t_k,
                         "kernel_share_of_step": (k_ms / 1e3) / t_dev, "peak_source": "profiles/r2_peak_pipes.json (2 x DADD-rate + DFMA-rate) / 3", "traffic": None},
            "clocks": clocks,
            "parity": {"sample_queries": int(len(sample)), "against": "oracle port fir_oracle_pnn / fir_oracle_knn (pinned to classification.cpp in oracle/_ref)",
                       "pnn_labels_equal": bool(np.array_equal(plab, lab_h[sample])), "pnn_scores_max_rel_err": float(rel[big].max()),
                       "pnn_scores_within_1e-5": bool(rel[big].max() <= 1e-5), "knn_labels_equal": bool(np.array_equal(pk, knn_lab.cpu().numpy()[sample]))},
            "cpu_baseline": {"value": len(sample) * float(n) / t_cpu, "unit": "evals/s", "cores": 1, "kind": "port",
                             "sample": "%d of %d queries, PNNClassifier::predict_bf restated in C (oracle/fir_oracle.c), 1 thread" % (len(sample), nq)}}
    print(json.dumps(line))
    clf.close()


def parity_and_cpu(env, synth, cfg, args, q_host, res_host, gal, k, dem):
    """After the timed regions: the reference's own code over the FULL gallery on a query sample, compared with what the GPU
    returned through the host-buffer call; the same run, timed, is the cpu_baseline."""
    import numpy as np
    from oracle import oracle_py
    nthreads = os.cpu_count() or 1
    t0 = time.perf_counter()
    g, gl = full_gallery_to_host(env, synth, cfg)
    nq, n = cfg["nq"], cfg["n"]
    ns = min(cfg["parity_q"], nq)
    sample = np.unique(np.linspace(0, nq - 1, ns).astype(np.int64))
    qs = np.ascontiguousarray(q_host[sample])
    par = {"sample_queries": int(len(sample)), "gallery_rows": n}
    cpu = None
    m = cfg["metric"]
    have_ref = oracle_py.Ref.available(m)
    ref = oracle_py.Ref(m) if have_ref else None
    port = oracle_py.Port()
    if cfg["kind"] == "bf":
        idx, dist = res_host
        if have_ref:
            ri, rd, t = ref.bf(g, qs, None, nthreads=nthreads, timing=True)
        else:
            ri, rd, t = port.bf(m, g, qs, nthreads=nthreads, timing=True)
        par["against"] = ("oracle/_ref BruteForce::recognize (unmodified reference)" if have_ref else "oracle port (reference build absent)") + ", top-1 of every sample query"
        par["idx_equal"] = bool(np.array_equal(ri, idx[sample, 0]))
        par["dist_bits_equal"] = bool(np.array_equal(rd.view(np.uint32), dist[sample, 0].view(np.uint32)))
        nt = min(cfg["parity_topk_q"], len(sample))
        if k > 1 and nt > 0:
            ti, td = port.topk(m, g, qs[:nt], k, nthreads=nthreads)
            par["topk"] = {"queries": int(nt), "against": "oracle port fir_oracle_topk (pinned to oracle/_ref)", "idx_equal": bool(np.array_equal(ti, idx[sample[:nt]])),
                           "dist_bits_equal": bool(np.array_equal(td.view(np.uint32), dist[sample[:nt]].view(np.uint32)))}
        par["label_match_vs_reference"] = float((gl[np.maximum(ri, 0)] == gl[np.maximum(idx[sample, 0], 0)]).mean())
        cpu = {"value": len(sample) * float(n) / t, "unit": "evals/s", "cores": nthreads, "kind": "reference" if have_ref else "port",
               "sample": "%d of %d queries x full %d-row gallery, BruteForce::recognize (ann.cpp:113-126), %d threads" % (len(sample), nq, n, nthreads)}
    elif cfg["kind"] == "pnn":
        sc, lab = res_host
        C = cfg["classes"]
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=nthreads) as ex:                  # the port is single-threaded; ctypes releases the GIL: one query per task
            parts = list(ex.map(lambda i: port.pnn_div(m, g, gl, C, qs[i:i + 1], cfg["var"]), range(len(qs))))
        psc, plab = np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])
        rel = np.abs(sc[sample] - psc) / np.maximum(np.abs(psc), 1e-300)
        big = psc > psc.max(axis=1, keepdims=True) * 1e-30
        par["against"] = "oracle port fir_oracle_pnn_div (reference-exact distances, fp64 Parzen sums) on the full gallery"
        par["scores_max_rel_err"] = float(rel[big].max()) if big.any() else 0.0
        par["scores_within_1e-5"] = bool(par["scores_max_rel_err"] <= 1e-5)
        par["labels_equal"] = bool(np.array_equal(plab, lab[sample]))
        i1, d1 = gal.search(env.torch.from_numpy(qs).to(env.dev), k=1, path=env.fir.PATH_EXACT) if env.world == 1 else (None, None)
        if have_ref:
            ri, rd, t = ref.bf(g, qs, None, nthreads=nthreads, timing=True)
            if i1 is not None:
                par["top1"] = {"against": "oracle/_ref BruteForce::recognize", "idx_equal": bool(np.array_equal(ri, i1.cpu().numpy()[:, 0])),
                               "dist_bits_equal": bool(np.array_equal(rd.view(np.uint32), d1.cpu().numpy()[:, 0].view(np.uint32)))}
            cpu = {"value": len(sample) * float(n) / t, "unit": "evals/s", "cores": nthreads, "kind": "reference",
                   "sample": "%d of %d queries x full %d-row gallery, %s feature_distance argmin (ann.cpp:113-126), %d threads" % (len(sample), nq, n, m, nthreads)}
    else:
        idx, dist, below, evals = res_host
        M = int(cfg["ratio"] * n)
        P, piv, thr = dem.P, dem.pivots, float(dem.threshold)
        if have_ref:
            rdem = ref.dem_create_injected(g, gl, piv, P, thr)
            out = rdem.search(qs, M, nthreads=nthreads, timing=True)
            ri, rd, rb, re_, t = out
            par["against"] = "oracle/_ref DirectedEnumeration::recognize (verbatim) over the GPU-built pivots / pivot-distance rows / threshold"
            par["idx_equal"] = bool(np.array_equal(ri, idx[sample]))
            par["dist_bits_equal"] = bool(np.array_equal(np.asarray(rd, np.float32).view(np.uint32), dist[sample].view(np.uint32)))
            par["below_equal"] = bool(np.array_equal(np.asarray(rb).astype(np.uint8), below[sample]))
            par["evals_equal"] = bool(np.array_equal(np.asarray(re_).astype(np.int32), evals[sample]))
            cpu = {"value": len(sample) / t, "unit": "queries/s", "cores": nthreads, "kind": "reference",
                   "sample": "%d of %d queries, DirectedEnumeration::recognize (ann.cpp:416-507) over the full %d-row gallery, %d threads" % (len(sample), nq, n, nthreads)}
            # the other check budgets (ratio of the gallery a query may examine, ann.h:26): agreement with the verbatim recognize(), checked-%, q/s
            sweep = []
            qd = env.torch.from_numpy(q_host).to(env.dev)
            for ratio in (0.025, 0.1, 0.25, 0.5):
                Mr = int(ratio * n)
                dem.search(qd, Mr)
                a, b = env.torch.cuda.Event(enable_timing=True), env.torch.cuda.Event(enable_timing=True)
                a.record(); gi, gd, gb, ge = dem.search(qd, Mr); b.record(); env.torch.cuda.synchronize()
                gi, gd, gb, ge = (x.cpu().numpy() for x in (gi, gd, gb, ge))
                ri, rd, rb, re_, _ = rdem.search(qs, Mr, nthreads=nthreads, timing=True)
                sweep.append({"ratio": ratio, "queries_per_s": nq / (a.elapsed_time(b) * 1e-3), "checked_percent": float(100.0 * ge.mean() / n),
                              "below_threshold_frac": float(gb.mean()),
                              "equal_to_reference": bool(np.array_equal(ri, gi[sample]) and np.array_equal(np.asarray(rd, np.float32).view(np.uint32), gd[sample].view(np.uint32))
                                                         and np.array_equal(np.asarray(rb).astype(np.uint8), gb[sample]) and np.array_equal(np.asarray(re_).astype(np.int32), ge[sample]))})
            par["ratio_sweep"] = sweep
            rdem.close()
        qd_all = env.torch.from_numpy(q_host).to(env.dev)
        bi, _ = gal.search(qd_all, k=1)
        try:        # the exact GPU brute force over the same gallery, timed beside the approximate search (speed-up vs GPU BF)
            ea, eb = env.torch.cuda.Event(enable_timing=True), env.torch.cuda.Event(enable_timing=True)
            ea.record(); gal.search(qd_all, k=1); eb.record(); env.torch.cuda.synchronize()
            par["gpu_brute_force_ms"] = float(ea.elapsed_time(eb))
        except Exception:
            pass
        par["recall_at_1_vs_bf"] = float((bi.cpu().numpy()[:, 0] == idx).mean())
        par["below_threshold_frac"] = float(below.mean())
        par["checked_percent"] = float(100.0 * evals.mean() / n)
    par["seconds"] = time.perf_counter() - t0
    return par, cpu


def bind_to_gpu_numa_node(torch, local_rank):
    """Multi-rank runs: run this process (and so allocate its pinned buffers) on the CPUs of the NUMA node the GPU hangs off."""
    try:
        p = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c5", choices=sorted(CONFIGS))
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--dem-chain", type=int, default=0, help="c4: pivot-chain rows (0 = the reference's 0.015 N)")
    ap.add_argument("--skip-cpu", action="store_true", help="omit the cpu_baseline leg")
    ap.add_argument("--skip-parity", action="store_true", help="omit the parity leg (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
