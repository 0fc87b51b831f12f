#!/usr/bin/env python
"""bench.py — headline benchmark of the matching hot path (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--k 10]

A "step" is one pass of the hot path over one batch: every query of the batch (10 000 x 512-d, L2-normalised)
matched against the labelled gallery (100 000 x 512-d per GPU) — tensor-core candidate generation, exact fp32
rerank, certificate, top-k — returning the k nearest (feature_distance, index) pairs per query.

  value      distance evaluations per second (queries x gallery rows / s), inputs resident in HBM,
             CUDA-event timed on the launching stream, max over ranks.
  e2e        the same through the C-ABI with HOST (pinned) buffers: H2D of the queries and D2H of the
             indices/distances inside the timed region.
  roofline   dominant kernel (l2_candidates_kernel, tcgen05): 2*D flops per evaluation, timed live with
             CUDA events around each launch, against MEASURED_PEAKS.json's sustained bf16/fp16 tensor peak.
  cpu_baseline  the reference's own BruteForce::recognize (oracle/_ref, unmodified reference code; or the C
             restatement if that library is absent) on the host cores, on a bounded query sample.

N > 1 (torchrun): the gallery is row-sharded (one 100k-row shard per rank, global indices by shard offset),
queries are replicated, each rank's exact top-k is all-gathered over NCCL and merged on the GPU — weak scaling.
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

N_GALLERY, N_QUERY, DIM, N_CLASSES = 100_000, 10_000, 512, 1000
METRIC, UNIT = "distance_evals_per_s", "evals/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            j = json.load(f)
        return {"tensor_tflops": float(j.get("bf16_tflops_sustained", j.get("bf16_tflops", 1418.0))),
                "tensor_tflops_burst": float(j.get("bf16_tflops", 1687.2)), "hbm_gbs": float(j.get("hbm_gbs", 6545.9)),
                "source": "measured"}
    return {"tensor_tflops": 1400.0, "tensor_tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback"}


def traffic_from_profiles():
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed ncu --set full
    capture of this workload (profiles/r1_traffic.json); None when no capture is recorded."""
    p = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            j = json.load(f)
        return j.get("l2_candidates_kernel_2cta", {}).get("dram_bytes_per_launch")
    return None


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons while the timed region runs: NVML (sub-millisecond per sample, every 2 ms) when
    pynvml is importable, else `nvidia-smi` (the recipe's clocks line, ~5 samples/s)."""
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


def cpu_reference_arm(g, q, k_unused, seconds_target=15.0, nthreads=None):
    """The reference's CPU implementation of the path on the host cores (bounded sample of the workload)."""
    from oracle import oracle_py
    nthreads = nthreads or os.cpu_count() or 1
    per_thread = 4
    if oracle_py.Ref.available("l2"):
        ref, kind = oracle_py.Ref("l2"), "reference"
        run = lambda qq: ref.bf(g, qq, None, nthreads=nthreads, timing=True)[2]
    else:
        port, kind = oracle_py.Port(), "port"
        run = lambda qq: port.bf("l2", g, qq, nthreads=nthreads, timing=True)[2]
    t = run(q[: per_thread * nthreads])                                   # calibration pass
    rate = per_thread * nthreads * g.shape[0] / max(t, 1e-9)
    nq = int(min(q.shape[0], max(per_thread * nthreads, seconds_target * rate / g.shape[0])))
    nq = max(nthreads, nq // nthreads * nthreads)
    t = run(q[:nq])
    return {"value": nq * g.shape[0] / t, "unit": UNIT, "cores": nthreads, "kind": kind, "seconds": t,
            "sample": "%d of %d queries x full %d-row gallery, BruteForce::recognize (ann.cpp:113-126), %d threads"
                      % (nq, q.shape[0], g.shape[0], nthreads)}


def run_reference(args):
    """--impl reference: the reference's own CPU path, all host threads, same config/metric; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    import importlib
    synth = importlib.import_module("fast-image-recognition_b200.synth")
    from oracle import oracle_py
    port = oracle_py.Port()
    g, gl, q, ql = synth.make_split(N_GALLERY, N_QUERY, DIM, N_CLASSES, "l2")
    g, q = port.normalize_rows("l2", g), port.normalize_rows("l2", q)      # loader normalisation (oracle restatement)
    nthreads = os.cpu_count() or 1
    steps = max(1, args.steps)
    per_step_s = max(2.0, min(20.0, 150.0 / (steps + args.warmup)))
    res = None
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_reference_arm(g, q, args.k, per_step_s / 2, nthreads)
    tot_evals, tot_t = 0.0, 0.0
    for _ in range(steps):
        res = cpu_reference_arm(g, q, args.k, per_step_s, nthreads)
        tot_evals += res["value"] * res["seconds"]
        tot_t += res["seconds"]
    value = tot_evals / tot_t
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "queries_per_s": value / N_GALLERY,
            "config": {"workload": "C2: 100k x 512 gallery, 10k queries, L2 1-NN (reference CPU loop on a bounded query sample)",
                       "gallery": N_GALLERY, "queries": N_QUERY, "dim": DIM, "k": 1},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": res["cores"], "kind": res["kind"], "sample": res["sample"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_gpu(args):
    import importlib
    import numpy as np
    import torch
    import fir_b200
    synth = importlib.import_module("fast-image-recognition_b200.synth")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else None      # pinned staging buffers next to the GPU's PCIe root
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    k = args.k
    peaks = measured_peaks()

    # ---- synthetic workload: one 100k-row class-major shard per rank; queries replicated ------------------
    g_np, gl_np, q_np, ql_np = synth.make_split(N_GALLERY, N_QUERY, DIM, N_CLASSES, "l2", seed=0, shard=rank)   # same classes + queries on every rank
    g_dev = torch.from_numpy(g_np).to(dev)
    q_dev = torch.from_numpy(q_np).to(dev)
    fir_b200.normalize_rows(g_dev, "l2")         # loader normalisation (db_features.cpp:79-101) on the GPU
    fir_b200.normalize_rows(q_dev, "l2")
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream().cuda_stream
    gal = fir_b200.Gallery(g_dev, torch.from_numpy(gl_np).to(dev), "l2", index_offset=rank * N_GALLERY, stream=stream)
    q_host = torch.empty((N_QUERY, DIM), dtype=torch.float32).pin_memory()
    q_host.copy_(q_dev)
    idx_host = torch.empty((N_QUERY, k), dtype=torch.int32).pin_memory()
    dist_host = torch.empty((N_QUERY, k), dtype=torch.float32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2

    gathered_d = gathered_i = q_stage_dev = q_gather_dev = q_slice_dev = None
    if world > 1:
        per = -(-N_QUERY // world)                                           # equal slices, the last one padded
        q_lo, q_hi = min(N_QUERY, rank * per), min(N_QUERY, (rank + 1) * per)
        q_gather_dev = torch.zeros((world * per, DIM), dtype=torch.float32, device=dev)
        q_stage_dev = q_gather_dev[:N_QUERY]
        q_slice_dev = torch.zeros((per, DIM), dtype=torch.float32, device=dev)
        gathered_d = torch.empty((world * N_QUERY, k), dtype=torch.float32, device=dev)
        gathered_i = torch.empty((world * N_QUERY, k), dtype=torch.int32, device=dev)

    def step_device():
        idx, dd = gal.search(q_dev, k=k, path=fir_b200.PATH_AUTO)
        if world > 1:
            dist.all_gather_into_tensor(gathered_d, dd)
            dist.all_gather_into_tensor(gathered_i, idx)
            idx, dd = fir_b200.merge_topk(gathered_d.view(world, N_QUERY, k), gathered_i.view(world, N_QUERY, k), stream=stream)
        return idx, dd

    def step_host():
        # the call a user of the C-ABI makes: host buffers in, host buffers out (copies + sync inside)
        if world > 1:
            # sharded gallery: the rank's own top-k is an intermediate, so it stays on the device — pinned queries in,
            # search (device pointers), all-gather + merge, merged result out to pinned host memory
            q_slice_dev[: q_hi - q_lo].copy_(q_host[q_lo:q_hi], non_blocking=True)   # 1/world of the batch over this rank's own PCIe link
            dist.all_gather_into_tensor(q_gather_dev, q_slice_dev)                   # the rest over NVLink
            idx, dd = gal.search(q_stage_dev, k=k, path=fir_b200.PATH_AUTO)
            dist.all_gather_into_tensor(gathered_d, dd)
            dist.all_gather_into_tensor(gathered_i, idx)
            mi, md = fir_b200.merge_topk(gathered_d.view(world, N_QUERY, k), gathered_i.view(world, N_QUERY, k), stream=stream)
            idx_host.copy_(mi, non_blocking=True)
            dist_host.copy_(md, non_blocking=True)
            torch.cuda.synchronize()
            return idx_host.numpy(), dist_host.numpy()
        q_stage = q_host.numpy()
        return gal.search(q_stage, k=k, path=fir_b200.PATH_AUTO)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, host_clock=False):
        """K steps, L2 flushed before each, per-step CUDA events (device path) or host clock (e2e: the call
        synchronises itself); returns total seconds, max over ranks."""
        barrier()
        tot = 0.0
        evs = []
        for _ in range(steps):
            flush.fill_(1)
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
        barrier()
        if not host_clock:
            tot = sum(a.elapsed_time(b) for a, b in evs) / 1e3
        if world > 1:
            t = torch.tensor([tot], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tot = float(t.item())
        return tot

    for _ in range(max(3, args.warmup)):
        step_device()
    step_host()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    gal.profile(True)
    t_dev = timed(step_device, args.steps)
    k_ms, k_n = gal.profile_read(0)
    gal.profile(False)
    st = gal.stats()
    launches_per_step = st["gpu_launches"] + (1 if world > 1 else 0)
    t_e2e = timed(step_host, args.steps, host_clock=True)
    clocks = sampler.stop() if sampler else None

    # k = 1 on the same data (BASELINE config quotes k=1 / k=10)
    for _ in range(2):
        gal.search(q_dev, k=1)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush.fill_(1)
    a.record()
    idx1, _ = gal.search(q_dev, k=1)
    b.record()
    torch.cuda.synchronize()
    t_k1 = a.elapsed_time(b) / 1e3
    acc = float((torch.from_numpy(gl_np).to(dev)[(idx1[:, 0].long() - rank * N_GALLERY).clamp(0, N_GALLERY - 1)] ==
                 torch.from_numpy(ql_np).to(dev)).float().mean().item()) if world == 1 else None

    if rank == 0:
        evals_per_step = float(N_QUERY) * N_GALLERY * world
        value = evals_per_step * args.steps / t_dev
        e2e = evals_per_step * args.steps / t_e2e
        flops_per_launch = 2.0 * DIM * N_QUERY * N_GALLERY
        achieved = flops_per_launch / (k_ms / max(k_n, 1) * 1e-3) / 1e12 if k_n else None
        cpu = cpu_reference_arm(g_np_norm(g_dev), q_host.numpy(), k) if (world == 1 and not args.skip_cpu) else None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
                "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f16 tensor-core candidates + f32 exact rerank", "data": "synthetic",
                "queries_per_s": value / (N_GALLERY * world),
                "config": {"workload": "C2: synthetic 512-d L2-normalised embeddings, 100k gallery x 10k queries, Euclidean top-%d "
                                       "(BASELINE.json configs[1]); per GPU: one 100k-row shard, queries replicated" % k,
                           "gallery_per_gpu": N_GALLERY, "queries": N_QUERY, "dim": DIM, "k": k, "classes": N_CLASSES,
                           "parallelism": "gallery row-shards x%d + NCCL all-gather top-k merge" % world if world > 1 else "single GPU",
                           "l2": "flushed before every timed step (256 MiB fill)", "timing": "per-step CUDA events on the launching stream",
                           "e2e_path": ("pinned host queries: each rank uploads 1/N of the rows, NVLink all-gather assembles the batch; search; all-gather + merge of the top-k on the device; merged top-k -> pinned host"
                                        if world > 1 else "fir_search_topk with host buffers (H2D of the queries, D2H of the top-k inside the call)"),
                           "numa_node": numa},
                "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": 1e3 * t_e2e / args.steps,
                        "h2d_bytes_per_step": N_QUERY * DIM * 4, "d2h_bytes_per_step": N_QUERY * k * 8},
                "gpu_launches": launches_per_step * args.steps,
                "k1": {"value": float(N_QUERY) * N_GALLERY / t_k1, "unit": UNIT, "ms": 1e3 * t_k1, "label_accuracy": acc},
                "certificate_fallback_queries": st["n_fallback"], "reranked_candidates_per_query": st["n_candidates"],
                "roofline": {"bound": "tensor", "kernel": "l2_candidates_kernel_2cta, first pass (tcgen05.mma.cta_group::2 kind::f16, TMA, TMEM)",
                             "achieved": achieved, "peak": peaks["tensor_tflops"], "unit": "TFLOP/s",
                             "frac": (achieved / peaks["tensor_tflops"]) if achieved else None,
                             "peak_source": "%s sustained fp16/bf16 dense (MEASURED_PEAKS.json)" % peaks["source"],
                             "kernel_ms": k_ms / max(k_n, 1), "kernel_share_of_step": (k_ms / 1e3) / t_dev if t_dev else None,
                             "flops_per_launch": flops_per_launch, "traffic": traffic_from_profiles()},
                "clocks": clocks}
        if cpu:
            line["cpu_baseline"] = {kk: cpu[kk] for kk in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    gal.close()


def g_np_norm(g_dev):
    return g_dev.cpu().numpy()


def bind_to_gpu_numa_node(torch, local_rank):
    """Multi-rank runs: run this process (and so allocate its pinned buffers) on the CPUs of the NUMA node the GPU hangs off.
    Best effort — returns the node or None."""
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
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--skip-cpu", action="store_true", help="omit the cpu_baseline leg (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
