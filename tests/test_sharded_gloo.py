"""N>1 host-side logic on CPU: 2 ranks over gloo.  The rank-local search is served by the oracle (test
infrastructure) so that only the sharding / gather / merge plumbing of sharded.py is under test here; the
same plumbing runs over NCCL with the CUDA kernels in bench.py --gpus N."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _OracleLocal:
    def __init__(self, port, rows, labels, metric, lo):
        self.port, self.rows, self.labels, self.metric, self.lo = port, rows, labels, metric, lo

    def search(self, q, k=1):
        i, d = self.port.topk(self.metric, self.rows, q, k)
        return np.where(i < 0, -1, i + self.lo).astype(np.int32), d

    def class_min(self, q):
        mn, arg = self.port.class_min(self.metric, self.rows, self.labels, 6, q)
        return mn, np.where(arg < 0, -1, arg + self.lo).astype(np.int32)

    def pnn_scores(self, q, var, n_total=0):
        sc, lab = self.port.pnn_div(self.metric, self.rows, self.labels, 6, q, var)
        return sc * (self.rows.shape[0] / float(n_total)), lab


def _worker(rank, world, port_no, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle_py
    from util import make_data
    sharded = importlib.import_module("fast-image-recognition_b200.sharded")
    port = oracle_py.Port()
    g, gl, q, ql = make_data(port, "l2", 501, 23, 32, 6, seed=3)
    g[400:] = g[:101]                                   # duplicates across the shard boundary ⇒ ties resolved by global index
    lo, hi = sharded.shard_bounds(len(g), world, rank)
    sg = sharded.ShardedGallery(g[lo:hi], gl[lo:hi], "l2", len(g), lo, dist=dist,
                                local_factory=lambda r, l, m, o: _OracleLocal(port, r, l, m, o), merge=sharded.merge_topk_host)
    idx, dd = sg.search(q, k=7)
    oi, od = port.topk("l2", g, q, 7)
    ok = np.array_equal(idx, oi) and np.array_equal(dd.view(np.uint32), od.view(np.uint32))
    mn, arg = sg.class_min(q)                           # all-reduce(min) of packed (dist, idx) keys
    omn, oarg = port.class_min("l2", g, gl, 6, q)
    ok = ok and np.array_equal(arg, oarg) and np.array_equal(mn.view(np.uint32), omn.view(np.uint32))
    sc, lab = sg.pnn_scores(q, 2e-4)
    osc, olab = port.pnn_div("l2", g, gl, 6, q, 2e-4)
    ok = ok and np.allclose(sc, osc, rtol=1e-9) and np.array_equal(lab, olab)
    full = sg.gather_queries(q)                         # 23 rows over 2 ranks: slices of 12 + 11 (padded), assembled everywhere
    ok = ok and tuple(full.shape) == q.shape and np.array_equal(full.numpy().view(np.uint32), q.view(np.uint32))
    if rank == 0:
        with open(out, "w") as f:
            f.write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_gather_merge(tmp_path, port):
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert open(out).read() == "ok"


def test_shard_bounds_cover_everything():
    sharded = importlib.import_module("fast-image-recognition_b200.sharded")
    for n in (1, 7, 100, 10_000_001):
        for w in (1, 2, 4, 8):
            b = [sharded.shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_key_format_roundtrip_and_order():
    sharded = importlib.import_module("fast-image-recognition_b200.sharded")
    rng = np.random.default_rng(0)
    d = rng.standard_normal(4096).astype(np.float32)
    d[:8] = [0.0, 1e-38, 3.4e38, -1.0, 0.5, 0.5, 100000.0, 1e-45]
    i = rng.integers(0, 2**31 - 1, 4096).astype(np.int32)
    i[100:110] = -1
    keys = sharded.pack_keys(d, i)
    dd, ii = sharded.unpack_keys(keys)
    live = i >= 0
    assert np.array_equal(ii, i) and np.array_equal(dd[live].view(np.uint32), d[live].view(np.uint32))
    order = np.argsort(keys[live], kind="stable")
    want = np.lexsort((i[live], d[live]))
    assert np.array_equal(order, want)                  # key order = lexicographic (dist, idx); empties sort last
    assert (keys[~live] == sharded.EMPTY_KEY).all() and keys[live].max() < sharded.EMPTY_KEY


def test_merge_host_ties_and_empties():
    sharded = importlib.import_module("fast-image-recognition_b200.sharded")
    pd = np.array([[[0.1, 0.2, 0.0]], [[0.1, 0.15, 0.3]]], np.float32)
    pi = np.array([[[5, 9, -1]], [[2, 7, 8]]], np.int32)
    oi, od = sharded.merge_topk_host(pd, pi, 4)
    assert oi.tolist() == [[2, 5, 7, 9]]
    assert np.allclose(od, [[0.1, 0.1, 0.15, 0.2]])
