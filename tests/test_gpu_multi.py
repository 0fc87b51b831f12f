"""GPU: the row-sharded gallery behind the C-ABI (csrc/sharded.cu) returns exactly what ONE gallery returns.
  * single process driving the GPUs (fir_sharded_*): runs with however many devices the box has — with one device the
    pack / merge / unpack kernels and the phase driver are still exercised (world = 1);
  * one process per GPU over NCCL (fir_comm_* + fir_shard_*): needs >= 2 devices, skipped otherwise (the driver's test box
    has one GPU; profiles/r2_multigpu_tests.log records the 2-GPU run of this file)."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _data(port, metric="l2", n=6000, nq=300, d=256, classes=24, seed=5):
    from util import make_data
    g, gl, q, ql = make_data(port, metric, n, nq, d, classes, seed=seed)
    g[5000:5500] = g[100:600]                          # duplicates across shards ⇒ ties resolved by global index
    return g, gl, q, ql


def _check_against_oracle(port, fir, who, g, gl, q, classes, device_queries=None):
    ok = True
    for k, path in ((1, fir.PATH_TENSOR), (10, fir.PATH_TENSOR), (5, fir.PATH_EXACT)):
        idx, dd = who.search(q if device_queries is None else device_queries, k=k, path=path)
        if device_queries is not None:
            import torch
            torch.cuda.synchronize()
            idx, dd = idx.cpu().numpy(), dd.cpu().numpy()
        oi, od = port.topk("l2", g, q, k, nthreads=4)
        ok = ok and np.array_equal(idx, oi) and np.array_equal(dd.view(np.uint32), od.view(np.uint32))
    mn, arg = who.class_min(q)
    omn, oarg = port.class_min("l2", g, gl, classes, q)
    ok = ok and np.array_equal(np.asarray(arg), oarg) and np.array_equal(np.asarray(mn).view(np.uint32), omn.view(np.uint32))
    sc, lab = who.pnn_scores(q, 2e-4)
    osc, olab = port.pnn_div("l2", g, gl, classes, q, 2e-4)
    ok = ok and np.allclose(sc, osc, rtol=1e-5, atol=0) and np.array_equal(lab, olab)       # tolerance: fp64 sums in shard order
    return ok


def test_single_process_sharded_matches_oracle(fir, port):
    import torch
    g, gl, q, ql = _data(port)
    for n_gpus in sorted({1, min(2, torch.cuda.device_count()), torch.cuda.device_count()}):
        sh = fir.Sharded(g, gl, "l2", n_gpus=n_gpus)
        assert sh.n_gpus == n_gpus and sh.n == len(g) and sh.n_classes == 24
        assert _check_against_oracle(port, fir, sh, g, gl, q, 24), "n_gpus=%d" % n_gpus
        sh.close()


def test_single_process_sharded_dem(fir, port):
    """fir_sharded_dem_*: one host thread per GPU runs the collective build / search; equals ONE index over the whole gallery."""
    import torch
    from util import make_data
    g, gl, q, ql = make_data(port, "l2", 18000, 100, 48, 120, seed=14, sigma=1.2)
    single = fir.Gallery(g, gl, "l2")
    sdem = fir.Dem(single, pivot0=9, max_chain=40)
    for n_gpus in sorted({1, torch.cuda.device_count()}):
        sh = fir.Sharded(g, gl, "l2", n_gpus=n_gpus)
        dem = sh.dem(pivot0=9, max_chain=40)
        assert np.array_equal(dem.pivots, sdem.pivots) and np.float32(dem.threshold).view(np.uint32) == np.float32(sdem.threshold).view(np.uint32)
        for M in (0, 40, 56, 300):
            for a, b in zip(dem.search(q, M), sdem.search(q, M)):
                assert np.array_equal(a, b), (n_gpus, M)
        dem.close(); sh.close()
    sdem.close(); single.close()


def test_sharded_rejects_bad_arguments(fir):
    import torch
    with pytest.raises(fir.FirError):
        fir.Sharded(np.zeros((4, 8), np.float32), None, "l2", n_gpus=torch.cuda.device_count() + 1)
    sh = fir.Sharded(np.eye(8, dtype=np.float32), np.arange(8, dtype=np.int32), "l2", n_gpus=1)
    with pytest.raises(fir.FirError):
        sh.search(np.eye(8, dtype=np.float32), k=0)
    with pytest.raises(fir.FirError):
        sh.pnn_scores(np.eye(8, dtype=np.float32), 0.0)
    idx, dd = sh.search(np.eye(8, dtype=np.float32), k=2, path=fir.PATH_EXACT)
    assert idx[:, 0].tolist() == list(range(8)) and (dd[:, 0] == 0).all()
    sh.close()


def _worker(rank, world, port_no, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)      # the launcher's group only carries the communicator id
    import fir_b200
    from oracle import oracle_py
    sharded = importlib.import_module("fast-image-recognition_b200.sharded")
    port = oracle_py.Port()
    g, gl, q, ql = _data(port)
    lo, hi = sharded.shard_bounds(len(g), world, rank)
    dev = torch.device("cuda", rank)
    comm = fir_b200.Comm.from_torch_distributed(dist)
    gal = fir_b200.Gallery(torch.from_numpy(g[lo:hi]).to(dev), torch.from_numpy(gl[lo:hi]).to(dev), "l2", index_offset=lo)
    gal.set_num_classes(24)
    rs = fir_b200.RankShard(gal, comm, len(g))
    ok = _check_against_oracle(port, fir_b200, rs, g, gl, q, 24)                                    # host buffers: sliced upload + NVLink all-gather
    qd = torch.from_numpy(q).to(dev)
    idx, dd = rs.search(qd, k=10)                                                                   # device buffers, asynchronous
    torch.cuda.synchronize()
    oi, od = port.topk("l2", g, q, 10, nthreads=4)
    ok = ok and np.array_equal(idx.cpu().numpy(), oi) and np.array_equal(dd.cpu().numpy().view(np.uint32), od.view(np.uint32))
    # ---- directed enumeration over the shards: ONE index over the whole gallery (pivots, threshold, candidate order) ----
    def dem_case(n, d, classes, seed, max_chain, budgets):
        from util import make_data
        good = True
        gg, ggl, qq, _ = make_data(port, "l2", n, 120, d, classes, seed=seed, sigma=1.2)
        a, b = sharded.shard_bounds(n, world, rank)
        shard = fir_b200.Gallery(torch.from_numpy(gg[a:b]).to(dev), torch.from_numpy(ggl[a:b]).to(dev), "l2", index_offset=a)
        shard.set_num_classes(classes)
        rsh = fir_b200.RankShard(shard, comm, n)
        dem = rsh.dem(pivot0=77, max_chain=max_chain)
        single = fir_b200.Gallery(torch.from_numpy(gg).to(dev), torch.from_numpy(ggl).to(dev), "l2")      # the same build on ONE gallery
        sdem = fir_b200.Dem(single, pivot0=77, max_chain=max_chain)
        good = good and np.array_equal(dem.pivots, sdem.pivots) and np.float32(dem.threshold).view(np.uint32) == np.float32(sdem.threshold).view(np.uint32)
        good = good and np.array_equal(dem.P.view(np.uint32), sdem.P[:, a:b].view(np.uint32))
        piv, P, thr = sdem.pivots, sdem.P, float(sdem.threshold)
        for t, low in ((thr, False), (thr * 0.05, True)):
            d1 = dem if not low else None
            if low:                                              # nothing is ever under the threshold: the walk runs its whole budget
                p_ = fir_b200.DemParams(77, 0, 0.01, t, max_chain, 0)
                h = fir_b200.C.c_void_p(None)
                fir_b200._check(fir_b200.lib().fir_shard_dem_build(shard._h, comm._h, n, fir_b200.C.byref(p_), fir_b200.C.byref(h)))
                d1 = fir_b200.Dem._adopt(shard, h)
            for M in budgets:
                got = d1.search(qq, M)
                want = port.dem_search("l2", gg, piv, P, t, M, qq)
                for name, x, y in zip(("idx", "dist", "below", "evals"), got, want):
                    if not np.array_equal(x, y):
                        good = False
                        print("DEM mismatch rank", rank, n, low, M, name, int((x != y).sum()), flush=True)
            if low:
                d1.close()
        dem.close(); sdem.close(); single.close(); shard.close()
        return good
    ok = ok and dem_case(5000, 48, 40, 7, 40, (0, 33, 60, 300, 900))             # small shards: CUDA-core rounds only
    ok = ok and dem_case(20000, 64, 150, 8, 40, (0, 40, 56, 57, 200, 1500))      # >= 8192 rows per shard: tensor first round + exchange
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)                                                     # every rank must have received the merged answer
    if rank == 0:
        with open(out, "w") as f:
            f.write("ok nccl=%d" % comm.nccl_version if int(flag.item()) else "mismatch")
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


def test_rank_shards_over_nccl_match_single_gallery_oracle(tmp_path, port):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert open(out).read().startswith("ok")


def test_shard_dem_world_one_matches_plain_build(fir, port):
    """fir_shard_dem_build with a one-rank communicator (no NCCL needed) walks the same code as the multi-rank build — the
    pivot record, the triple exchange, global row indices — and must equal the plain single-gallery DEM."""
    from util import make_data
    g, gl, q, ql = make_data(port, "l2", 9000, 90, 48, 60, seed=12, sigma=1.2)
    gal = fir.Gallery(g, gl, "l2")
    comm = fir.Comm(0, 1)
    rs = fir.RankShard(gal, comm, len(g))
    dem = rs.dem(pivot0=5, max_chain=40)
    plain = fir.Dem(gal, pivot0=5, max_chain=40)
    assert np.array_equal(dem.pivots, plain.pivots) and np.array_equal(dem.P.view(np.uint32), plain.P.view(np.uint32))
    assert np.float32(dem.threshold).view(np.uint32) == np.float32(plain.threshold).view(np.uint32)
    for M in (0, 40, 56, 500):
        for a, b in zip(dem.search(q, M), plain.search(q, M)):
            assert np.array_equal(a, b), M
    dem.close(); plain.close(); comm.close(); gal.close()
