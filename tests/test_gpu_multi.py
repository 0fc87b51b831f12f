"""GPU (needs >= 2 devices, skipped otherwise): the row-sharded gallery over NCCL returns exactly what one GPU returns."""
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


def _worker(rank, world, port_no, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import fir_b200
    from oracle import oracle_py
    from util import make_data
    sharded = importlib.import_module("fast-image-recognition_b200.sharded")
    port = oracle_py.Port()
    g, gl, q, ql = make_data(port, "l2", 6000, 300, 256, 24, seed=5)
    g[5000:5500] = g[100:600]                          # duplicates across shards ⇒ ties resolved by global index
    lo, hi = sharded.shard_bounds(len(g), world, rank)
    dev = torch.device("cuda", rank)
    sg = sharded.ShardedGallery(torch.from_numpy(g[lo:hi]).to(dev), torch.from_numpy(gl[lo:hi]).to(dev), "l2", len(g), lo, dist=dist)
    ok = True
    for k, path in ((1, fir_b200.PATH_TENSOR), (10, fir_b200.PATH_TENSOR), (5, fir_b200.PATH_EXACT)):
        idx, dd = sg.search(torch.from_numpy(q).to(dev), k=k, path=path)
        torch.cuda.synchronize()
        oi, od = port.topk("l2", g, q, k, nthreads=4)
        ok = ok and np.array_equal(idx.cpu().numpy(), oi) and np.array_equal(dd.cpu().numpy().view(np.uint32), od.view(np.uint32))
    sc, lab = sg.pnn_scores(torch.from_numpy(q).to(dev), 2e-4)
    osc, olab = port.pnn_div("l2", g, gl, 24, q, 2e-4)
    ok = ok and np.allclose(sc.cpu().numpy(), osc, rtol=1e-5, atol=0) and np.array_equal(lab.cpu().numpy(), olab)
    if rank == 0:
        with open(out, "w") as f:
            f.write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_shards_match_single_gpu_oracle(tmp_path, port):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert open(out).read() == "ok"
