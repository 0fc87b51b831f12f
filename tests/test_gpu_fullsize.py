"""GPU parity at BASELINE.json's full C2 size (100k x 512 gallery, 10k queries): the oracle on a query sample, and
size-independent properties on everything (tensor path == exact path, k=1 is the head of k=10, per-class minima fold
to the global minimum, shards merge to the unsharded answer, search is idempotent)."""
import importlib
import os

import numpy as np
import pytest
import torch

from util import bits

pytestmark = pytest.mark.gpu
synth = importlib.import_module("fast-image-recognition_b200.synth")
N, Q, D, C = 100_000, 10_000, 512, 1000


@pytest.fixture(scope="module")
def c2(fir):
    g, gl, q, ql = synth.make_split(N, Q, D, C, "l2")
    dev = torch.device("cuda", 0)
    gd, qd = torch.from_numpy(g).to(dev), torch.from_numpy(q).to(dev)
    fir.normalize_rows(gd, "l2")
    fir.normalize_rows(qd, "l2")
    gal = fir.Gallery(gd, torch.from_numpy(gl).to(dev), "l2")
    yield dict(gal=gal, g=gd, q=qd, gl=gl, ql=ql, dev=dev)
    gal.close()


def test_loader_normalisation_matches_oracle_sample(c2, port):
    raw_g, _, _, _ = synth.make_split(N, 8, D, C, "l2")
    want = port.normalize_rows("l2", raw_g[:2000])
    assert np.array_equal(bits(c2["g"][:2000].cpu().numpy()), bits(want))


def test_oracle_on_query_sample(c2, fir, port):
    g = c2["g"].cpu().numpy()
    sel = np.random.default_rng(0).choice(Q, 64, replace=False)
    qs = c2["q"][torch.from_numpy(sel).to(c2["dev"])]
    oi, od = port.topk("l2", g, qs.cpu().numpy(), 10, nthreads=os.cpu_count() or 1)
    for path in (fir.PATH_TENSOR, fir.PATH_EXACT):
        idx, dist = c2["gal"].search(qs, k=10, path=path)
        torch.cuda.synchronize()
        assert np.array_equal(idx.cpu().numpy(), oi) and np.array_equal(bits(dist.cpu().numpy()), bits(od))


def test_properties_on_all_queries(c2, fir):
    gal, q = c2["gal"], c2["q"]
    i10, d10 = gal.search(q, k=10, path=fir.PATH_TENSOR)
    st = gal.stats()
    i1, d1 = gal.search(q, k=1, path=fir.PATH_TENSOR)
    i10b, d10b = gal.search(q, k=10, path=fir.PATH_TENSOR)
    torch.cuda.synchronize()
    assert torch.equal(i10, i10b) and torch.equal(d10.view(torch.int32), d10b.view(torch.int32))        # idempotent
    assert torch.equal(i1[:, 0], i10[:, 0]) and torch.equal(d1[:, 0].view(torch.int32), d10[:, 0].view(torch.int32))
    assert bool((d10[:, 1:] >= d10[:, :-1]).all())                                                        # sorted by distance
    same = d10[:, 1:] == d10[:, :-1]
    assert bool((i10[:, 1:][same] > i10[:, :-1][same]).all())                                             # ties by index
    assert st["reserved"] == 0                                                                             # nothing needed the CUDA-core re-run
    # the tensor path agrees with the exact CUDA-core path on a 1024-query slice (exact path is ~50x slower)
    ie, de = gal.search(q[:1024], k=10, path=fir.PATH_EXACT)
    torch.cuda.synchronize()
    assert torch.equal(ie, i10[:1024]) and torch.equal(de.view(torch.int32), d10[:1024].view(torch.int32))
    # label agreement with the synthetic ground truth (clustered data: the nearest neighbour is of the query's class)
    gl = torch.from_numpy(c2["gl"]).to(c2["dev"])
    assert float((gl[i1[:, 0].long()] == torch.from_numpy(c2["ql"]).to(c2["dev"])).float().mean()) > 0.999


def test_class_min_folds_to_global_min(c2, fir):
    gal, q = c2["gal"], c2["q"][:256]
    mn, arg = gal.class_min(q)
    i1, d1 = gal.search(q, k=1, path=fir.PATH_TENSOR)
    torch.cuda.synchronize()
    best = mn.min(dim=1)
    assert torch.equal(best.values.view(torch.int32), d1[:, 0].view(torch.int32))
    rows = torch.arange(len(q), device=q.device)
    assert torch.equal(arg[rows, best.indices], i1[:, 0])


def test_two_shards_merge_to_unsharded(c2, fir):
    q = c2["q"][:2048]
    want_i, want_d = c2["gal"].search(q, k=10, path=fir.PATH_TENSOR)
    parts_i, parts_d = [], []
    for lo, hi in ((0, 43_210), (43_210, N)):
        sh = fir.Gallery(c2["g"][lo:hi].contiguous(), torch.from_numpy(c2["gl"][lo:hi]).to(c2["dev"]), "l2", index_offset=lo)
        i, d = sh.search(q, k=10, path=fir.PATH_TENSOR)
        torch.cuda.synchronize()
        parts_i.append(i)
        parts_d.append(d)
        sh.close()
    mi, md = fir.merge_topk(torch.stack(parts_d), torch.stack(parts_i))
    torch.cuda.synchronize()
    assert torch.equal(mi, want_i) and torch.equal(md.view(torch.int32), want_d.view(torch.int32))
