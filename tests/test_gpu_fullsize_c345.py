"""GPU parity at BASELINE.json's full sizes for configs[2..4] (C3: 1M x 1280 chi-square / KL with PNN class scores,
C4: directed enumeration over 1M x 512 with the reference's full pivot chain, C5: 10M x 512 Euclidean top-10): the
unmodified reference / the pinned port over the WHOLE gallery on a query sample, plus size-independent properties on
larger batches.  Galleries come from the counter-based generator on the device (the bench's inputs, tests/test_synth.py
pins device == host bits) and are copied to the host for the CPU side."""
import gc
import importlib
import os

import numpy as np
import pytest
import torch

from util import bits

pytestmark = pytest.mark.gpu
synth = importlib.import_module("fast-image-recognition_b200.synth")
SEED = 0x5EED0000
NT = os.cpu_count() or 1


def make(fir, n, nq, d, classes, metric):
    dev = torch.device("cuda", 0)
    relu = metric != "l2"
    g, gl = synth.synth_rows_device(synth.ROLE_GALLERY, 0, n, n, d, classes, SEED, relu=relu, device=dev)
    q, ql = synth.synth_rows_device(synth.ROLE_QUERY, 0, nq, nq, d, classes, SEED, relu=relu, device=dev)
    fir.normalize_rows(g, metric)
    fir.normalize_rows(q, metric)
    torch.cuda.synchronize()
    return g, gl, q, ql


def need_host_ram(gib):
    """The CPU side of these tests holds the whole gallery (and the reference's own copy of it): skip, do not die, on a small host."""
    try:
        import psutil
        free = psutil.virtual_memory().available / 2 ** 30
    except Exception:
        return
    if free < gib:
        pytest.skip("needs %d GiB of free host memory for the reference side, %.0f available" % (gib, free))


def release(*gals):
    for x in gals:
        x.close()
    gc.collect()
    torch.cuda.empty_cache()


def to_host(t, step=1 << 20):
    out = np.empty(tuple(t.shape), np.float32)
    for lo in range(0, t.shape[0], step):
        torch.from_numpy(out[lo:lo + step]).copy_(t[lo:lo + step])
    return out


@pytest.mark.parametrize("metric,nq_sample", [("chi2", 12), ("kl", 2)])
def test_c3_full_size_scores_and_neighbours(fir, port, metric, nq_sample):
    """C3: 1M x 1280 ReLU'd, L1-normalised features, 1000 classes.  PNN class scores (<= 1e-5 relative, the bar north_star
    states), predicted labels, per-class minima and the nearest neighbour (bit-exact) against the port over the full gallery;
    the batched approximate top-k path agrees with the exact tiles on a larger batch."""
    need_host_ram(16)
    n, d, C, var = 1_000_000, 1280, 1000, 2e-5
    g, gl, q, ql = make(fir, n, 512, d, C, metric)
    gal = fir.Gallery(g, gl, metric)
    gh, glh = to_host(g), gl.cpu().numpy()
    del g
    qs = q[:nq_sample]
    qh = qs.cpu().numpy()
    sc, lab = gal.pnn_scores(qs, var)
    mn, arg = gal.class_min(qs)
    i1, d1 = gal.search(qs, k=1, path=fir.PATH_EXACT)
    torch.cuda.synchronize()
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=NT) as ex:                  # the port's class reducers are single-threaded; ctypes releases the GIL
        parts = list(ex.map(lambda i: (port.pnn_div(metric, gh, glh, C, qh[i:i + 1], var), port.class_min(metric, gh, glh, C, qh[i:i + 1])), range(nq_sample)))
    psc = np.concatenate([p[0][0] for p in parts]); plab = np.concatenate([p[0][1] for p in parts])
    pmn = np.concatenate([p[1][0] for p in parts]); parg = np.concatenate([p[1][1] for p in parts])
    sc = sc.cpu().numpy()
    big = psc > psc.max(axis=1, keepdims=True) * 1e-30              # (scores that underflow towards 0 have no relative error to speak of)
    rel = np.abs(sc - psc)[big] / psc[big]
    assert rel.max() <= 1e-5, rel.max()                             # tolerance: north_star's "PNN scores within 1e-5 relative"
    assert np.array_equal(lab.cpu().numpy(), plab)
    assert np.array_equal(arg.cpu().numpy(), parg) and np.array_equal(bits(mn.cpu().numpy()), bits(pmn))
    best = pmn.argmin(axis=1)                                       # the nearest neighbour is the smallest class minimum (lowest class on ties = lowest row: class-major)
    assert np.array_equal(i1.cpu().numpy()[:, 0], parg[np.arange(nq_sample), best])
    assert np.array_equal(bits(d1.cpu().numpy()[:, 0]), bits(pmn[np.arange(nq_sample), best]))
    # batched top-10: the approximate tiles + exact rerank (PATH_AUTO for this shape) == the exact tiles, bit for bit
    ia, da = gal.search(q, k=10)
    used = gal.stats()["path_used"]
    ie, de = gal.search(q[:64], k=10, path=fir.PATH_EXACT)
    torch.cuda.synchronize()
    assert used == fir.PATH_APPROX
    assert torch.equal(ia[:64], ie) and torch.equal(da[:64].view(torch.int32), de.view(torch.int32))
    assert bool((da[:, 1:] >= da[:, :-1]).all())
    del gh
    release(gal)


def test_c4_full_size_directed_enumeration(fir, port, ref_l2):
    """C4: 1M x 512, 10 000 classes, 32 pivots, the reference's full max(5, 0.015 N) = 15 000-row chain built on the GPU; the
    verbatim DirectedEnumeration::recognize over the GPU-built pivots / pivot-distance rows / threshold answers a query sample
    identically (index, distance bits, below-threshold flag, number of distance evaluations) for three check budgets."""
    need_host_ram(12)
    n, d, C = 1_000_000, 512, 10_000
    g, gl, q, ql = make(fir, n, 4096, d, C, "l2")
    gal = fir.Gallery(g, gl, "l2")
    dem = fir.Dem(gal, pivot0=12345)
    assert dem.chain_rows == 15_000 and dem.n_pivots == 32
    gh, glh = to_host(g), gl.cpu().numpy()
    # latency mode at size, on a handle whose workspace is sized by this very call, with query counts that are not powers of two
    # (the streaming kernel runs 4 / 8 query slots and must drop the spare ones — it once wrote them past the end of the buffer)
    fresh = fir.Gallery(g, gl, "l2")
    for m in (3, 5, 7):
        si, sd = fresh.search(q[:m], k=1, path=fir.PATH_EXACT)
        smn, sarg = fresh.class_min(q[:m])
        ti, td = gal.search(q[:m], k=1, path=fir.PATH_TENSOR)
        torch.cuda.synchronize()
        assert torch.equal(si, ti) and torch.equal(sd.view(torch.int32), td.view(torch.int32))
        best = smn.min(dim=1)
        assert torch.equal(best.values.view(torch.int32), td[:, 0].view(torch.int32))
    fresh.close()
    del g
    rdem = ref_l2.dem_create_injected(gh, glh, dem.pivots, dem.P, float(dem.threshold))
    sample = np.linspace(0, 4095, 48).astype(np.int64)
    qh = q.cpu().numpy()[sample]
    for ratio in (0.025, 0.05, 0.5):
        M = int(ratio * n)
        idx, dist, below, evals = (x.cpu().numpy() for x in dem.search(q, M))
        ri, rd, rb, re_ = rdem.search(qh, M, nthreads=NT)[:4]
        assert np.array_equal(ri, idx[sample])
        assert np.array_equal(bits(np.asarray(rd, np.float32)), bits(dist[sample]))
        assert np.array_equal(np.asarray(rb).astype(np.uint8), below[sample])
        assert np.array_equal(np.asarray(re_).astype(np.int32), evals[sample])
    rdem.close()
    dem.close()
    del gh
    release(gal)


def test_c5_full_size_topk(fir, port, ref_l2):
    """C5: 10M x 512 L2-normalised embeddings, top-10 for 25 600 queries on the tensor path (full rounds with re-aligned pairs
    and a phased remainder: 100 query blocks on 74 pairs).  The unmodified reference's BruteForce::recognize (top-1) and the
    pinned port (top-10) over the whole gallery on a query sample; on all queries: k = 1 is the head of k = 10, distances
    ascend, ties ascend by index, the call is idempotent, and nothing needed the CUDA-core re-run."""
    need_host_ram(52)
    n, d, C, nq = 10_000_000, 512, 1000, 25_600
    g, gl, q, ql = make(fir, n, nq, d, C, "l2")
    gal = fir.Gallery(g, gl, "l2")
    i10, d10 = gal.search(q, k=10, path=fir.PATH_TENSOR)
    st = gal.stats()
    i1, d1 = gal.search(q, k=1, path=fir.PATH_TENSOR)
    i10b, d10b = gal.search(q, k=10, path=fir.PATH_TENSOR)
    torch.cuda.synchronize()
    assert st["path_used"] == fir.PATH_TENSOR and st["reserved"] == 0
    assert torch.equal(i10, i10b) and torch.equal(d10.view(torch.int32), d10b.view(torch.int32))
    assert torch.equal(i1[:, 0], i10[:, 0]) and torch.equal(d1[:, 0].view(torch.int32), d10[:, 0].view(torch.int32))
    assert bool((d10[:, 1:] >= d10[:, :-1]).all())
    same = d10[:, 1:] == d10[:, :-1]
    assert bool((i10[:, 1:][same] > i10[:, :-1][same]).all())
    assert float((gl[i1[:, 0].long()] == ql).float().mean()) > 0.999
    gh = to_host(g)
    del g
    sample = np.linspace(0, nq - 1, 24).astype(np.int64)               # covers full-round and remainder query blocks
    qh = q.cpu().numpy()[sample]
    ri, rd = ref_l2.bf(gh, qh, None, nthreads=NT)[:2]
    assert np.array_equal(ri, i1.cpu().numpy()[sample, 0]) and np.array_equal(bits(np.asarray(rd, np.float32)), bits(d1.cpu().numpy()[sample, 0]))
    ti, td = port.topk("l2", gh, qh[:8], 10, nthreads=NT)
    assert np.array_equal(ti, i10.cpu().numpy()[sample[:8]]) and np.array_equal(bits(td), bits(d10.cpu().numpy()[sample[:8]]))
    del gh
    release(gal)
