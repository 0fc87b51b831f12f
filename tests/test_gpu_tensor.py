"""GPU parity: the tcgen05/TMA tensor-core L2 path (candidates + exact rerank + certificate) against the
oracle, through the C-ABI.  Indices and distances must be bit-identical to the reference argmin / top-k."""
import os

import numpy as np
import pytest

from util import bits, make_data

pytestmark = pytest.mark.gpu

SHAPES = [  # n, nq, d, classes
    (3000, 300, 512, 30),     # A resident in shared memory (D <= 512)
    (1000, 130, 128, 10),
    (777, 45, 100, 7),        # ragged everything: D % 64 != 0, N % 256 != 0, Q % 128 != 0
    (200, 5, 64, 4),          # gallery smaller than one tile
    (2500, 140, 1536, 25),    # reference default D (db.h:86): A streamed per k-block
    (700, 20100, 64, 8),      # more query blocks than SMs: full wave-synchronous rounds + a remainder round
]


@pytest.mark.parametrize("n,nq,d,c", SHAPES)
def test_tensor_topk_matches_oracle(fir, port, n, nq, d, c):
    g, gl, q, ql = make_data(port, "l2", n, nq, d, c, seed=n % 7)
    gal = fir.Gallery(g, gl, "l2")
    for k in (1, 10, 20):
        idx, dist = gal.search(q, k=k, path=fir.PATH_TENSOR)
        st = gal.stats()
        assert st["path_used"] == fir.PATH_TENSOR
        oi, od = port.topk("l2", g, q, k, nthreads=8)
        assert np.array_equal(idx, oi), "k=%d fallback=%d" % (k, st["n_fallback"])
        assert np.array_equal(bits(dist), bits(od))
    gal.close()


def test_tensor_matches_reference_build(fir, port, ref_l2):
    g, gl, q, ql = make_data(port, "l2", 4000, 256, 512, 40, seed=5)
    gal = fir.Gallery(g, gl, "l2")
    idx, dist = gal.search(q, k=1, path=fir.PATH_TENSOR)
    ri, rd = ref_l2.bf(g, q, gl, nthreads=8)
    assert np.array_equal(idx[:, 0], ri) and np.array_equal(bits(dist[:, 0]), bits(rd))
    assert gal.stats()["n_fallback"] == 0          # clustered data: every query certified by the bound
    gal.close()


def test_tensor_duplicates_fall_back_exactly(fir, port):
    """Mass exact ties defeat any finite candidate list: the certificate must fail and the exact re-run must
    still return the reference answer (lowest index first)."""
    g, gl, q, ql = make_data(port, "l2", 600, 64, 128, 3, seed=9)
    g[:] = g[:40].repeat(15, axis=0)               # 15 copies of each of 40 rows
    gal = fir.Gallery(g, None, "l2")
    idx, dist = gal.search(q, k=10, path=fir.PATH_TENSOR)
    oi, od = port.topk("l2", g, q, 10)
    assert np.array_equal(idx, oi) and np.array_equal(bits(dist), bits(od))
    gal.close()


def test_tensor_iid_noise_and_unnormalised(fir, port):
    rng = np.random.default_rng(3)
    g = rng.standard_normal((5000, 256)).astype(np.float32)          # i.i.d., NOT normalised, |x| up to ~5
    q = rng.standard_normal((200, 256)).astype(np.float32)
    gal = fir.Gallery(g, None, "l2")
    idx, dist = gal.search(q, k=5, path=fir.PATH_TENSOR)
    oi, od = port.topk("l2", g, q, 5, nthreads=8)
    assert np.array_equal(idx, oi) and np.array_equal(bits(dist), bits(od))
    gal.close()


def test_tensor_error_bound_is_respected(fir, port):
    """Evidence for the certificate: the measured |approx - exact| of every candidate stays below the bound E."""
    g, gl, q, ql = make_data(port, "l2", 8192, 256, 512, 64, seed=2)
    gal = fir.Gallery(g, gl, "l2")
    gal.search(q, k=10, path=fir.PATH_TENSOR)
    st = gal.stats()
    ci, ca, ce = gal.debug_candidates(q.shape[0])
    ok = ci >= 0
    err = np.abs(ca[ok].astype(np.float64) - ce[ok].astype(np.float64) * 512)
    assert ok.sum() > 1000
    assert err.max() < st["approx_err_bound"], (err.max(), st["approx_err_bound"])
    assert err.max() < 0.25 * st["approx_err_bound"]     # the Cauchy-Schwarz bound is loose by a wide margin
    print("max |approx-exact| = %.3g, certified bound E = %.3g, fallback = %d" % (err.max(), st["approx_err_bound"], st["n_fallback"]))
    gal.close()


def test_tensor_device_pointers(fir, port):
    import torch
    g, gl, q, ql = make_data(port, "l2", 2048, 128, 512, 16, seed=4)
    gal = fir.Gallery(torch.from_numpy(g).cuda(), torch.from_numpy(gl).cuda(), "l2")
    gal.set_stream(torch.cuda.current_stream().cuda_stream)
    idx, dist = gal.search(torch.from_numpy(q).cuda(), k=10, path=fir.PATH_TENSOR)
    torch.cuda.synchronize()
    oi, od = port.topk("l2", g, q, 10, nthreads=8)
    assert np.array_equal(idx.cpu().numpy(), oi) and np.array_equal(bits(dist.cpu().numpy()), bits(od))
    gal.close()


@pytest.mark.parametrize("env", [{"FIR_TENSOR_SEED": "0"}, {"FIR_TENSOR_SEED_M": "1"}, {"FIR_TENSOR_CTAS": "1"},
                                 {"FIR_TENSOR_PHASED_MIN_BYTES": "0", "FIR_TENSOR_SYNC_TILES": "8", "FIR_TEST_NQ": "24220"},
                                 {"FIR_TENSOR_PHASED_MIN_BYTES": "0", "FIR_TENSOR_SYNC_TILES": "8", "FIR_TEST_NQ": "19000", "FIR_TENSOR_R1": "8"}])
def test_tensor_path_variants_stay_exact(env):
    """The list seeds are an optimisation, never a correctness input: with the sample pass switched off, with a seed rank that
    is deliberately too small (most queries then fail the first certificate and take the second pass / the exact re-run), and
    with the single-CTA kernel, the tensor path still returns the exact kernel's answer bit for bit.  The last two run what a
    gallery larger than L2 gets — full rounds with the pairs re-aligned every few tiles, the remainder's query blocks in phases
    of shared tile ranges (21 and 1 remainder blocks on 74 pairs) — at a size the exact kernel can check in full.  (The switches
    are read once per process, hence the subprocess.)"""
    import subprocess
    import sys
    code = r'''
import importlib, sys, numpy as np, torch
sys.path.insert(0, %r)
import fir_b200
synth = importlib.import_module("fast-image-recognition_b200.synth")
import os
g, gl, q, ql = synth.make_split(60000, int(os.environ.get("FIR_TEST_NQ", "1536")), 512, 300, "l2", seed=4)
dev = torch.device("cuda", 0)
gd, qd = torch.from_numpy(g).to(dev), torch.from_numpy(q).to(dev)
fir_b200.normalize_rows(gd, "l2"); fir_b200.normalize_rows(qd, "l2")
gal = fir_b200.Gallery(gd, torch.from_numpy(gl).to(dev), "l2")
for k in (1, 10):
    ti, td = gal.search(qd, k=k, path=fir_b200.PATH_TENSOR)
    st = gal.stats()
    ei, ed = gal.search(qd, k=k, path=fir_b200.PATH_EXACT)
    assert torch.equal(ti, ei) and torch.equal(td.view(torch.int32), ed.view(torch.int32)), k
    print("k", k, "flagged", st["n_fallback"], "exact", st["reserved"])
print("OK")
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr


@pytest.mark.parametrize("n,nq,d", [(20000, 1300, 200), (9000, 700, 33), (12000, 500, 62), (12000, 500, 63)])
def test_tensor_folded_norms(fir, port, n, nq, d):
    """With at least two spare columns in the last k-block (D mod 64 <= 62) the gallery's row norms ride in the shadow and the
    epilogue compares raw accumulators (fold_norm_kernel): same bits as the port for un-normalised rows of mixed magnitude; queries
    whose scale cannot be folded (1e-8 x and 1e+7 x the gallery's magnitude: U leaves the fp16 range) come back exact through the
    fallback; D = 63 has one spare column and keeps the norm loads."""
    rng = np.random.default_rng(d)
    g = (rng.standard_normal((n, d)) * rng.uniform(0.2, 3.0, (n, 1))).astype(np.float32)       # row norms spread over a decade
    q = (rng.standard_normal((nq, d)) * rng.uniform(0.2, 3.0, (nq, 1))).astype(np.float32)
    q[5] *= np.float32(1e-8); q[6] *= np.float32(1e7); q[7] = 0
    g[100:103] = g[7]
    gal = fir.Gallery(g, None, "l2")
    for k in (1, 10):
        idx, dist = gal.search(q, k=k, path=fir.PATH_TENSOR)
        assert gal.stats()["path_used"] == fir.PATH_TENSOR
        oi, od = port.topk("l2", g, q, k, nthreads=os.cpu_count() or 1)
        assert np.array_equal(idx, oi) and np.array_equal(bits(dist), bits(od)), k
    gal.close()


@pytest.mark.parametrize("n,nq,d", [(30000, 1500, 512), (20011, 900, 200)])
def test_tensor_prefix_max_features(fir, port, n, nq, d):
    """recognize_image_bf's prefix distance (db_features.cpp:319-335: the first max_features dimensions, mean over them) on the
    tensor path: the k-blocks of the shadow up to the prefix, prefix row norms, queries packed with zeros beyond the prefix —
    same bits as the port's prefix brute force (top-1) and as the exact kernels (top-5), for prefixes that end inside a k-block,
    on a k-block boundary and one short of D; the cached prefix norms follow the prefix from call to call."""
    g, gl, q, ql = make_data(port, "l2", n, nq, d, 40, seed=3)
    g[100:104] = g[7]                                            # exact ties: lowest index wins on every path
    gal = fir.Gallery(g, gl, "l2")
    for mf in (64, 100, d - 1, 37, 128):
        ti, td = gal.search(q, k=5, max_features=mf, path=fir.PATH_TENSOR)
        assert gal.stats()["path_used"] == fir.PATH_TENSOR
        ei, ed = gal.search(q, k=5, max_features=mf, path=fir.PATH_EXACT)
        assert np.array_equal(ti, ei) and np.array_equal(bits(td), bits(ed)), mf
        oi, od = port.bf("l2", g, q[:200], max_features=mf, nthreads=os.cpu_count() or 1)
        assert np.array_equal(ti[:200, 0], oi) and np.array_equal(bits(td[:200, 0]), bits(od)), mf
    ti, td = gal.search(q, k=5, path=fir.PATH_TENSOR)             # and back to all dimensions
    ei, ed = gal.search(q, k=5, path=fir.PATH_EXACT)
    assert np.array_equal(ti, ei) and np.array_equal(bits(td), bits(ed))
    gal.close()


@pytest.mark.parametrize("n,nq,d,c", [(30000, 700, 512, 60), (20011, 1100, 200, 333), (70000, 300, 96, 7)])
def test_class_min_on_tensor_cores_matches_oracle(fir, port, n, nq, d, c):
    """fir_class_min (Euclidean, large batch) through the tcgen05 candidate passes: per-class approximate minima in the epilogue
    (run-length over the class-major rows), a second pass listing every row within the certified margin of its class minimum,
    exact fp32 rerank — bit-equal to the exact kernels' answer, including classes with exact ties across rows (lowest index),
    a class with more tied rows than the list holds (→ that query re-runs on the exact tiles), empty classes, ragged sizes."""
    g, gl, q, ql = make_data(port, "l2", n, nq, d, c, seed=n % 13)
    g[50:58] = g[40]                                     # 9 identical rows: more than the four list slots per class
    g[n - 30:n - 27] = g[n - 31]                         # a tie of four
    q[0], q[1] = g[40], g[n - 31]                        # queries sitting exactly on the tied rows
    keep = gl != 3                                       # class 3 has no rows at all
    g, gl = g[keep], gl[keep]
    gal = fir.Gallery(g, gl, "l2")
    mn, arg = gal.class_min(q)
    assert gal.stats()["path_used"] == fir.PATH_TENSOR
    omn, oarg = port.class_min("l2", g, gl, gal.n_classes, q)
    assert np.array_equal(arg, oarg) and np.array_equal(bits(mn), bits(omn))
    import torch
    mn2, arg2 = gal.class_min(torch.from_numpy(q).cuda())            # device buffers
    torch.cuda.synchronize()
    assert np.array_equal(arg2.cpu().numpy(), oarg) and np.array_equal(bits(mn2.cpu().numpy()), bits(omn))
    gal.close()
    # a gallery that is not class-major is served by the exact kernels
    perm = np.random.default_rng(0).permutation(len(g))
    gal = fir.Gallery(g[perm], gl[perm], "l2")
    mn, arg = gal.class_min(q[:300])
    omn, oarg = port.class_min("l2", g[perm], gl[perm], gal.n_classes, q[:300])
    assert np.array_equal(arg, oarg) and np.array_equal(bits(mn), bits(omn))
    gal.close()
