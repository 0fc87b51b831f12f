"""CPU: randomised pinning of the plain-C restatement (oracle/fir_oracle.c) against the unmodified reference (oracle/_ref) on
ragged shapes the fixed cases do not hit — distances over arbitrary windows, brute force with ties, the TWD classifiers with
chunks that do not divide 256, sequential PNN.  Small sizes, a few dozen draws each."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from util import bits

SET = settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture], derandomize=True)


def _rows(seed, n, d, metric, dup=0):
    r = np.random.default_rng(seed)
    x = r.normal(size=(n, d)).astype(np.float32)
    if metric != "l2":
        x = np.maximum(x, 0)                         # ReLU-style features: exact zeros exercise the (l+r)>0 / l>0 guards
    if dup:
        x[n - dup:] = x[:dup]                        # exact duplicates: ties
    return x


@SET
@given(seed=st.integers(0, 10**6), d=st.integers(1, 70), a=st.integers(0, 69), b=st.integers(1, 70), metric=st.sampled_from(["l2", "chi2", "kl"]))
def test_distance_over_any_window(port, request, seed, d, a, b, metric):
    ref = request.getfixturevalue("ref_" + metric)
    lo, hi = min(a, d - 1), min(max(b, min(a, d - 1) + 1), d)
    x = _rows(seed, 2, d, metric)
    assert bits(port.distance(metric, x[0], x[1], lo, hi)) == bits(ref.distance(x[0], x[1], lo, hi))


@SET
@given(seed=st.integers(0, 10**6), n=st.integers(1, 60), nq=st.integers(1, 9), d=st.integers(1, 40), dup=st.integers(0, 5),
       metric=st.sampled_from(["l2", "chi2", "kl"]))
def test_brute_force_with_ties(port, request, seed, n, nq, d, dup, metric):
    ref = request.getfixturevalue("ref_" + metric)
    g = port.normalize_rows(metric, _rows(seed, n, d, metric, min(dup, n // 2)) + (0.01 if metric != "l2" else 0))
    q = np.concatenate([g[: min(nq, n)], port.normalize_rows(metric, _rows(seed + 1, nq, d, metric) + (0.01 if metric != "l2" else 0))])
    pi, pd = port.bf(metric, g, q)
    ri, rd = ref.bf(g, q)
    assert np.array_equal(pi, ri) and np.array_equal(bits(pd), bits(rd))


@SET
@given(seed=st.integers(0, 10**6), n=st.integers(6, 50), nq=st.integers(1, 8), fc=st.integers(1, 255), c=st.integers(5, 8),
       th=st.sampled_from([0.5, 0.7, 0.9, 1.3]), kind=st.sampled_from(["posteriors", "diff", "ratio"]))
def test_twd_classifiers(port, ref_l2, seed, n, nq, fc, c, th, kind):
    d = ((256 + fc - 1) // fc) * fc                  # the last proposed-TWD chunk may run past 256 (ImageTesting.cpp:223)
    d = max(d, 256)
    g = port.normalize_rows("l2", _rows(seed, n, d, "l2"))
    q = port.normalize_rows("l2", _rows(seed + 7, nq, d, "l2") * 0.3 + g[np.arange(nq) % n])
    gl = (np.arange(n) % c).astype(np.int32)
    idx, cls, unrel = port.twd_proposed("l2", g, gl, q, fc, th)
    rc, ru = ref_l2.twd("proposed", g, gl, c, q, fc, th)
    assert np.array_equal(cls, rc) and np.array_equal(unrel, ru)
    idx, cls, unrel = port.twd_conventional("l2", g, gl, c, q, kind, th * 0.01 if kind == "diff" else th * 0.5, fc)
    rc, ru = ref_l2.twd("conventional", g, gl, c, q, fc, th * 0.01 if kind == "diff" else th * 0.5, kind)
    assert np.array_equal(cls, rc) and np.array_equal(unrel, ru)


@settings(max_examples=8, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture], derandomize=True)
@given(seed=st.integers(0, 10**6), d=st.integers(20, 90), c=st.integers(2, 6), per=st.integers(2, 9))
def test_sequential_pnn_and_fpnn(port, ref_l2, seed, d, c, per):
    r = np.random.default_rng(seed)
    n = c * (per + 4)
    cen = r.normal(size=(c, d))
    labels = (np.arange(n) % c).astype(np.int32)
    rows = cen[labels] + r.normal(scale=0.8, size=(n, d))
    rows /= np.linalg.norm(rows, axis=1, keepdims=True)
    tr, trl, te, avg = ref_l2.cls_setup(rows, labels, c, per, seed=seed % 1000)
    if len(te) == 0:
        return
    assert np.array_equal(port.pnn_seq(rows[tr], trl, c, avg, rows[te]), ref_l2.cls_pnn_seq(0, len(te)))
    sd = ref_l2.cls_std()
    lab, a, J = ref_l2.cls_fpnn(0, len(te), scale=1.0, bf=False, output_ratio=0.95, coefficients=True)
    pa, pJ = port.fpnn_train(rows[tr], trl, c, avg, sd, 1.0)
    assert pJ == J and np.array_equal(pa.view(np.int64), a.view(np.int64))
    assert np.array_equal(port.fpnn_predict(pa, pJ, c, avg, sd, rows[te], 1.0, sequential=True, output_ratio=0.95), lab)


@settings(max_examples=10, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture], derandomize=True)
@given(seed=st.integers(0, 10**6), n=st.integers(340, 1500), nq=st.integers(1, 25), d=st.integers(4, 48), c=st.integers(2, 25),
       metric=st.sampled_from(["l2", "chi2", "kl"]), m_frac=st.sampled_from([0.0, 0.02, 0.2, 1.0]), thr_scale=st.sampled_from([1.0, 0.3, 0.02]))
def test_directed_enumeration(port, request, seed, n, nq, d, c, metric, m_frac, thr_scale):
    """DirectedEnumeration: verbatim constructor + recognize against the restatement — pivot chain, pivot-distance rows,
    threshold, then the candidate walk with the state injected at a scaled threshold (early exits ... full budgets)."""
    ref = request.getfixturevalue("ref_" + metric)
    r = np.random.default_rng(seed)
    labels = np.sort(r.integers(0, c, size=n)).astype(np.int32)
    cen = r.normal(size=(c, d)).astype(np.float32)
    x = cen[labels] + r.normal(scale=0.7, size=(n, d)).astype(np.float32)
    qx = cen[r.integers(0, c, size=nq)] + r.normal(scale=0.7, size=(nq, d)).astype(np.float32)
    if metric != "l2":
        x, qx = np.abs(x) + np.float32(1e-3), np.abs(qx) + np.float32(1e-3)
    g, q = port.normalize_rows(metric, x), port.normalize_rows(metric, qx)
    rd = ref.dem_create(g, labels, seed=seed % 977 + 1)
    pb = port.dem_build(metric, g, labels, int(rd.pivots[0]))
    assert pb["n_pivots"] == rd.n_pivots and np.array_equal(pb["pivots"][: rd.n_pivots], rd.pivots)
    assert np.array_equal(bits(pb["P"]), bits(rd.P())) and bits(pb["threshold"]) == bits(np.float32(rd.threshold))
    M = int(m_frac * n)
    thr = float(rd.threshold) * thr_scale
    inj = rd if thr_scale == 1.0 else ref.dem_create_injected(g, labels, rd.pivots, rd.P(), thr)
    for a, b in zip(inj.search(q, M), port.dem_search(metric, g, rd.pivots, rd.P(), thr, M, q)):
        assert np.array_equal(a, b)
    if inj is not rd:
        inj.close()
    rd.close()
