"""GPU parity: the sequential three-way-decision classifiers of qt_cpp/ImageTesting.cpp against the
unmodified reference classes (oracle/_ref) and the C port, through the C-ABI."""
import numpy as np
import pytest

from util import make_data

pytestmark = pytest.mark.gpu

CONV = [("posteriors", 0.24), ("diff", 0.003), ("ratio", 0.7), ("diff", 0.0003), ("ratio", 0.9), ("posteriors", 0.21)]
PROP = [(32, 0.7), (64, 0.7), (32, 0.9), (16, 0.5)]


def _ref(request, metric):
    return request.getfixturevalue("ref_" + metric)


@pytest.mark.parametrize("metric,sigma", [("l2", 0.5), ("l2", 1.5), ("l2", 3.0), ("chi2", 1.5), ("kl", 1.5)])
def test_twd_matches_reference_build(fir, port, request, metric, sigma):
    ref = _ref(request, metric)
    g, gl, q, ql = make_data(port, metric, 600, 300, 256, 12, seed=3, sigma=sigma)
    gal = fir.Gallery(g, gl, metric)
    unreliable_seen = 0
    for fc, th in PROP:
        idx, lab, unrel = gal.twd_proposed(q, fc, th)
        rc, ru = ref.twd("proposed", g, gl, 12, q, fc, th)
        pi, pc, pu = port.twd_proposed(metric, g, gl, q, fc, th)
        assert np.array_equal(lab, rc) and np.array_equal(unrel, ru), (fc, th)
        assert np.array_equal(idx, pi) and np.array_equal(lab, pc) and np.array_equal(unrel, pu), (fc, th)
        unreliable_seen += int(unrel.sum())
    for kind, th in CONV:
        idx, lab, unrel = gal.twd_conventional(q, kind, th, 64)
        rc, ru = ref.twd("conventional", g, gl, 12, q, 64, th, kind)
        pi, pc, pu = port.twd_conventional(metric, g, gl, 12, q, kind, th, 64)
        assert np.array_equal(lab, rc) and np.array_equal(unrel, ru), (kind, th)
        assert np.array_equal(idx, pi) and np.array_equal(lab, pc) and np.array_equal(unrel, pu), (kind, th)
        unreliable_seen += int(unrel.sum())
    assert unreliable_seen > 0
    gal.close()


def test_twd_chunk_past_last_feature_and_ragged(fir, port, ref_l2):
    """feat_count = 96 does not divide 256: the third chunk covers [192, 288) (ImageTesting.cpp:223,243); ragged sizes."""
    g, gl, q, ql = make_data(port, "l2", 333, 77, 300, 7, seed=9, sigma=2.0)
    gal = fir.Gallery(g, gl, "l2")
    idx, lab, unrel = gal.twd_proposed(q, 96, 0.8)
    rc, ru = ref_l2.twd("proposed", g, gl, 7, q, 96, 0.8)
    assert np.array_equal(lab, rc) and np.array_equal(unrel, ru)
    assert np.array_equal(idx, port.twd_proposed("l2", g, gl, q, 96, 0.8)[0])
    for fc in (1, 100, 255):
        idx, lab, unrel = gal.twd_conventional(q, "ratio", 0.8, fc)
        rc, ru = ref_l2.twd("conventional", g, gl, 7, q, fc, 0.8, "ratio")
        assert np.array_equal(lab, rc) and np.array_equal(unrel, ru), fc
    gal.close()


def test_twd_device_queries_and_threshold_above_one(fir, port):
    """th > 1 makes 1/th < 1: the best match itself is dropped and the walk stops with bestInd kept (ImageTesting.cpp:256-263)."""
    import torch
    g, gl, q, ql = make_data(port, "l2", 500, 130, 256, 9, seed=4, sigma=2.5)
    gal = fir.Gallery(g, gl, "l2")
    qd = torch.from_numpy(q).cuda()
    for fc, th in ((32, 1.5), (64, 0.7)):
        idx, lab, unrel = gal.twd_proposed(qd, fc, th)
        pi, pc, pu = port.twd_proposed("l2", g, gl, q, fc, th)
        assert np.array_equal(idx.cpu().numpy(), pi) and np.array_equal(lab.cpu().numpy(), pc) and np.array_equal(unrel.cpu().numpy(), pu)
    idx, lab, unrel = gal.twd_conventional(qd, "diff", 0.001)
    pi, pc, pu = port.twd_conventional("l2", g, gl, 9, q, "diff", 0.001)
    assert np.array_equal(idx.cpu().numpy(), pi) and np.array_equal(unrel.cpu().numpy(), pu)
    gal.close()


def test_twd_argument_errors(fir, port):
    g, gl, q, ql = make_data(port, "l2", 100, 5, 128, 4, seed=1)
    gal = fir.Gallery(g, gl, "l2")
    with pytest.raises(fir.FirError):
        gal.twd_proposed(q, 32, 0.7)                    # last_feature 256 > D
    with pytest.raises(fir.FirError):
        gal.twd_conventional(q, "posteriors", 0.24, 32, last_feature=128)    # < 5 classes
    with pytest.raises(fir.FirError):
        gal.twd_conventional(q, "diff", 0.003, 128, last_feature=128)        # empty refinement range
    idx, lab, unrel = gal.twd_proposed(q, 32, 0.7, last_feature=128)
    assert np.array_equal(idx, port.twd_proposed("l2", g, gl, q, 32, 0.7, last_feature=128)[0])
    gal.close()


def test_twd_larger_gallery_matches_port_on_a_query_sample(fir, port):
    """20k x 320 gallery, 3k queries on the device (several 64-query tiles, compaction across chunks); the port checks a
    sample of the queries, the rest through properties: unreliable => more than one chunk was needed, decided class = class of
    the returned row."""
    import torch
    g, gl, q, ql = make_data(port, "l2", 20000, 3000, 320, 200, seed=21, sigma=2.0)
    dev = torch.device("cuda", 0)
    gal = fir.Gallery(torch.from_numpy(g).to(dev), torch.from_numpy(gl).to(dev), "l2")
    qd = torch.from_numpy(q).to(dev)
    sample = np.arange(0, 3000, 47)
    for fc, th in ((32, 0.7), (64, 0.8)):
        idx, lab, unrel = (t.cpu().numpy() for t in gal.twd_proposed(qd, fc, th))
        pi, pc, pu = port.twd_proposed("l2", g, gl, q[sample], fc, th)
        assert np.array_equal(idx[sample], pi) and np.array_equal(lab[sample], pc) and np.array_equal(unrel[sample], pu)
        assert np.array_equal(lab, gl[idx]) and idx.min() >= 0
    for kind, th in (("ratio", 0.8), ("diff", 0.0005), ("posteriors", 0.22)):
        idx, lab, unrel = (t.cpu().numpy() for t in gal.twd_conventional(qd, kind, th, 64))
        pi, pc, pu = port.twd_conventional("l2", g, gl, 200, q[sample], kind, th, 64)
        assert np.array_equal(idx[sample], pi) and np.array_equal(unrel[sample], pu), kind
        assert np.array_equal(lab, gl[idx])
        assert 0 < unrel.mean() < 1, kind                  # both branches of the decision are exercised
    gal.close()
