"""GPU: persisted index (fir_index_save / fir_index_load): a loaded gallery + DirectedEnumeration state answers bit-identically
to the one saved; damaged files are refused."""
import os

import numpy as np
import pytest

from util import bits, make_data

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("metric", ["l2", "chi2"])
def test_index_round_trip(fir, port, tmp_path, metric):
    g, gl, q, ql = make_data(port, metric, 700, 90, 72, 9, seed=12)      # d not a multiple of the row padding
    gal = fir.Gallery(g, gl, metric)
    dem = fir.Dem(gal, pivot0=5)
    path = str(tmp_path / "index.firb200")
    fir.save_index(path, gal, dem)
    assert os.path.getsize(path) == 72 + 700 * 4 + 700 * 72 * 4 + dem.n_pivots * 4 + dem.n_pivots * 700 * 4 + 8
    gal2, dem2 = fir.load_index(path)
    assert (gal2.n, gal2.d, gal2.n_classes, gal2.metric_name) == (700, 72, gal.n_classes, metric)
    assert dem2 is not None and dem2.n_pivots == dem.n_pivots and bits(dem2.threshold) == bits(dem.threshold)
    assert np.array_equal(dem2.pivots, dem.pivots) and np.array_equal(bits(dem2.P), bits(dem.P))
    for k in (1, 5):
        a, b = gal.search(q, k=k), gal2.search(q, k=k)
        assert np.array_equal(a[0], b[0]) and np.array_equal(bits(a[1]), bits(b[1]))
    for M in (0, 40, 300):
        for x, y in zip(dem.search(q, M), dem2.search(q, M)):
            assert np.array_equal(x, y)
    mn, arg = gal.class_min(q)
    mn2, arg2 = gal2.class_min(q)
    assert np.array_equal(arg, arg2) and np.array_equal(bits(mn), bits(mn2))
    for h in (dem2, gal2, dem, gal):
        h.close()


def test_index_without_dem_and_damage(fir, port, tmp_path):
    g, gl, q, ql = make_data(port, "l2", 300, 20, 64, 5, seed=3)
    gal = fir.Gallery(g, gl, "l2")
    path = str(tmp_path / "gallery_only.firb200")
    fir.save_index(path, gal)
    gal2, dem2 = fir.load_index(path)
    assert dem2 is None
    assert np.array_equal(gal.search(q, k=3)[0], gal2.search(q, k=3)[0])
    gal2.close()
    blob = bytearray(open(path, "rb").read())
    cases = {"flipped payload byte": bytes(blob[:5000]) + bytes([blob[5000] ^ 0x40]) + bytes(blob[5001:]),
             "truncated": bytes(blob[:-100]),
             "bad magic": b"NOTANIDX" + bytes(blob[8:]),
             "header claims more rows": bytes(blob[:16]) + (10 ** 6).to_bytes(8, "little") + bytes(blob[24:]),
             "empty": b""}
    for name, data in cases.items():
        bad = str(tmp_path / "bad.firb200")
        open(bad, "wb").write(data)
        with pytest.raises(fir.FirError):
            fir.load_index(bad)
    with pytest.raises(fir.FirError):
        fir.load_index(str(tmp_path / "does_not_exist"))
    other = fir.Gallery(g[:100], gl[:100], "l2")
    dem = fir.Dem(gal, pivot0=1)
    with pytest.raises(fir.FirError):
        fir.save_index(str(tmp_path / "mismatch"), other, dem)       # DEM of another gallery
    for h in (dem, other, gal):
        h.close()
