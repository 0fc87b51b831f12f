"""The counter-based workload generator (csrc/synth_common.h): numpy restatement == host C twin (CPU), == CUDA kernel (GPU).
bench.py relies on this: rank r materialises rows [lo, hi) of the gallery on its GPU, the reference arm regenerates the
same bits on the host."""
import importlib

import numpy as np
import pytest

synth = importlib.import_module("fast-image-recognition_b200.synth")


def test_philox_known_answer():
    # Random123 known-answer vectors for philox4x32-10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = synth.philox4x32_10(*[np.array([c]) for c in ctr], key[0], key[1])
        assert tuple(int(g[0]) for g in got) == want


@pytest.mark.parametrize("role,relu,d", [(0, False, 512), (1, False, 77), (0, True, 1280), (1, True, 33)])
def test_host_twin_equals_numpy(role, relu, d):
    a, la = synth.synth_rows_numpy(role, 12345, 257, 1_000_000, d, 50, seed=7, relu=relu)
    b, lb = synth.synth_rows_host(role, 12345, 257, 1_000_000, d, 50, seed=7, relu=relu, threads=3)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and np.array_equal(la, lb)
    if relu:
        assert (a >= 0).all() and (a == 0).mean() > 0.2


def test_rows_do_not_depend_on_the_range_asked_for():
    whole, lw = synth.synth_rows_host(0, 0, 1000, 1000, 64, 10, seed=3)
    part, lp = synth.synth_rows_host(0, 600, 400, 1000, 64, 10, seed=3)
    assert np.array_equal(whole[600:], part) and np.array_equal(lw[600:], lp)
    assert (np.diff(lw) >= 0).all() and lw[0] == 0 and lw[-1] == 9            # class-major, equal blocks
    z = synth._z(1, np.arange(4000), 64)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01


def test_caltech_shape():
    s = synth.caltech_sizes()
    assert len(s) == 101 and sum(s) == 8677 and sum(1 for x in s if x > 400) == 4 and min(s) >= 31
    assert sum(min(x, 400) for x in s) - 30 * 101 == 4779                      # queries left after 30 per class with the 400 cap


@pytest.mark.gpu
@pytest.mark.parametrize("role,relu,d", [(0, False, 512), (1, True, 1280), (0, False, 77)])
def test_device_generator_equals_host_twin(role, relu, d):
    import torch
    dev, lab = synth.synth_rows_device(role, 999_000, 3000, 10_000_000, d, 1000, seed=0x5EED0000, relu=relu)
    torch.cuda.synchronize()
    host, hl = synth.synth_rows_host(role, 999_000, 3000, 10_000_000, d, 1000, seed=0x5EED0000, relu=relu)
    assert np.array_equal(dev.cpu().numpy().view(np.uint32), host.view(np.uint32))
    assert np.array_equal(lab.cpu().numpy(), hl)
