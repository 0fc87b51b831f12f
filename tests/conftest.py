import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def port():
    from oracle import oracle_py
    oracle_py.build()
    return oracle_py.Port()


def _ref(metric):
    from oracle import oracle_py
    if not oracle_py.Ref.available(metric):
        pytest.skip("oracle/_ref/libfir_ref_%s.so not built (reference sources absent)" % metric)
    return oracle_py.Ref(metric)


@pytest.fixture(scope="session")
def ref_l2():
    return _ref("l2")


@pytest.fixture(scope="session")
def ref_chi2():
    return _ref("chi2")


@pytest.fixture(scope="session")
def ref_kl():
    return _ref("kl")


@pytest.fixture(scope="session")
def fir():
    import fir_b200
    return fir_b200
