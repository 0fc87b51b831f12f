"""CPU: libfir_b200.so loads and exports exactly what include/fir_b200.h declares; without a GPU every compute
entry point fails loudly (there is no CPU fallback behind the C-ABI)."""
import ctypes
import os
import re

import numpy as np
import pytest

from util import has_cuda

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "fir_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fir_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(fir):
    names = declared_functions()
    assert len(names) >= 28
    L = fir.lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(fir.EXPORTS) == names           # the Python binding tracks the header one to one
    assert L.fir_version() == 100


def test_product_never_touches_the_oracle():
    """The package must not import, load or link anything under oracle/."""
    pkg = os.path.join(ROOT, "fast-image-recognition_b200")
    for dp, _, files in os.walk(pkg):
        if os.path.basename(dp) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".sh")):
                text = open(os.path.join(dp, f)).read()
                assert "oracle_py" not in text and "libfir_oracle" not in text and "libfir_ref" not in text, f


@pytest.mark.skipif(has_cuda(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_a_gpu(fir):
    with pytest.raises(fir.FirError) as e:
        fir.Gallery(np.ones((4, 8), np.float32))
    assert e.value.code in (2, 3)                 # FIR_ERR_CUDA: no device, and no fallback


def test_missing_library_fails_loudly(fir, monkeypatch, tmp_path):
    monkeypatch.setattr(fir, "_lib", None)
    monkeypatch.setattr(fir, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(fir.FirError):
        fir.lib()
