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


def test_new_entry_points_reject_null_arguments(fir):
    """The multi-GPU / DEM / classifier additions validate their arguments before touching a device (status codes, never a crash)."""
    L, C = fir.lib(), ctypes
    BAD = 1
    null = C.c_void_p(None)
    out = C.c_void_p(None)
    assert L.fir_comm_unique_id(null) == BAD
    assert L.fir_comm_init_rank(null, 0, 0, C.byref(out)) == BAD            # world < 1
    assert L.fir_comm_init_rank(null, 3, 2, C.byref(out)) == BAD            # rank outside the world
    assert L.fir_comm_info(null, None, None, None) == BAD
    assert L.fir_comm_destroy(null) == 0                                     # destroying nothing is fine, like free(NULL)
    assert L.fir_shard_search_topk(null, null, null, 1, 1, 0, 0, null, null) == BAD
    assert L.fir_shard_class_min(null, null, null, 1, 0, null, null) == BAD
    assert L.fir_shard_pnn_scores(null, null, null, 1, 1e-3, 10, 0, null, null) == BAD
    assert L.fir_shard_dem_build(null, null, 10, None, C.byref(out)) == BAD
    assert L.fir_sharded_create(null, null, 0, 8, 0, 1, C.byref(out)) == BAD
    assert L.fir_sharded_info(null, None, None, None, None) == BAD
    assert L.fir_sharded_search_topk(null, null, 1, 1, 0, null, null) == BAD
    assert L.fir_sharded_dem_build(null, None, C.byref(out)) == BAD
    assert L.fir_sharded_dem_search(null, null, 1, 0, null, null, null, null) == BAD
    assert L.fir_sharded_destroy(null) == 0 and L.fir_sharded_dem_destroy(null) == 0
    assert L.fir_dem_search_stats(null, None, None, None, None) == BAD
    assert L.fir_classifier_knn_ex(null, null, 1, 3, 0, null) == BAD
    assert L.fir_classifier_pnn_ex(null, null, 1, 0, null, null) == BAD
    assert L.fir_classifier_profile(null, 1, None, None) == BAD
    assert L.fir_gallery_index_offset(null, None) == BAD
    assert L.fir_synth_rows(null, null, 0, 1, 1, 8, 2, 0, 1, 0.5, 0, null) == BAD
    assert b"" != L.fir_last_error_string()


def test_sharded_host_model_matches_c_key_format():
    """sharded.py's key format is the one csrc/sharded.cu packs: ordered distance bits << 32 | global index, ~0 = empty."""
    import importlib
    sh = importlib.import_module("fast-image-recognition_b200.sharded")
    k = sh.pack_keys(np.array([0.0, 1.5, -2.0], np.float32), np.array([7, 2**31 - 1, -1]))
    assert k[0] == (0x80000000 << 32) | 7 and k[1] == ((0x3FC00000 | 0x80000000) << 32) | (2**31 - 1) and k[2] == sh.EMPTY_KEY
    d, i = sh.unpack_keys(k)
    assert i.tolist() == [7, 2**31 - 1, -1] and d[:2].tolist() == [0.0, 1.5]


def test_tensor_work_partition_covers_every_item_once(fir):
    """The candidate kernel's work partition (full rounds, balanced remainder, phased remainder for galleries larger than L2):
    every (query block, tile) item exactly once, distinct candidate slots per query block — checked on the host for the
    BASELINE shapes and a sweep of ragged ones; the phased form is what C5 runs (21 remainder blocks on 74 pairs → 4 + 24 ranges)."""
    L, C = fir.lib(), ctypes
    ph, sl = C.c_int32(0), C.c_int32(0)
    chk = lambda nq, n, sm, ctas, rb: L.fir_debug_partition_check(nq, n, sm, ctas, rb, C.byref(ph), C.byref(sl))
    assert chk(100_000, 2_000_000, 148, 2, 1024) == 0 and ph.value == 2 and sl.value == 24     # C5's query side (gallery cut to keep the check small)
    assert chk(10_000, 100_000, 148, 2, 1024) == 0 and ph.value == 0                              # C2: no full round → balanced cut
    assert chk(100_000, 40_000, 148, 2, 1024) == 0 and ph.value == 0                              # shadow fits L2 → balanced cut
    assert chk(3_000_000 + 300, 40_000, 148, 2, 10**6) == 0 and ph.value >= 1 and sl.value <= 5   # very many queries: few ranges, small candidate arrays
    rng = np.random.default_rng(5)
    seen_phased = 0
    for _ in range(300):
        sm = int(rng.choice([148, 132, 64, 20, 2]))
        ctas = int(rng.choice([1, 2]))
        nq = int(rng.integers(1, 60_000))
        n = int(rng.integers(1, 400_000))
        rb = int(rng.choice([0, 128, 1024, 4096]))
        assert chk(nq, n, sm, ctas, rb) == 0, (nq, n, sm, ctas, rb)
        seen_phased += ph.value > 0
        assert ph.value <= 4 and sl.value <= 160
    assert seen_phased > 10


def test_library_carries_the_blackwell_data_path():
    """The shipped library is sm_100a code with the tcgen05 / TMA / TMEM data path in its candidate kernels (no PTX-only or
    recompiled mma.sync build can pass for it): UTCHMMA.2CTA (tcgen05.mma cta_group::2), UTMALDG (TMA loads), LDTM
    (tcgen05.ld), UTCBAR (tcgen05.commit), FFMA2 in the epilogue.  Needs cuobjdump (CUDA toolkit); skipped without it."""
    import shutil
    import subprocess
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    so = os.path.join(ROOT, "fast-image-recognition_b200", "libfir_b200.so")
    if not os.path.exists(exe) or not os.path.exists(so):
        pytest.skip("cuobjdump or the built library not present")
    lst = subprocess.run([exe, "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in lst and "sm_90" not in lst and "sm_80" not in lst          # one architecture, the target
    sass = subprocess.run([exe, "-sass", "-fun", "_ZN3fir25l2_candidates_kernel_2ctaILi16ELb1ELi0EEEv14CUtensorMap_stS1_NS_10CandParamsE", so],
                          capture_output=True, text=True).stdout
    for op in ("UTCHMMA.2CTA", "UTMALDG.2D.2CTA", "LDTM", "UTCBAR.2CTA.MULTICAST", "FFMA2", "SYNCS"):
        assert op in sass, op
