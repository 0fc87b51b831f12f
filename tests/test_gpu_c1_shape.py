"""GPU: BASELINE config 1 at the reference's own shape through the C++ drop-in adapters.

A Caltech-101-shaped text features file (101 classes, 8677 images, D = 1536, `{:f}` formatting —
qt_cpp/dnn_feature_extractor.py:58-64) goes through loadImages → getTrainingAndTestImages(randomize = true) under srand(13) →
BruteForce → DirectedEnumeration in a C++ program compiled against include/fir_b200_compat.hpp, and every output is compared
with the UNMODIFIED reference (oracle/_ref: loadImages, getTrainingAndTestImages, BruteForce, the verbatim DirectedEnumeration
constructor and recognize, feature_distance with a window) run on the same file with the same seeds."""
import importlib
import os
import subprocess

import numpy as np
import pytest

from util import bits

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
synth = importlib.import_module("fast-image-recognition_b200.synth")
D, SPLIT_SEED, DEM_SEED = 1536, 13, 5


def _write_caltech_file(path):
    sizes = synth.caltech_sizes()
    labels = np.repeat(np.arange(101), sizes)
    cen = synth._z(0xC1 + 3, np.arange(101), D)
    rows = (cen[labels] + np.float32(4.0) * synth._z(0xC1, np.arange(len(labels)), D)).astype(np.float32)    # sigma 4: ~11 % 1-NN errors
    with open(path, "w") as f:
        f.write("bg_0.jpg\nBACKGROUND_Google\n" + "0.500000 " * D + "\n")       # filtered by the Caltech switch (db_features.cpp:59-62)
        for i, row in enumerate(rows):
            f.write("img_%05d.jpg\nclass_%03d\n" % (i, labels[i]))
            f.write(" ".join(map("{:f}".format, row.tolist())) + " \n")
    return len(rows)


def test_c1_shape_adapters_match_reference(ref_l2, tmp_path):
    txt = str(tmp_path / "caltech_shape.txt")
    total = _write_caltech_file(txt)
    exe = str(tmp_path / "c1_test")
    pkg = os.path.join(ROOT, "fast-image-recognition_b200")
    subprocess.run(["g++", "-std=c++11", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "c1_test.cpp"),
                    "-o", exe, "-L", pkg, "-lfir_b200", "-Wl,-rpath," + pkg], check=True)
    out = subprocess.run([exe, txt, str(D), str(SPLIT_SEED), str(DEM_SEED)], check=True, capture_output=True, text=True).stdout
    lines = {l.split()[0]: l.split()[1:] for l in out.splitlines() if l and l.split()[0].isupper()}
    ints = lambda key: np.array([int(x) for x in lines[key]], np.int64)

    db, dbl, dbi, te, tel, tei = ref_l2.load_split(txt, D, seed=SPLIT_SEED, randomize=True)
    assert int(lines["LOADED"][0]) == total == 8677 and int(lines["LOADED"][2]) == 101
    assert len(db) == 3030 and len(te) == 4779                                  # 30 per class; the four classes above 400 are capped
    assert np.array_equal(ints("DBIDX"), dbi) and np.array_equal(ints("TESTIDX"), tei)         # the randomised split itself
    bi, bd = ref_l2.bf(db, te, dbl, nthreads=os.cpu_count() or 1)
    assert np.array_equal(ints("BF"), bi)                                       # per-query neighbour indices
    err = 100.0 * float(np.mean((bi < 0) | (dbl[np.maximum(bi, 0)] != tel)))
    assert abs(float(lines["BFERR"][0]) - err) < 1e-6                           # error % of testSetRecognition (ann.cpp:99-103)
    assert bits(np.float32(float(lines["WINDOW"][0]))) == bits(ref_l2.distance(te[0], db[1], 64, 320))
    assert bits(np.float32(float(lines["WINDOW"][1]))) == bits(ref_l2.distance(te[1], db[2], 100, 101))
    dem = ref_l2.dem_create(db, dbl, seed=DEM_SEED)                             # the verbatim constructor under the same srand
    assert bits(np.float32(float(lines["THRESHOLD"][0]))) == bits(dem.threshold)
    for r, ratio in enumerate((0.05, 0.2)):
        di = dem.search(te, int(ratio * len(db)), nthreads=os.cpu_count() or 1)[0]
        assert np.array_equal(ints("DEM%d" % r), di)
        derr = 100.0 * float(np.mean((di < 0) | (dbl[np.maximum(di, 0)] != tel)))
        assert abs(float(lines["DEMERR%d" % r][0]) - derr) < 1e-6
    dem.close()
    # the same program with fir::n_gpus() = all devices: BruteForce over row shards + NCCL merge inside ONE process
    import torch
    if torch.cuda.device_count() >= 2:
        out2 = subprocess.run([exe, txt, str(D), str(SPLIT_SEED), str(DEM_SEED), "0"], check=True, capture_output=True, text=True).stdout
        lines2 = {l.split()[0]: l.split()[1:] for l in out2.splitlines() if l and l.split()[0].isupper()}
        assert lines2["BF"] == lines["BF"] and lines2["BFERR"] == lines["BFERR"]
        assert lines2["THRESHOLD"] == lines["THRESHOLD"] and lines2["DEM0"] == lines["DEM0"] and lines2["DEM1"] == lines["DEM1"]   # DEM over the shards
