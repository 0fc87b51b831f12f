"""GPU: the C++ drop-in adapters (include/fir_b200_compat.hpp) compiled with the reference's own language level
(-std=c++11, recognition_testing.pro:38) against libfir_b200.so, driven like testANN, checked against the oracle."""
import importlib
import os
import subprocess

import numpy as np
import pytest

from util import bits

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
synth = importlib.import_module("fast-image-recognition_b200.synth")


def test_cpp_adapters_match_oracle(port, tmp_path):
    d, n_classes = 48, 9
    g, gl, q, ql = synth.make_split(9 * 45, 1, d, n_classes, "l2", seed=21)    # USE_CALTECH split: 30 per class to the gallery, rest to test
    txt = str(tmp_path / "features.txt")
    synth.write_features_file(txt, g, ["class_%02d" % c for c in gl])
    exe = str(tmp_path / "compat_test")
    pkg = os.path.join(ROOT, "fast-image-recognition_b200")
    subprocess.run(["g++", "-std=c++11", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "compat_test.cpp"),
                    "-o", exe, "-L", pkg, "-lfir_b200", "-Wl,-rpath," + pkg], check=True)
    # expected: parse the same text, loader normalisation, class-major split without shuffling
    parsed = np.array([[np.float32(float("{:f}".format(float(v)))) for v in row] for row in g], np.float32)
    rows = port.normalize_rows("l2", parsed)
    db_i, te_i = [], []
    for c in range(n_classes):
        idx = np.flatnonzero(gl == c)
        db_i += list(idx[:30])
        te_i += list(idx[30:400])
    db, test = rows[db_i], rows[te_i]
    dbl, tel = gl[db_i], gl[te_i]
    pivot0 = 7
    out = subprocess.run([exe, txt, str(d), str(pivot0)], check=True, capture_output=True, text=True).stdout
    lines = {l.split()[0]: l.split()[1:] for l in out.splitlines() if l and l.split()[0].isupper()}
    assert lines["LOADED"][0] == str(len(g)) and int(lines["LOADED"][4]) == len(db) and int(lines["LOADED"][6]) == len(test)
    bi, bd = port.bf("l2", db, test)
    assert [int(x) for x in lines["BF"]] == bi.tolist()
    assert int(lines["BF1"][0]) == bi[0] and int(lines["BF1"][2]) == bi[0]
    assert int(lines["BF1"][4]) == port.bf("l2", db, test[:1], max_features=d // 2)[0][0]
    assert bits(np.float32(float(lines["DIST"][0]))) == bits(port.distance("l2", test[0], db[3]))
    b = port.dem_build("l2", db, dbl, pivot0)
    assert bits(np.float32(float(lines["THRESHOLD"][0]))) == bits(np.float32(b["threshold"]))
    di = port.dem_search("l2", db, b["pivots"][: b["n_pivots"]], b["P"], b["threshold"], int(0.2 * len(db)), test)[0]
    assert [int(x) for x in lines["DEM"]] == di.tolist()
    tr64, te64 = db.astype(np.float64), test.astype(np.float64)
    avg = np.array([np.add.reduce(tr64[:, f]) for f in range(d)])          # sequential left-to-right like the adapter
    avg = np.array([sum(tr64[:, f].tolist()) for f in range(d)]) / len(tr64)
    assert [int(x) for x in lines["KNN3"]] == port.knn(tr64, dbl, n_classes, avg, te64, 3).tolist()
    sc, pl = port.pnn(tr64, dbl, n_classes, avg, te64)
    assert [int(x) for x in lines["PNN"]] == pl.tolist() and int(lines["PNN1"][0]) == pl[0]
    assert [int(x) for x in lines["PNNSEQ"]] == port.pnn_seq(tr64, dbl, n_classes, avg, te64).tolist()
    sq = np.array([sum((tr64[:, f] * tr64[:, f]).tolist()) for f in range(d)])
    sd = np.sqrt((sq - avg * avg * len(tr64)) / (len(tr64) - 1))                      # stdValues, classification.cpp:988
    a, J = port.fpnn_train(tr64, dbl, n_classes, avg, sd, 1.0)
    assert [int(x) for x in lines["FPNN"]] == port.fpnn_predict(a, J, n_classes, avg, sd, te64, 1.0).tolist()
    a, J = port.fpnn_train(tr64, dbl, n_classes, avg, sd, 0.33)
    assert [int(x) for x in lines["FPNNSEQ"]] == port.fpnn_predict(a, J, n_classes, avg, sd, te64, 0.33, sequential=True, output_ratio=0.9).tolist()
    sel = port.kmedoids(tr64, dbl, n_classes, 4)
    assert [int(x) for x in lines["PNNCLUST"]] == port.pnn(tr64[sel], dbl[sel], n_classes, avg, te64)[1].tolist()
    assert " ".join(lines["NAMES"]) == "FPNN, 1|FPNN, 0.33 (seq)|PNN with clustering, 4"


def test_cpp_twd_adapters_match_oracle(port, tmp_path):
    """fir_compat::image_testing (ImageTesting.cpp's Classifier family) driven like testRecognition, vs the oracle port."""
    d, n_classes = 256, 8
    g, gl, q, ql = synth.make_split(8 * 50, 1, d, n_classes, "l2", sigma=2.0, seed=5)
    txt = str(tmp_path / "features.txt")
    synth.write_features_file(txt, g, ["class_%02d" % c for c in gl])
    exe = str(tmp_path / "twd_test")
    pkg = os.path.join(ROOT, "fast-image-recognition_b200")
    subprocess.run(["g++", "-std=c++11", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "twd_test.cpp"),
                    "-o", exe, "-L", pkg, "-lfir_b200", "-Wl,-rpath," + pkg], check=True)
    parsed = np.array([[np.float32(float("{:f}".format(float(v)))) for v in row] for row in g], np.float32)
    rows = port.normalize_rows("l2", parsed)
    db_i, te_i = [], []
    for c in range(n_classes):
        idx = np.flatnonzero(gl == c)
        db_i += list(idx[:30])
        te_i += list(idx[30:400])
    db, test, dbl = rows[db_i], rows[te_i], gl[db_i]
    out = subprocess.run([exe, txt, str(d)], check=True, capture_output=True, text=True).stdout
    lines = {l.split()[0]: l.split()[1:] for l in out.splitlines() if l}
    assert [int(v) for v in lines["SIZES"]] == [n_classes, len(db), len(test)]
    want = []
    for mf in (0, 64, 0):                                                  # BF, BF 64, BF 256 (= all 256 dimensions)
        want.append((dbl[port.bf("l2", db, test, max_features=mf)[0]], np.zeros(len(test), np.uint8)))
    for kind, th in (("posteriors", 0.24), ("diff", 0.003), ("ratio", 0.7)):
        want.append(port.twd_conventional("l2", db, dbl, n_classes, test, kind, th, 64)[1:])
    for fc in (32, 64):
        want.append(port.twd_proposed("l2", db, dbl, test, fc, 0.7)[1:])
    names = ["BF, 256", "BF, 64", "BF, 256", "TWD posteriors, 0.24", "TWD diff, 0.003", "TWD ratio, 0.7", "Proposed TWD, 32, 0.7", "Proposed TWD, 64, 0.7"]
    some_unreliable = 0
    for c, (cls, unrel) in enumerate(want):
        assert " ".join(lines["NAME%d" % c]) == names[c]
        assert [int(v) for v in lines["CLS%d" % c]] == cls.tolist(), names[c]
        assert int(lines["UNREL%d" % c][0]) == int(unrel.sum()), names[c]
        assert [int(v) for v in lines["ONE%d" % c]] == [int(cls[1]), int(unrel[1])], names[c]
        some_unreliable += int(unrel.sum())
    assert some_unreliable > 0


def _ytf_files(tmp_path, d, n_people, sigma, seed):
    """Still images of people 0..n-2 (plus nobody from the last one), videos of people 1..n-1: two videos each."""
    from oracle import oracle_py
    g, gl, q, ql = synth.make_split(n_people * 12, n_people * 60, d, n_people, "l2", sigma=sigma, seed=seed)
    names = ["person_%02d" % c for c in range(n_people)]
    keep = np.flatnonzero(gl < n_people - 1)
    synth.write_features_file(str(tmp_path / oracle_py.Ref.TRAIN_FILE), g[keep], [names[c] for c in gl[keep]])
    people = {}
    for c in range(1, n_people):
        rows = q[ql == c]
        cut = len(rows) // 2 - 3
        people[names[c]] = [rows[:cut], rows[cut:]]
    synth.write_video_file(str(tmp_path / oracle_py.Ref.VIDEO_FILE), people)
    return g[keep], gl[keep], people, names


def test_cpp_video_adapters_match_reference(port, ref_l2, tmp_path):
    """fir_compat::loadVideos / buildYTFSplit + BruteForce / DirectedEnumeration, driven like testYTFRecognition (video.cpp),
    against the verbatim testYTFRecognition() run by oracle/_ref on the same two files."""
    d, n_people = 64, 9
    stills, still_labels, people, names = _ytf_files(tmp_path, d, n_people, sigma=3.0, seed=8)
    exe = str(tmp_path / "ytf_test")
    pkg = os.path.join(ROOT, "fast-image-recognition_b200")
    subprocess.run(["g++", "-std=c++11", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "ytf_test.cpp"),
                    "-o", exe, "-L", pkg, "-lfir_b200", "-Wl,-rpath," + pkg], check=True)
    out = subprocess.run([exe, str(tmp_path), str(d), "3"], check=True, capture_output=True, text=True).stdout.splitlines()
    want = ref_l2.ytf_run(str(tmp_path), d).splitlines()
    for prefix in ("total size=", "lfw names size=", "dbSize="):           # loader + split bookkeeping lines, verbatim
        assert [l for l in out if l.startswith(prefix)] == [l for l in want if l.startswith(prefix)], prefix
    bf_err = [l.split("%")[0] for l in out if l.startswith("BF error=")]
    assert bf_err == [l.split("%")[0] for l in want if l.startswith("BF error=")] and float(bf_err[0].split("=")[1]) > 0
    # every frame the reference loads, bit for bit: the queries are every 10th frame of every video in name order
    rnames, frames, person, video, frame = ref_l2.videos_load(str(tmp_path), d)
    common = [n for n in rnames if n in names[:n_people - 1]]
    sel = np.array([i for i in range(len(frames)) if rnames[person[i]] in common and frame[i] % 10 == 0])
    test = frames[sel]
    lines = {l.split()[0]: l.split()[1:] for l in out if l and l.split()[0].isupper()}
    assert int(lines["TESTBITS"][0]) == int(test.view(np.uint32).astype(np.uint64).sum())
    test_class = np.array([common.index(rnames[person[i]]) for i in sel])
    assert [int(x) for x in lines["TESTCLASS"]] == test_class.tolist()
    # gallery: the stills of the common people through the loader restatement; classes are positions in `common`
    parsed = np.array([[np.float32(float("{:f}".format(float(v)))) for v in row] for row in stills], np.float32)
    rows = port.normalize_rows("l2", parsed)
    gsel = np.array([i for i in range(len(rows)) if names[still_labels[i]] in common])
    gallery, gclass = rows[gsel], np.array([common.index(names[still_labels[i]]) for i in gsel])
    bi, bd = port.bf("l2", gallery, test)
    assert [int(x) for x in lines["BFCLASS"]] == gclass[bi].tolist()
    assert abs(100.0 * np.mean(gclass[bi] != test_class) - float(bf_err[0].split("=")[1])) < 1e-3
    dem_err = [float(l.split("%")[0].split("=")[1]) for l in out if l.startswith("dem error=")]
    assert len(dem_err) == 7 and dem_err[-1] <= dem_err[0] + 1e-9          # more candidates never hurt


def test_video_loader_normalisation_variants(fir, ref_l2, ref_chi2, tmp_path):
    """loadVideos' normalisation (video.cpp:64-84) on the GPU: the L2 build divides by sqrt(sum x^2), the other builds by
    sum x^2 (FIR_NORM_VIDEO_SUMSQ) — both against the frames the verbatim loader produces."""
    d = 48
    stills, still_labels, people, names = _ytf_files(tmp_path, d, 5, sigma=1.0, seed=2)
    raw = np.concatenate([np.concatenate(v) for k, v in sorted(people.items())])
    parsed = np.array([[np.float32(float("{:f}".format(float(v)))) for v in row] for row in raw], np.float32)
    for ref, mode in ((ref_l2, "l2"), (ref_chi2, 3)):
        rnames, frames, person, video, frame = ref.videos_load(str(tmp_path), d)
        got = fir.normalize_rows(parsed.copy(), mode)
        assert np.array_equal(bits(got), bits(frames))
