"""TEST INFRASTRUCTURE ONLY — ctypes loaders for the two CPU oracles.

* ``Port``  : oracle/libfir_oracle.so, our plain-C restatement (fir_oracle.c).
* ``Ref``   : oracle/_ref/libfir_ref_{l2,chi2,kl}.so, the UNMODIFIED reference translation units
              (/root/reference/qt_cpp/{db_features,ann,classification}.cpp) behind ref_capi.cpp.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) import this.
Nothing here reads /root/reference at run time; building the libraries is oracle/Makefile's job.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
METRICS = {"l2": 0, "chi2": 1, "kl": 2}

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build(verbose=False):
    """(Re)build the oracle libraries. The _ref part is skipped by the Makefile when the reference
    sources are absent (GPU box): the prebuilt oracle/_ref/*.so that travelled with the snapshot is used."""
    r = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Port:
    """Plain-C restatement (oracle/fir_oracle.c)."""

    def __init__(self):
        path = os.path.join(HERE, "libfir_oracle.so")
        if not os.path.exists(path):
            build()
        L = self.L = C.CDLL(path)
        L.fir_oracle_logf.restype = C.c_float
        L.fir_oracle_logf.argtypes = [C.c_float, C.c_int]
        L.fir_oracle_logf_host_variant.restype = C.c_int
        L.fir_oracle_distance.restype = C.c_float
        L.fir_oracle_distance.argtypes = [C.c_int, _f32p, _f32p, C.c_int, C.c_int]
        L.fir_oracle_normalize_rows.argtypes = [C.c_int, _f32p, C.c_int64, C.c_int]
        L.fir_oracle_bf.restype = C.c_double
        L.fir_oracle_bf.argtypes = [C.c_int, _f32p, C.c_int64, C.c_int, _f32p, C.c_int64, C.c_int, C.c_int, _i32p, _f32p]
        L.fir_oracle_topk.restype = C.c_double
        L.fir_oracle_topk.argtypes = [C.c_int, _f32p, C.c_int64, C.c_int, _f32p, C.c_int64, C.c_int, C.c_int, _i32p, _f32p]
        L.fir_oracle_class_min.argtypes = [C.c_int, _f32p, _i32p, C.c_int64, C.c_int, C.c_int, _f32p, C.c_int64, _f32p, _i32p]
        L.fir_oracle_pnn_div.argtypes = [C.c_int, _f32p, _i32p, C.c_int64, C.c_int, C.c_int, _f32p, C.c_int64, C.c_double, _f64p, _i32p]
        L.fir_oracle_knn.argtypes = [_f64p, _i32p, C.c_int64, C.c_int, C.c_int, _f64p, _f64p, C.c_int64, C.c_int, _i32p]
        L.fir_oracle_pnn.argtypes = [_f64p, _i32p, C.c_int64, C.c_int, C.c_int, _f64p, _f64p, C.c_int64, _f64p, _i32p]
        L.fir_oracle_pnn_seq.argtypes = [_f64p, _i32p, C.c_int64, C.c_int, C.c_int, _f64p, _f64p, C.c_int64, _i32p]
        L.fir_oracle_fasterlog2.restype = C.c_float
        L.fir_oracle_fasterlog2.argtypes = [C.c_float]
        L.fir_oracle_fpnn_J.argtypes = [C.c_int64, C.c_int]
        L.fir_oracle_fpnn_train.argtypes = [_f64p, _i32p, C.c_int64, C.c_int, C.c_int, _f64p, _f64p, C.c_double, C.c_int, _f64p]
        L.fir_oracle_fpnn_predict.argtypes = [_f64p, C.c_int, C.c_int, C.c_int, _f64p, _f64p, C.c_double, _f64p, C.c_int64, C.c_int, C.c_float, _i32p]
        L.fir_oracle_kmedoids.restype = C.c_int64
        L.fir_oracle_kmedoids.argtypes = [_f64p, _i32p, C.c_int64, C.c_int, C.c_int, C.c_int, _i64p]
        L.fir_oracle_twd_conventional.argtypes = [C.c_int, _f32p, _i32p, C.c_int64, C.c_int, C.c_int, _f32p, C.c_int64, C.c_int, C.c_double,
                                                  C.c_int, C.c_int, _i32p, _i32p, _u8p]
        L.fir_oracle_twd_proposed.argtypes = [C.c_int, _f32p, _i32p, C.c_int64, C.c_int, _f32p, C.c_int64, C.c_int, C.c_double, C.c_int,
                                              _i32p, _i32p, _u8p]
        L.fir_oracle_dem_build.restype = C.c_int
        L.fir_oracle_dem_build.argtypes = [C.c_int, _f32p, _i32p, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int,
                                           C.POINTER(C.c_int), _i32p, _f32p, _f32p, C.POINTER(C.c_float)]
        L.fir_oracle_dem_search.argtypes = [C.c_int, _f32p, C.c_int64, C.c_int, _i32p, C.c_int, _f32p, C.c_float, C.c_int,
                                            _f32p, C.c_int64, _i32p, _f32p, _u8p, _i32p]

    def logf(self, x, fma=1):
        return self.L.fir_oracle_logf(float(np.float32(x)), fma)

    def logf_host_variant(self):
        return self.L.fir_oracle_logf_host_variant()

    def distance(self, metric, l, r, start=0, end=None):
        l, r = _f32(l), _f32(r)
        return np.float32(self.L.fir_oracle_distance(METRICS[metric], l, r, start, len(l) if end is None else end))

    def normalize_rows(self, metric, rows):
        rows = _f32(rows).copy()
        self.L.fir_oracle_normalize_rows(METRICS[metric], rows, rows.shape[0], rows.shape[1])
        return rows

    def bf(self, metric, g, q, max_features=0, nthreads=1, timing=False):
        g, q = _f32(g), _f32(q)
        idx = np.empty(q.shape[0], np.int32)
        dist = np.empty(q.shape[0], np.float32)
        t = self.L.fir_oracle_bf(METRICS[metric], g, g.shape[0], g.shape[1], q, q.shape[0], max_features, nthreads, idx, dist)
        return (idx, dist, t) if timing else (idx, dist)

    def topk(self, metric, g, q, k, nthreads=1, timing=False):
        g, q = _f32(g), _f32(q)
        idx = np.empty((q.shape[0], k), np.int32)
        dist = np.empty((q.shape[0], k), np.float32)
        t = self.L.fir_oracle_topk(METRICS[metric], g, g.shape[0], g.shape[1], q, q.shape[0], k, nthreads, idx, dist)
        return (idx, dist, t) if timing else (idx, dist)

    def class_min(self, metric, g, labels, n_classes, q):
        g, q, labels = _f32(g), _f32(q), _i32(labels)
        mn = np.empty((q.shape[0], n_classes), np.float32)
        arg = np.empty((q.shape[0], n_classes), np.int32)
        self.L.fir_oracle_class_min(METRICS[metric], g, labels, g.shape[0], g.shape[1], n_classes, q, q.shape[0], mn, arg)
        return mn, arg

    def pnn_div(self, metric, g, labels, n_classes, q, var):
        g, q, labels = _f32(g), _f32(q), _i32(labels)
        sc = np.empty((q.shape[0], n_classes), np.float64)
        lab = np.empty(q.shape[0], np.int32)
        self.L.fir_oracle_pnn_div(METRICS[metric], g, labels, g.shape[0], g.shape[1], n_classes, q, q.shape[0], float(var), sc, lab)
        return sc, lab

    def knn(self, train, train_label, n_classes, avg, q, K):
        train, q, avg, train_label = _f64(train), _f64(q), _f64(avg), _i32(train_label)
        lab = np.empty(q.shape[0], np.int32)
        self.L.fir_oracle_knn(train, train_label, train.shape[0], train.shape[1], n_classes, avg, q, q.shape[0], K, lab)
        return lab

    def pnn(self, train, train_label, n_classes, avg, q):
        train, q, avg, train_label = _f64(train), _f64(q), _f64(avg), _i32(train_label)
        sc = np.empty((q.shape[0], n_classes), np.float64)
        lab = np.empty(q.shape[0], np.int32)
        self.L.fir_oracle_pnn(train, train_label, train.shape[0], train.shape[1], n_classes, avg, q, q.shape[0], sc, lab)
        return sc, lab

    def pnn_seq(self, train, train_label, n_classes, avg, q):
        train, q, avg, train_label = _f64(train), _f64(q), _f64(avg), _i32(train_label)
        lab = np.empty(q.shape[0], np.int32)
        self.L.fir_oracle_pnn_seq(train, train_label, train.shape[0], train.shape[1], n_classes, avg, q, q.shape[0], lab)
        return lab

    def fpnn_train(self, train, train_label, n_classes, avg, sd, scale=1.0):
        train, avg, sd, train_label = _f64(train), _f64(avg), _f64(sd), _i32(train_label)
        J = self.L.fir_oracle_fpnn_J(train.shape[0], n_classes)
        a = np.zeros(train.shape[1] * n_classes * (2 * J + 1), np.float64)
        self.L.fir_oracle_fpnn_train(train, train_label, train.shape[0], train.shape[1], n_classes, avg, sd, float(scale), J, a)
        return a, J

    def fpnn_predict(self, a, J, n_classes, avg, sd, q, scale=1.0, sequential=False, output_ratio=0.9):
        q, avg, sd = _f64(q), _f64(avg), _f64(sd)
        lab = np.empty(q.shape[0], np.int32)
        self.L.fir_oracle_fpnn_predict(_f64(a), J, q.shape[1], n_classes, avg, sd, float(scale), q, q.shape[0], int(sequential), float(output_ratio), lab)
        return lab

    def kmedoids(self, train, train_label, n_classes, num_clusters):
        train, train_label = _f64(train), _i32(train_label)
        sel = np.empty(train.shape[0], np.int64)
        cnt = self.L.fir_oracle_kmedoids(train, train_label, train.shape[0], train.shape[1], n_classes, int(num_clusters), sel)
        if cnt < 0:
            raise ValueError("a cluster ran empty (undefined behaviour in the reference)")
        return sel[:cnt].copy()

    TWD_TYPES = {"posteriors": 0, "diff": 1, "ratio": 2}

    def twd_conventional(self, metric, g, labels, n_classes, q, kind, threshold, feat_count=64, last_feature=256):
        """ConventionalTWDClassifier::recognize → (index, class, unreliable)."""
        g, q, labels = _f32(g), _f32(q), _i32(labels)
        idx, cls, unrel = np.empty(len(q), np.int32), np.empty(len(q), np.int32), np.empty(len(q), np.uint8)
        self.L.fir_oracle_twd_conventional(METRICS[metric], g, labels, g.shape[0], g.shape[1], n_classes, q, q.shape[0],
                                           self.TWD_TYPES[kind], float(threshold), feat_count, last_feature, idx, cls, unrel)
        return idx, cls, unrel

    def twd_proposed(self, metric, g, labels, q, feat_count, th, last_feature=256):
        """ProposedTWDClassifier::recognize → (index, class, unreliable)."""
        g, q, labels = _f32(g), _f32(q), _i32(labels)
        idx, cls, unrel = np.empty(len(q), np.int32), np.empty(len(q), np.int32), np.empty(len(q), np.uint8)
        self.L.fir_oracle_twd_proposed(METRICS[metric], g, labels, g.shape[0], g.shape[1], q, q.shape[0], feat_count, float(th),
                                       last_feature, idx, cls, unrel)
        return idx, cls, unrel

    def dem_build(self, metric, g, labels, pivot0, far=0.01, threshold=0.0, keep_rows=32):
        g, labels = _f32(g), _i32(labels)
        n = g.shape[0]
        np_total = max(5, int(n * 0.015))
        pivots = np.zeros(np_total, np.int32)
        keep = min(keep_rows, np_total)
        P = np.zeros((keep, n), np.float32)
        min_other = np.zeros(np_total, np.float32)
        thr = C.c_float(0)
        npo = C.c_int(0)
        s = self.L.fir_oracle_dem_build(METRICS[metric], g, labels, n, g.shape[1], int(pivot0), far, threshold, keep,
                                        C.byref(npo), pivots, P, min_other, C.byref(thr))
        return dict(n_pivots=s, np_total=npo.value, pivots=pivots, P=P[:s].copy(), min_other=min_other, threshold=np.float32(thr.value))

    def dem_search(self, metric, g, pivots, P, threshold, count_to_check, q):
        g, q, P, pivots = _f32(g), _f32(q), _f32(P), _i32(pivots)
        nq = q.shape[0]
        idx = np.empty(nq, np.int32)
        dist = np.empty(nq, np.float32)
        below = np.empty(nq, np.uint8)
        evals = np.empty(nq, np.int32)
        self.L.fir_oracle_dem_search(METRICS[metric], g, g.shape[0], g.shape[1], pivots, len(pivots), P, float(threshold),
                                     int(count_to_check), q, nq, idx, dist, below, evals)
        return idx, dist, below, evals


class Ref:
    """The unmodified reference code (oracle/_ref/libfir_ref_<metric>.so)."""

    @staticmethod
    def available(metric="l2"):
        return os.path.exists(os.path.join(HERE, "_ref", "libfir_ref_%s.so" % metric))

    def __init__(self, metric="l2"):
        self.metric = metric
        path = os.path.join(HERE, "_ref", "libfir_ref_%s.so" % metric)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        L = self.L = C.CDLL(path)
        assert L.fir_ref_metric() == METRICS[metric]
        L.fir_ref_feature_distance.restype = C.c_float
        L.fir_ref_feature_distance.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int]
        L.fir_ref_bf_search.restype = C.c_double
        L.fir_ref_bf_search.argtypes = [_f32p, _i32p, C.c_long, C.c_int, _f32p, C.c_long, C.c_int, C.c_int, _i32p, _f32p]
        L.fir_ref_all_distances.argtypes = [_f32p, C.c_long, C.c_int, _f32p, C.c_long, C.c_int, _f32p]
        L.fir_ref_dataset_load.restype = C.c_void_p
        L.fir_ref_dataset_load.argtypes = [C.c_char_p, C.c_int, C.c_uint, C.c_int]
        L.fir_ref_dataset_count.restype = C.c_long
        L.fir_ref_dataset_count.argtypes = [C.c_void_p, C.c_int]
        L.fir_ref_dataset_get.argtypes = [C.c_void_p, C.c_int, _f32p, _i32p, _i32p]
        L.fir_ref_dataset_free.argtypes = [C.c_void_p]
        L.fir_ref_dem_create.restype = C.c_void_p
        L.fir_ref_dem_create.argtypes = [_f32p, _i32p, C.c_long, C.c_int, C.c_uint, C.c_float, C.c_float, C.c_int, C.POINTER(C.c_double)]
        L.fir_ref_dem_create_injected.restype = C.c_void_p
        L.fir_ref_dem_create_injected.argtypes = [_f32p, _i32p, C.c_long, C.c_int, _i32p, C.c_int, _f32p, C.c_float]
        L.fir_ref_dem_num_pivots.argtypes = [C.c_void_p]
        L.fir_ref_dem_threshold.restype = C.c_float
        L.fir_ref_dem_threshold.argtypes = [C.c_void_p]
        L.fir_ref_dem_get_pivots.argtypes = [C.c_void_p, _i32p]
        L.fir_ref_dem_get_P.argtypes = [C.c_void_p, C.c_int, _f32p]
        L.fir_ref_dem_search.restype = C.c_double
        L.fir_ref_dem_search.argtypes = [C.c_void_p, _f32p, C.c_long, C.c_int, C.c_int, _i32p, _f32p, _u8p, _i32p]
        L.fir_ref_dem_free.argtypes = [C.c_void_p]
        L.fir_ref_videos_load.restype = C.c_void_p
        L.fir_ref_videos_load.argtypes = [C.c_char_p, C.c_int]
        L.fir_ref_videos_counts.argtypes = [C.c_void_p, C.POINTER(C.c_long), C.POINTER(C.c_long), C.POINTER(C.c_long)]
        L.fir_ref_videos_get.argtypes = [C.c_void_p, _f32p, _i32p, _i32p, _i32p, C.c_char_p, C.c_long]
        L.fir_ref_videos_free.argtypes = [C.c_void_p]
        L.fir_ref_ytf_run.restype = C.c_long
        L.fir_ref_ytf_run.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_long]
        L.fir_ref_twd_run.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _f32p, _i32p, C.c_int64, C.c_int, _f32p, C.c_int64, _i32p, _u8p]
        if metric == "l2":
            L.fir_ref_cls_setup.argtypes = [_f64p, _i32p, C.c_long, C.c_int, C.c_int, C.c_double, C.c_uint]
            L.fir_ref_cls_counts.restype = C.c_long
            L.fir_ref_cls_counts.argtypes = [C.c_int]
            L.fir_ref_cls_get_split.argtypes = [_i64p, _i32p, _i64p, _f64p]
            L.fir_ref_cls_knn.restype = C.c_double
            L.fir_ref_cls_knn.argtypes = [C.c_int, C.c_long, C.c_long, C.c_int, _i32p]
            L.fir_ref_cls_pnn.restype = C.c_double
            L.fir_ref_cls_pnn.argtypes = [C.c_long, C.c_long, _i32p, C.c_void_p]
            L.fir_ref_cls_get_std.argtypes = [_f64p]
            L.fir_ref_cls_fpnn.argtypes = [C.c_double, C.c_int, C.c_float, C.c_long, C.c_long, _i32p, C.c_void_p, C.POINTER(C.c_int)]
            L.fir_ref_cls_pnn_clustered.argtypes = [C.c_int, C.c_long, C.c_long, _i32p, _i64p]
            L.fir_ref_cls_pnn_seq.restype = C.c_double
            L.fir_ref_cls_pnn_seq.argtypes = [C.c_long, C.c_long, _i32p]

    VIDEO_FILE, TRAIN_FILE = "vgg_mean_dnn_features.txt", "dnn_vgg_features_all_mean.txt"      # video.cpp:30-33

    def videos_load(self, directory, d):
        """The verbatim loadVideos() on <directory>/VIDEO_FILE → (names, frames [F,d], person, video, frame)."""
        h = self.L.fir_ref_videos_load(os.fsencode(directory), d)
        assert h
        a, b, c = C.c_long(0), C.c_long(0), C.c_long(0)
        self.L.fir_ref_videos_counts(h, C.byref(a), C.byref(b), C.byref(c))
        frames = np.empty((c.value, d), np.float32)
        person, video, frame = (np.empty(c.value, np.int32) for _ in range(3))
        names = C.create_string_buffer(1 << 20)
        self.L.fir_ref_videos_get(h, frames, person, video, frame, names, len(names))
        self.L.fir_ref_videos_free(h)
        return names.value.decode().split("\n")[:-1], frames, person, video, frame

    def ytf_run(self, directory, d):
        """The verbatim testYTFRecognition() inside <directory>; returns what it printed."""
        buf = C.create_string_buffer(1 << 20)
        n = self.L.fir_ref_ytf_run(os.fsencode(directory), d, buf, len(buf))
        assert n >= 0
        return buf.value.decode()

    def twd(self, kind, g, labels, n_classes, q, feat_count, th, twd_type="diff"):
        """kind 'conventional' | 'proposed' | 'bf' through the verbatim ImageTesting.cpp classes → (class, unreliable)."""
        g, q, labels = _f32(g), _f32(q), _i32(labels)
        cls, unrel = np.empty(len(q), np.int32), np.empty(len(q), np.uint8)
        k = {"conventional": 0, "proposed": 1, "bf": 2}[kind]
        rc = self.L.fir_ref_twd_run(k, Port.TWD_TYPES[twd_type], float(th), feat_count, n_classes, g, labels, g.shape[0], g.shape[1],
                                    q, q.shape[0], cls, unrel)
        assert rc == 0
        return cls, unrel

    def distance(self, l, r, start=0, end=None):
        l, r = _f32(l), _f32(r)
        return np.float32(self.L.fir_ref_feature_distance(l, r, len(l), start, len(l) if end is None else end))

    def all_distances(self, g, q, gallery_is_lhs=False):
        g, q = _f32(g), _f32(q)
        out = np.empty((q.shape[0], g.shape[0]), np.float32)
        self.L.fir_ref_all_distances(g, g.shape[0], g.shape[1], q, q.shape[0], int(gallery_is_lhs), out)
        return out

    def bf(self, g, q, labels=None, max_features=0, nthreads=1, timing=False):
        g, q = _f32(g), _f32(q)
        labels = _i32(np.zeros(g.shape[0]) if labels is None else labels)
        idx = np.empty(q.shape[0], np.int32)
        dist = np.empty(q.shape[0], np.float32)
        t = self.L.fir_ref_bf_search(g, labels, g.shape[0], g.shape[1], q, q.shape[0], nthreads, max_features, idx, dist)
        return (idx, dist, t) if timing else (idx, dist)

    def load_split(self, path, d, seed=13, randomize=True):
        h = self.L.fir_ref_dataset_load(path.encode(), d, seed, int(randomize))
        out = []
        for which in (0, 1):
            n = self.L.fir_ref_dataset_count(h, which)
            rows = np.empty((n, d), np.float32)
            labels = np.empty(n, np.int32)
            iid = np.empty(n, np.int32)
            self.L.fir_ref_dataset_get(h, which, rows, labels, iid)
            out += [rows, labels, iid]
        self.L.fir_ref_dataset_free(h)
        return out  # gallery rows, labels, indexInDatabase, test rows, labels, indexInDatabase

    class Dem:
        def __init__(self, L, h, n):
            self.L, self.h, self.n = L, h, n

        @property
        def n_pivots(self):
            return self.L.fir_ref_dem_num_pivots(self.h)

        @property
        def threshold(self):
            return np.float32(self.L.fir_ref_dem_threshold(self.h))

        @property
        def pivots(self):
            out = np.empty(self.n_pivots, np.int32)
            self.L.fir_ref_dem_get_pivots(self.h, out)
            return out

        def P(self, rows=None):
            rows = self.n_pivots if rows is None else rows
            out = np.empty((rows, self.n), np.float32)
            self.L.fir_ref_dem_get_P(self.h, rows, out)
            return out

        def search(self, q, count_to_check=0, nthreads=1, timing=False):
            q = _f32(q)
            nq = q.shape[0]
            idx = np.empty(nq, np.int32)
            dist = np.empty(nq, np.float32)
            below = np.empty(nq, np.uint8)
            evals = np.empty(nq, np.int32)
            t = self.L.fir_ref_dem_search(self.h, q, nq, int(count_to_check), nthreads, idx, dist, below, evals)
            return (idx, dist, below, evals, t) if timing else (idx, dist, below, evals)

        def close(self):
            if self.h:
                self.L.fir_ref_dem_free(self.h)
                self.h = None

    def dem_create(self, g, labels, seed=1, far=0.01, threshold=0.0, count_to_check=0):
        g, labels = _f32(g), _i32(labels)
        secs = C.c_double(0)
        h = self.L.fir_ref_dem_create(g, labels, g.shape[0], g.shape[1], seed, far, threshold, count_to_check, C.byref(secs))
        d = Ref.Dem(self.L, h, g.shape[0])
        d.build_seconds = secs.value
        return d

    def dem_create_injected(self, g, labels, pivots, P, threshold):
        g, labels, pivots, P = _f32(g), _i32(labels), _i32(pivots), _f32(P)
        h = self.L.fir_ref_dem_create_injected(g, labels, g.shape[0], g.shape[1], pivots, len(pivots), P, float(threshold))
        if not h:
            raise ValueError("gallery too small for injection")
        return Ref.Dem(self.L, h, g.shape[0])

    # ---- kNN / PNN (classification.cpp), L2 variant only -------------------------------------
    def cls_setup(self, rows, labels, n_classes, fraction, seed=7):
        rows, labels = _f64(rows), _i32(labels)
        self._cls_d, self._cls_c = rows.shape[1], n_classes
        self.L.fir_ref_cls_setup(rows, labels, rows.shape[0], rows.shape[1], n_classes, float(fraction), seed)
        ntr, nte = self.L.fir_ref_cls_counts(0), self.L.fir_ref_cls_counts(1)
        self._cls_ntrain = ntr
        tr = np.empty(ntr, np.int64)
        trl = np.empty(ntr, np.int32)
        te = np.empty(nte, np.int64)
        avg = np.empty(rows.shape[1], np.float64)
        self.L.fir_ref_cls_get_split(tr, trl, te, avg)
        return tr, trl, te, avg

    def cls_knn(self, K, first, count, timing=False):
        lab = np.empty(count, np.int32)
        t = self.L.fir_ref_cls_knn(K, first, count, 1, lab)
        return (lab, t) if timing else lab

    def cls_std(self):
        sd = np.empty(self._cls_d, np.float64)
        self.L.fir_ref_cls_get_std(sd)
        return sd

    def cls_fpnn(self, first, count, scale=1.0, bf=True, output_ratio=0.9, coefficients=False):
        lab = np.empty(count, np.int32)
        J = C.c_int(0)
        size = self.L.fir_ref_cls_fpnn(float(scale), int(bf), float(output_ratio), 0, 0, lab[:0].copy(), None, C.byref(J))
        a = np.empty(size, np.float64) if coefficients else None
        self.L.fir_ref_cls_fpnn(float(scale), int(bf), float(output_ratio), first, count, lab, a.ctypes.data if coefficients else None, C.byref(J))
        return (lab, a, J.value) if coefficients else lab

    def cls_pnn_clustered(self, num_clusters, first, count):
        lab = np.empty(count, np.int32)
        med = np.empty(self._cls_ntrain, np.int64)
        cnt = self.L.fir_ref_cls_pnn_clustered(int(num_clusters), first, count, lab, med)
        return lab, med[:cnt].copy()

    def cls_pnn_seq(self, first, count, timing=False):
        lab = np.empty(count, np.int32)
        t = self.L.fir_ref_cls_pnn_seq(first, count, lab)
        return (lab, t) if timing else lab

    def cls_pnn(self, first, count, scores=True, timing=False):
        lab = np.empty(count, np.int32)
        sc = np.empty((count, self._cls_c), np.float64) if scores else None
        t = self.L.fir_ref_cls_pnn(first, count, lab, sc.ctypes.data if scores else None)
        if t < 0:
            raise AssertionError("restated PNN scores disagree with verbatim predict_bf label")
        return (lab, sc, t) if timing else (lab, sc)
