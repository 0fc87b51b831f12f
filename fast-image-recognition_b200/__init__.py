"""fast-image-recognition_b200 — B200-native matching engine for the qt_cpp recognizer hot path.

Python face of the C-ABI in include/fir_b200.h (libfir_b200.so, hand-written sm_100a CUDA).  The
classes mirror the reference's C++ entry points for this path:

    Gallery            ~ std::vector<ImageInfo> dbImages            (qt_cpp/db_features.h:14-29)
    Gallery.search     ~ BruteForce::recognize / recognize_image_bf looped by testSetRecognition
                                                                     (qt_cpp/ann.cpp:94-126, db_features.cpp:319-335)
    Gallery.distances  ~ ImageInfo::distance / feature_distance      (qt_cpp/db_features.cpp:22-42)
    Classifier         ~ KNNClassifier / PNNClassifier::predict_bf   (qt_cpp/classification.cpp:116-226)
    Dem                ~ DirectedEnumeration                         (qt_cpp/ann.cpp:270-507)

Arguments may be numpy arrays (host memory: copies happen inside the call, which synchronises) or
CUDA torch tensors (device memory: asynchronous on the gallery's stream).  There is no CPU
fallback: if the CUDA library is missing or no device is present, calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FIR_B200_LIB") or os.path.join(_HERE, "libfir_b200.so")   # override: A/B builds of the same library

L2, CHI2, KL = 0, 1, 2
METRICS = {"l2": L2, "chi2": CHI2, "kl": KL}
HOST, DEVICE = 0, 1
PATH_AUTO, PATH_EXACT, PATH_TENSOR, PATH_APPROX = 0, 1, 2, 3

EXPORTS = [
    "fir_last_error_string", "fir_version", "fir_device_count", "fir_set_device",
    "fir_gallery_create", "fir_gallery_destroy", "fir_gallery_set_stream", "fir_gallery_info", "fir_gallery_index_offset", "fir_gallery_set_num_classes",
    "fir_normalize_rows", "fir_search_topk", "fir_search_last_stats", "fir_pair_distances",
    "fir_class_min", "fir_pnn_scores", "fir_merge_topk", "fir_debug_tensor_candidates", "fir_debug_partition_check", "fir_profile_enable", "fir_profile_read",
    "fir_classifier_create", "fir_classifier_destroy", "fir_classifier_knn", "fir_classifier_pnn", "fir_classifier_pnn_sequential", "fir_classifier_knn_ex", "fir_classifier_pnn_ex", "fir_classifier_set_stream", "fir_classifier_profile",
    "fir_twd_conventional", "fir_twd_proposed", "fir_kmedoids_select", "fir_classifier_set_total",
    "fir_fpnn_create", "fir_fpnn_destroy", "fir_fpnn_info", "fir_fpnn_get_coefficients", "fir_fpnn_predict",
    "fir_dem_build", "fir_dem_from_state", "fir_dem_destroy", "fir_dem_info", "fir_dem_get_pivots", "fir_dem_get_pivot_matrix",
    "fir_dem_get_min_other", "fir_dem_search", "fir_dem_search_stats", "fir_index_save", "fir_index_load", "fir_synth_rows",
    "fir_comm_unique_id", "fir_comm_init_rank", "fir_comm_destroy", "fir_comm_info",
    "fir_shard_search_topk", "fir_shard_class_min", "fir_shard_pnn_scores", "fir_shard_dem_build",
    "fir_sharded_create", "fir_sharded_destroy", "fir_sharded_info", "fir_sharded_shard",
    "fir_sharded_search_topk", "fir_sharded_class_min", "fir_sharded_pnn_scores",
    "fir_sharded_dem_build", "fir_sharded_dem_destroy", "fir_sharded_dem_info", "fir_sharded_dem_get_pivots", "fir_sharded_dem_search",
]


class FirError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("fir_b200 error %d: %s" % (code, msg))
        self.code = code


class SearchStats(C.Structure):
    _fields_ = [("path_used", C.c_int32), ("n_fallback", C.c_int32), ("n_candidates", C.c_int32),
                ("gpu_launches", C.c_int32), ("approx_err_bound", C.c_float), ("reserved", C.c_float)]


class DemParams(C.Structure):
    _fields_ = [("pivot0", C.c_int32), ("seed", C.c_uint32), ("false_accept_rate", C.c_float),
                ("threshold", C.c_float), ("max_chain", C.c_int32), ("max_pivots", C.c_int32)]


_lib = None


def lib():
    """Load libfir_b200.so (built in-tree by build.sh / __graft_entry__.build()). Fails loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FirError(-1, "CUDA extension %s is missing; run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.fir_last_error_string.restype = C.c_char_p
    vp, i32, i64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    L.fir_gallery_create.argtypes = [vp, vp, i64, i32, i32, i32, i64, C.POINTER(vp)]
    L.fir_gallery_destroy.argtypes = [vp]
    L.fir_gallery_set_stream.argtypes = [vp, vp]
    L.fir_gallery_info.argtypes = [vp, C.POINTER(i64), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.fir_gallery_index_offset.argtypes = [vp, C.POINTER(i64)]
    L.fir_gallery_set_num_classes.argtypes = [vp, i32]
    L.fir_normalize_rows.argtypes = [vp, i64, i32, i32, i32, vp]
    L.fir_search_topk.argtypes = [vp, vp, i64, i32, i32, i32, i32, vp, vp]
    L.fir_search_last_stats.argtypes = [vp, C.POINTER(SearchStats)]
    L.fir_debug_tensor_candidates.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), vp, vp, vp]
    L.fir_debug_partition_check.argtypes = [C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int64, C.POINTER(i32), C.POINTER(i32)]
    L.fir_profile_enable.argtypes = [vp, i32]
    L.fir_profile_read.argtypes = [vp, i32, C.POINTER(f64), C.POINTER(i32)]
    L.fir_pair_distances.argtypes = [vp, vp, i64, vp, i32, i32, i32, vp]
    L.fir_class_min.argtypes = [vp, vp, i64, i32, vp, vp]
    L.fir_pnn_scores.argtypes = [vp, vp, i64, f64, i64, i32, vp, vp]
    L.fir_merge_topk.argtypes = [vp, vp, i32, i64, i32, vp, vp, vp]
    L.fir_classifier_create.argtypes = [vp, vp, i64, i32, i32, vp, C.POINTER(vp)]
    L.fir_classifier_destroy.argtypes = [vp]
    L.fir_classifier_knn.argtypes = [vp, vp, i64, i32, vp]
    L.fir_classifier_pnn.argtypes = [vp, vp, i64, vp, vp]
    L.fir_classifier_pnn_sequential.argtypes = [vp, vp, i64, vp]
    L.fir_classifier_knn_ex.argtypes = [vp, vp, i64, i32, i32, vp]
    L.fir_classifier_pnn_ex.argtypes = [vp, vp, i64, i32, vp, vp]
    L.fir_classifier_set_stream.argtypes = [vp, vp]
    L.fir_classifier_profile.argtypes = [vp, i32, C.POINTER(f64), C.POINTER(i32)]
    L.fir_kmedoids_select.argtypes = [vp, vp, i64, i32, i32, i32, vp, C.POINTER(i64)]
    L.fir_classifier_set_total.argtypes = [vp, i64]
    L.fir_fpnn_create.argtypes = [vp, vp, i64, i32, i32, vp, vp, f64, C.POINTER(vp)]
    L.fir_fpnn_destroy.argtypes = [vp]
    L.fir_fpnn_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i64)]
    L.fir_fpnn_get_coefficients.argtypes = [vp, vp]
    L.fir_fpnn_predict.argtypes = [vp, vp, i64, i32, C.c_float, vp]
    L.fir_twd_conventional.argtypes = [vp, vp, i64, i32, C.c_double, i32, i32, i32, vp, vp, vp]
    L.fir_twd_proposed.argtypes = [vp, vp, i64, i32, C.c_double, i32, i32, vp, vp, vp]
    L.fir_dem_build.argtypes = [vp, C.POINTER(DemParams), C.POINTER(vp)]
    L.fir_index_save.argtypes = [vp, vp, C.c_char_p]
    L.fir_index_load.argtypes = [C.c_char_p, C.POINTER(vp), C.POINTER(vp)]
    L.fir_dem_from_state.argtypes = [vp, vp, i32, vp, C.c_float, C.POINTER(vp)]
    L.fir_dem_destroy.argtypes = [vp]
    L.fir_dem_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(C.c_float)]
    L.fir_dem_get_pivots.argtypes = [vp, vp]
    L.fir_dem_get_pivot_matrix.argtypes = [vp, vp]
    L.fir_dem_get_min_other.argtypes = [vp, vp]
    L.fir_dem_search.argtypes = [vp, vp, i64, i32, i32, vp, vp, vp, vp]
    L.fir_dem_search_stats.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(f64), C.POINTER(i32)]
    L.fir_synth_rows.argtypes = [vp, vp, i64, i64, i64, i32, i32, i32, C.c_uint32, C.c_float, i32, vp]
    L.fir_comm_unique_id.argtypes = [vp]
    L.fir_comm_init_rank.argtypes = [vp, i32, i32, C.POINTER(vp)]
    L.fir_comm_destroy.argtypes = [vp]
    L.fir_comm_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.fir_shard_search_topk.argtypes = [vp, vp, vp, i64, i32, i32, i32, vp, vp]
    L.fir_shard_class_min.argtypes = [vp, vp, vp, i64, i32, vp, vp]
    L.fir_shard_pnn_scores.argtypes = [vp, vp, vp, i64, f64, i64, i32, vp, vp]
    L.fir_shard_dem_build.argtypes = [vp, vp, i64, C.POINTER(DemParams), C.POINTER(vp)]
    L.fir_sharded_create.argtypes = [vp, vp, i64, i32, i32, i32, C.POINTER(vp)]
    L.fir_sharded_destroy.argtypes = [vp]
    L.fir_sharded_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i64), C.POINTER(i32), C.POINTER(i32)]
    L.fir_sharded_shard.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(vp)]
    L.fir_sharded_search_topk.argtypes = [vp, vp, i64, i32, i32, vp, vp]
    L.fir_sharded_class_min.argtypes = [vp, vp, i64, vp, vp]
    L.fir_sharded_pnn_scores.argtypes = [vp, vp, i64, f64, vp, vp]
    L.fir_sharded_dem_build.argtypes = [vp, C.POINTER(DemParams), C.POINTER(vp)]
    L.fir_sharded_dem_destroy.argtypes = [vp]
    L.fir_sharded_dem_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(C.c_float)]
    L.fir_sharded_dem_get_pivots.argtypes = [vp, vp]
    L.fir_sharded_dem_search.argtypes = [vp, vp, i64, i32, vp, vp, vp, vp]
    _lib = L
    return L


def _check(code):
    if code != 0:
        raise FirError(code, lib().fir_last_error_string().decode(errors="replace"))


def _is_torch(a):
    return type(a).__module__.startswith("torch")


def _ptr(a):
    if a is None:
        return None
    if _is_torch(a):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(a.ctypes.data)


def _prep(a, dtype_np, dtype_name):
    """Returns (array, memspace). numpy → contiguous host array; torch CUDA tensor → contiguous device tensor."""
    if _is_torch(a):
        import torch
        want = getattr(torch, dtype_name)
        if not a.is_cuda:
            return np.ascontiguousarray(a.numpy(), dtype=dtype_np), HOST
        if a.dtype != want:
            a = a.to(want)
        return a.contiguous(), DEVICE
    return np.ascontiguousarray(a, dtype=dtype_np), HOST


def _out(shape, dtype_np, dtype_name, like_device, device=None):
    if like_device:
        import torch
        return torch.empty(shape, dtype=getattr(torch, dtype_name), device=device)
    return np.empty(shape, dtype=dtype_np)


def _check_out(buf, shape, dtype_name, space, like):
    """A caller-supplied result buffer goes to the C ABI as a raw pointer: wrong shape / dtype / memory space would corrupt memory."""
    if _is_torch(buf):
        ok = buf.is_cuda == (space == DEVICE) and tuple(buf.shape) == tuple(shape) and str(buf.dtype) == "torch." + dtype_name and buf.is_contiguous()
        ok = ok and (space != DEVICE or buf.device == like.device)
    else:
        ok = space == HOST and isinstance(buf, np.ndarray) and buf.shape == tuple(shape) and buf.dtype == np.dtype(dtype_name) and buf.flags["C_CONTIGUOUS"]
    if not ok:
        raise ValueError("out buffer must be a contiguous %s array of shape %s in the same memory space as the queries" % (dtype_name, tuple(shape)))


def device_count():
    n = C.c_int(0)
    _check(lib().fir_device_count(C.byref(n)))
    return n.value


def normalize_rows(rows, metric="l2", stream=None):
    """Loader normalisation of db_features.cpp:79-101, in place (numpy array or CUDA tensor)."""
    m = METRICS[metric] if isinstance(metric, str) else metric
    if _is_torch(rows) and rows.is_cuda:
        assert rows.is_contiguous() and rows.dtype.is_floating_point and rows.element_size() == 4
        _check(lib().fir_normalize_rows(_ptr(rows), rows.shape[0], rows.shape[1], m, DEVICE, stream))
        return rows
    assert rows.dtype == np.float32 and rows.flags["C_CONTIGUOUS"]
    _check(lib().fir_normalize_rows(_ptr(rows), rows.shape[0], rows.shape[1], m, HOST, stream))
    return rows


class Gallery:
    @classmethod
    def _adopt(cls, handle):
        """Wrap a handle created by the library (fir_index_load)."""
        self = cls.__new__(cls)
        self._h = handle
        n, d, m, nc = C.c_int64(0), C.c_int32(0), C.c_int32(0), C.c_int32(0)
        _check(lib().fir_gallery_info(self._h, C.byref(n), C.byref(d), C.byref(m), C.byref(nc)))
        self.n, self.d, self.n_classes = n.value, d.value, nc.value
        self.metric_name = {v: k for k, v in METRICS.items()}[m.value]
        off = C.c_int64(0)
        _check(lib().fir_gallery_index_offset(self._h, C.byref(off)))
        self.index_offset = off.value                       # a loaded index keeps the offset it was saved with
        self._device = None
        return self

    def __init__(self, rows, labels=None, metric="l2", index_offset=0, stream=None):
        self.metric_name = metric if isinstance(metric, str) else {v: k for k, v in METRICS.items()}[metric]
        m = METRICS[metric] if isinstance(metric, str) else metric
        rows, space = _prep(rows, np.float32, "float32")
        lab = None
        if labels is not None:
            lab, lspace = _prep(labels, np.int32, "int32")
            if lspace != space:
                raise ValueError("rows and labels must live in the same memory space")
        self.n, self.d = int(rows.shape[0]), int(rows.shape[1])
        self.index_offset = int(index_offset)
        self._device = rows.device if space == DEVICE else None
        h = C.c_void_p(None)
        _check(lib().fir_gallery_create(_ptr(rows), _ptr(lab), self.n, self.d, m, space, self.index_offset, C.byref(h)))
        self._h = h
        nc = C.c_int32(0)
        _check(lib().fir_gallery_info(self._h, None, None, None, C.byref(nc)))
        self.n_classes = nc.value
        if stream is not None:
            self.set_stream(stream)

    def set_num_classes(self, n_classes):
        _check(lib().fir_gallery_set_num_classes(self._h, int(n_classes)))
        self.n_classes = int(n_classes)

    def set_stream(self, stream):
        _check(lib().fir_gallery_set_stream(self._h, C.c_void_p(int(stream))))

    def close(self):
        if getattr(self, "_h", None):
            lib().fir_gallery_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def search(self, queries, k=1, max_features=0, path=PATH_AUTO, out=None):
        """k smallest (feature_distance, index) per query; k=1 is BruteForce::recognize.  out=(idx, dist) reuses result buffers."""
        q, space = _prep(queries, np.float32, "float32")
        nq = int(q.shape[0])
        if out is not None:
            idx, dist = out
            _check_out(idx, (nq, k), "int32", space, q)
            _check_out(dist, (nq, k), "float32", space, q)
        else:
            idx = _out((nq, k), np.int32, "int32", space == DEVICE, getattr(q, "device", None))
            dist = _out((nq, k), np.float32, "float32", space == DEVICE, getattr(q, "device", None))
        _check(lib().fir_search_topk(self._h, _ptr(q), nq, k, max_features, path, space, _ptr(idx), _ptr(dist)))
        return idx, dist

    def stats(self):
        s = SearchStats()
        _check(lib().fir_search_last_stats(self._h, C.byref(s)))
        return {f[0]: getattr(s, f[0]) for f in SearchStats._fields_}

    def profile(self, on=True):
        _check(lib().fir_profile_enable(self._h, int(on)))

    def profile_read(self, kernel=0):
        """(total ms, launches) of the given kernel class since profile(True)."""
        ms, n = C.c_double(0), C.c_int32(0)
        _check(lib().fir_profile_read(self._h, kernel, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def debug_candidates(self, nq):
        """(idx, approx d^2, exact feature_distance) of the last tensor-path search: arrays [nq, n_slots*R]."""
        ns, r = C.c_int32(0), C.c_int32(0)
        _check(lib().fir_debug_tensor_candidates(self._h, C.byref(ns), C.byref(r), None, None, None))
        shape = (nq, ns.value * r.value)
        idx = np.empty(shape, np.int32)
        approx = np.empty(shape, np.float32)
        exact = np.empty(shape, np.float32)
        _check(lib().fir_debug_tensor_candidates(self._h, None, None, _ptr(idx), _ptr(approx), _ptr(exact)))
        return idx, approx, exact

    def distances(self, queries, cand_idx, gallery_is_lhs=False):
        q, space = _prep(queries, np.float32, "float32")
        c, cspace = _prep(cand_idx, np.int32, "int32")
        if cspace != space:
            raise ValueError("queries and cand_idx must live in the same memory space")
        nq, r = int(c.shape[0]), int(c.shape[1])
        out = _out((nq, r), np.float32, "float32", space == DEVICE, getattr(q, "device", None))
        _check(lib().fir_pair_distances(self._h, _ptr(q), nq, _ptr(c), r, int(gallery_is_lhs), space, _ptr(out)))
        return out

    def class_min(self, queries):
        q, space = _prep(queries, np.float32, "float32")
        nq = int(q.shape[0])
        mn = _out((nq, self.n_classes), np.float32, "float32", space == DEVICE, getattr(q, "device", None))
        arg = _out((nq, self.n_classes), np.int32, "int32", space == DEVICE, getattr(q, "device", None))
        _check(lib().fir_class_min(self._h, _ptr(q), nq, space, _ptr(mn), _ptr(arg)))
        return mn, arg

    TWD_TYPES = {"posteriors": 0, "diff": 1, "ratio": 2}

    def _twd_out(self, q, space):
        nq = int(q.shape[0])
        dev = getattr(q, "device", None)
        return (_out((nq,), np.int32, "int32", space == DEVICE, dev), _out((nq,), np.int32, "int32", space == DEVICE, dev),
                _out((nq,), np.uint8, "uint8", space == DEVICE, dev))

    def twd_conventional(self, queries, kind, threshold, feat_count=64, last_feature=256):
        """ConventionalTWDClassifier(cls_num, kind, threshold, feat_count).recognize → (index, class, unreliable)."""
        q, space = _prep(queries, np.float32, "float32")
        idx, lab, unrel = self._twd_out(q, space)
        _check(lib().fir_twd_conventional(self._h, _ptr(q), int(q.shape[0]), self.TWD_TYPES[kind], float(threshold), int(feat_count),
                                          int(last_feature), space, _ptr(idx), _ptr(lab), _ptr(unrel)))
        return idx, lab, unrel

    def twd_proposed(self, queries, feat_count, threshold, last_feature=256):
        """ProposedTWDClassifier(cls_num, feat_count, threshold).recognize → (index, class, unreliable)."""
        q, space = _prep(queries, np.float32, "float32")
        idx, lab, unrel = self._twd_out(q, space)
        _check(lib().fir_twd_proposed(self._h, _ptr(q), int(q.shape[0]), int(feat_count), float(threshold), int(last_feature), space,
                                      _ptr(idx), _ptr(lab), _ptr(unrel)))
        return idx, lab, unrel

    def pnn_scores(self, queries, var, n_total=0):
        q, space = _prep(queries, np.float32, "float32")
        nq = int(q.shape[0])
        sc = _out((nq, self.n_classes), np.float64, "float64", space == DEVICE, getattr(q, "device", None))
        lab = _out((nq,), np.int32, "int32", space == DEVICE, getattr(q, "device", None))
        _check(lib().fir_pnn_scores(self._h, _ptr(q), nq, float(var), int(n_total), space, _ptr(sc), _ptr(lab)))
        return sc, lab


def save_index(path, gallery, dem=None):
    """Gallery (+ optional DirectedEnumeration state) → one binary file."""
    _check(lib().fir_index_save(gallery._h, dem._h if dem is not None else None, os.fsencode(path)))


def load_index(path):
    """→ (Gallery, Dem or None)"""
    g, d = C.c_void_p(None), C.c_void_p(None)
    _check(lib().fir_index_load(os.fsencode(path), C.byref(g), C.byref(d)))
    gal = Gallery._adopt(g)
    return gal, (Dem._adopt(gal, d) if d.value else None)


def merge_topk(parts_dist, parts_idx, stream=None):
    """k-way (dist, idx) merge of per-shard lists: CUDA tensors [n_parts, nq, k] → ([nq,k], [nq,k])."""
    import torch
    assert parts_dist.is_cuda and parts_idx.is_cuda
    pd, pi = parts_dist.contiguous(), parts_idx.contiguous()
    n_parts, nq, k = pd.shape
    od = torch.empty((nq, k), dtype=torch.float32, device=pd.device)
    oi = torch.empty((nq, k), dtype=torch.int32, device=pd.device)
    s = stream if stream is not None else torch.cuda.current_stream().cuda_stream
    _check(lib().fir_merge_topk(_ptr(pd), _ptr(pi), n_parts, nq, k, _ptr(od), _ptr(oi), C.c_void_p(int(s))))
    return oi, od


class Classifier:
    """fp64 kNN / PNN of classification.cpp over a class-major training set."""

    def __init__(self, train_rows, train_labels, n_classes, avg):
        tr = np.ascontiguousarray(train_rows, dtype=np.float64)
        tl = np.ascontiguousarray(train_labels, dtype=np.int32)
        av = np.ascontiguousarray(avg, dtype=np.float64)
        self.n, self.d, self.n_classes = int(tr.shape[0]), int(tr.shape[1]), int(n_classes)
        h = C.c_void_p(None)
        _check(lib().fir_classifier_create(_ptr(tr), _ptr(tl), self.n, self.d, self.n_classes, _ptr(av), C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib().fir_classifier_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def knn(self, queries, K):
        q, space = _prep(queries, np.float64, "float64")
        lab = _out((q.shape[0],), np.int32, "int32", space == DEVICE, getattr(q, "device", None))
        _check(lib().fir_classifier_knn_ex(self._h, _ptr(q), q.shape[0], int(K), space, _ptr(lab)))
        return lab

    def set_stream(self, stream):
        _check(lib().fir_classifier_set_stream(self._h, C.c_void_p(int(stream))))

    def profile(self, on=True):
        """Switch the distance-kernel timing on/off; returns (ms, launches) accumulated since it was last switched on."""
        ms, n = C.c_double(0), C.c_int32(0)
        _check(lib().fir_classifier_profile(self._h, int(on), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def set_total(self, n_total):
        """PNN denominator when the rows are a reduced (clustered) training set."""
        _check(lib().fir_classifier_set_total(self._h, int(n_total)))

    def pnn_sequential(self, queries):
        """PNNClassifier(bf=False): predict_sequentional."""
        q = np.ascontiguousarray(queries, dtype=np.float64)
        lab = np.empty(q.shape[0], np.int32)
        _check(lib().fir_classifier_pnn_sequential(self._h, _ptr(q), q.shape[0], _ptr(lab)))
        return lab

    def pnn(self, queries, scores=True):
        q, space = _prep(queries, np.float64, "float64")
        dev = getattr(q, "device", None)
        lab = _out((q.shape[0],), np.int32, "int32", space == DEVICE, dev)
        sc = _out((q.shape[0], self.n_classes), np.float64, "float64", space == DEVICE, dev) if scores else None
        _check(lib().fir_classifier_pnn_ex(self._h, _ptr(q), q.shape[0], space, _ptr(sc), _ptr(lab)))
        return lab, sc


def kmedoids_select(train_rows, train_labels, n_classes, num_clusters):
    """PNNwithClusteringClassifier::train: positions (in the given class-major order) of the rows kept per class."""
    rows = np.ascontiguousarray(train_rows, dtype=np.float64)
    lab = np.ascontiguousarray(train_labels, dtype=np.int32)
    sel = np.empty(rows.shape[0], np.int64)
    cnt = C.c_int64(0)
    _check(lib().fir_kmedoids_select(_ptr(rows), _ptr(lab), rows.shape[0], rows.shape[1], int(n_classes), int(num_clusters), _ptr(sel), C.byref(cnt)))
    return sel[:cnt.value].copy()


class Fpnn:
    """FPNNClassifier (orthogonal-series PNN): trained at construction from RAW class-major rows."""

    def __init__(self, train_rows, train_labels, n_classes, avg, std, scale=1.0):
        rows = np.ascontiguousarray(train_rows, dtype=np.float64)
        lab = np.ascontiguousarray(train_labels, dtype=np.int32)
        avg, std = np.ascontiguousarray(avg, dtype=np.float64), np.ascontiguousarray(std, dtype=np.float64)
        h = C.c_void_p(None)
        _check(lib().fir_fpnn_create(_ptr(rows), _ptr(lab), rows.shape[0], rows.shape[1], int(n_classes), _ptr(avg), _ptr(std), float(scale), C.byref(h)))
        self._h = h
        J, na = C.c_int32(0), C.c_int64(0)
        _check(lib().fir_fpnn_info(self._h, C.byref(J), C.byref(na)))
        self.J, self.n_coefficients = J.value, na.value

    @property
    def coefficients(self):
        a = np.empty(self.n_coefficients, np.float64)
        _check(lib().fir_fpnn_get_coefficients(self._h, _ptr(a)))
        return a

    def predict(self, queries, sequential=False, output_ratio=0.9):
        q = np.ascontiguousarray(queries, dtype=np.float64)
        lab = np.empty(q.shape[0], np.int32)
        _check(lib().fir_fpnn_predict(self._h, _ptr(q), q.shape[0], int(bool(sequential)), float(output_ratio), _ptr(lab)))
        return lab

    def close(self):
        if getattr(self, "_h", None):
            lib().fir_fpnn_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Dem:
    """DirectedEnumeration over a Gallery (kept alive by this object)."""

    @classmethod
    def _adopt(cls, gallery, handle):
        self = cls.__new__(cls)
        self.gallery, self._h = gallery, handle
        self._read_info()
        return self

    def _read_info(self):
        a, b, t = C.c_int32(0), C.c_int32(0), C.c_float(0)
        _check(lib().fir_dem_info(self._h, C.byref(a), C.byref(b), C.byref(t)))
        self.n_pivots, self.chain_rows, self.threshold = a.value, b.value, np.float32(t.value)

    def __init__(self, gallery, pivot0=-1, seed=0, false_accept_rate=0.01, threshold=0.0, max_chain=0, max_pivots=0, state=None):
        self.gallery = gallery
        h = C.c_void_p(None)
        if state is not None:      # (pivots, P, threshold): adopt an existing build
            piv = np.ascontiguousarray(state[0], dtype=np.int32)
            P = np.ascontiguousarray(state[1], dtype=np.float32)
            _check(lib().fir_dem_from_state(gallery._h, _ptr(piv), len(piv), _ptr(P), float(state[2]), C.byref(h)))
        else:
            p = DemParams(int(pivot0), int(seed), float(false_accept_rate), float(threshold), int(max_chain), int(max_pivots))
            _check(lib().fir_dem_build(gallery._h, C.byref(p), C.byref(h)))
        self._h = h
        self._read_info()

    def close(self):
        if getattr(self, "_h", None):
            lib().fir_dem_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def pivots(self):
        out = np.empty(self.n_pivots, np.int32)
        _check(lib().fir_dem_get_pivots(self._h, _ptr(out)))
        return out

    @property
    def P(self):
        out = np.empty((self.n_pivots, self.gallery.n), np.float32)
        _check(lib().fir_dem_get_pivot_matrix(self._h, _ptr(out)))
        return out

    @property
    def min_other(self):
        out = np.empty(self.chain_rows, np.float32)
        _check(lib().fir_dem_get_min_other(self._h, _ptr(out)))
        return out

    def search_stats(self):
        """Counters of the last search: kernels launched, tensor first round on/off, its candidate kernel's (ms, launches) while profiling."""
        a, b, ms, n = C.c_int32(0), C.c_int32(0), C.c_double(0), C.c_int32(0)
        _check(lib().fir_dem_search_stats(self._h, C.byref(a), C.byref(b), C.byref(ms), C.byref(n)))
        return {"gpu_launches": a.value, "tensor_round": bool(b.value), "candidates_kernel_ms": ms.value, "candidates_kernel_launches": n.value}

    def search(self, queries, count_to_check=0):
        q, space = _prep(queries, np.float32, "float32")
        nq = int(q.shape[0])
        dev = getattr(q, "device", None)
        idx = _out((nq,), np.int32, "int32", space == DEVICE, dev)
        dist = _out((nq,), np.float32, "float32", space == DEVICE, dev)
        below = _out((nq,), np.uint8, "uint8", space == DEVICE, dev)
        evals = _out((nq,), np.int32, "int32", space == DEVICE, dev)
        _check(lib().fir_dem_search(self._h, _ptr(q), nq, int(count_to_check), space, _ptr(idx), _ptr(dist), _ptr(below), _ptr(evals)))
        return idx, dist, below, evals


COMM_ID_BYTES = 128


class Comm:
    """One rank of the library's NCCL communicator (fir_comm): one per process / GPU."""

    def __init__(self, rank=0, world=1, comm_id=None):
        h = C.c_void_p(None)
        buf = (C.c_char * COMM_ID_BYTES).from_buffer_copy(comm_id) if comm_id is not None else None
        _check(lib().fir_comm_init_rank(buf, int(rank), int(world), C.byref(h)))
        self._h, self.rank, self.world = h, int(rank), int(world)

    @staticmethod
    def unique_id():
        buf = (C.c_char * COMM_ID_BYTES)()
        _check(lib().fir_comm_unique_id(buf))
        return bytes(buf)

    @classmethod
    def from_torch_distributed(cls, dist):
        """The launcher's process group only carries the 128-byte id from rank 0 to the others."""
        if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
            return cls(0, 1)
        box = [cls.unique_id() if dist.get_rank() == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return cls(dist.get_rank(), dist.get_world_size(), box[0])

    @property
    def nccl_version(self):
        v = C.c_int32(0)
        _check(lib().fir_comm_info(self._h, None, None, C.byref(v)))
        return v.value

    def close(self):
        if getattr(self, "_h", None):
            lib().fir_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RankShard:
    """This rank's row shard of a gallery + the communicator: collective search / class reductions (fir_shard_*)."""

    def __init__(self, gallery, comm, n_total):
        self.gallery, self.comm, self.n_total = gallery, comm, int(n_total)

    def search(self, queries, k=1, path=PATH_AUTO, out=None):
        q, space = _prep(queries, np.float32, "float32")
        nq = int(q.shape[0])
        if out is not None:
            idx, dist = out
            _check_out(idx, (nq, k), "int32", space, q)
            _check_out(dist, (nq, k), "float32", space, q)
        else:
            idx = _out((nq, k), np.int32, "int32", space == DEVICE, getattr(q, "device", None))
            dist = _out((nq, k), np.float32, "float32", space == DEVICE, getattr(q, "device", None))
        _check(lib().fir_shard_search_topk(self.gallery._h, self.comm._h, _ptr(q), nq, k, path, space, _ptr(idx), _ptr(dist)))
        return idx, dist

    def dem(self, pivot0=-1, seed=0, false_accept_rate=0.01, threshold=0.0, max_chain=0, max_pivots=0):
        """Collective: ONE DirectedEnumeration over the whole sharded gallery; .search() on the result is collective too."""
        p = DemParams(int(pivot0), int(seed), float(false_accept_rate), float(threshold), int(max_chain), int(max_pivots))
        h = C.c_void_p(None)
        _check(lib().fir_shard_dem_build(self.gallery._h, self.comm._h, self.n_total, C.byref(p), C.byref(h)))
        return Dem._adopt(self.gallery, h)

    def class_min(self, queries):
        q, space = _prep(queries, np.float32, "float32")
        nq, nc = int(q.shape[0]), self.gallery.n_classes
        mn = _out((nq, nc), np.float32, "float32", space == DEVICE, getattr(q, "device", None))
        arg = _out((nq, nc), np.int32, "int32", space == DEVICE, getattr(q, "device", None))
        _check(lib().fir_shard_class_min(self.gallery._h, self.comm._h, _ptr(q), nq, space, _ptr(mn), _ptr(arg)))
        return mn, arg

    def pnn_scores(self, queries, var):
        q, space = _prep(queries, np.float32, "float32")
        nq, nc = int(q.shape[0]), self.gallery.n_classes
        sc = _out((nq, nc), np.float64, "float64", space == DEVICE, getattr(q, "device", None))
        lab = _out((nq,), np.int32, "int32", space == DEVICE, getattr(q, "device", None))
        _check(lib().fir_shard_pnn_scores(self.gallery._h, self.comm._h, _ptr(q), nq, float(var), self.n_total, space, _ptr(sc), _ptr(lab)))
        return sc, lab


class Sharded:
    """The whole class-major gallery (host arrays) row-sharded over the GPUs of the box by ONE process (fir_sharded_*)."""

    def __init__(self, rows, labels=None, metric="l2", n_gpus=0):
        m = METRICS[metric] if isinstance(metric, str) else metric
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        lab = np.ascontiguousarray(labels, dtype=np.int32) if labels is not None else None
        h = C.c_void_p(None)
        _check(lib().fir_sharded_create(_ptr(rows), _ptr(lab), rows.shape[0], rows.shape[1], m, int(n_gpus), C.byref(h)))
        self._h = h
        g, n, d, nc = C.c_int32(0), C.c_int64(0), C.c_int32(0), C.c_int32(0)
        _check(lib().fir_sharded_info(self._h, C.byref(g), C.byref(n), C.byref(d), C.byref(nc)))
        self.n_gpus, self.n, self.d, self.n_classes = g.value, n.value, d.value, nc.value

    def search(self, queries, k=1, path=PATH_AUTO):
        q = np.ascontiguousarray(queries, dtype=np.float32)
        idx, dist = np.empty((q.shape[0], k), np.int32), np.empty((q.shape[0], k), np.float32)
        _check(lib().fir_sharded_search_topk(self._h, _ptr(q), q.shape[0], k, path, _ptr(idx), _ptr(dist)))
        return idx, dist

    def class_min(self, queries):
        q = np.ascontiguousarray(queries, dtype=np.float32)
        mn, arg = np.empty((q.shape[0], self.n_classes), np.float32), np.empty((q.shape[0], self.n_classes), np.int32)
        _check(lib().fir_sharded_class_min(self._h, _ptr(q), q.shape[0], _ptr(mn), _ptr(arg)))
        return mn, arg

    def pnn_scores(self, queries, var):
        q = np.ascontiguousarray(queries, dtype=np.float32)
        sc, lab = np.empty((q.shape[0], self.n_classes), np.float64), np.empty(q.shape[0], np.int32)
        _check(lib().fir_sharded_pnn_scores(self._h, _ptr(q), q.shape[0], float(var), _ptr(sc), _ptr(lab)))
        return sc, lab

    def dem(self, pivot0=-1, seed=0, false_accept_rate=0.01, threshold=0.0, max_chain=0, max_pivots=0):
        """ONE DirectedEnumeration over the whole sharded gallery → ShardedDem."""
        return ShardedDem(self, DemParams(int(pivot0), int(seed), float(false_accept_rate), float(threshold), int(max_chain), int(max_pivots)))

    def close(self):
        if getattr(self, "_h", None):
            lib().fir_sharded_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShardedDem:
    """DirectedEnumeration over a Sharded gallery (fir_sharded_dem_*); keeps the gallery alive."""

    def __init__(self, sharded, params):
        self.sharded = sharded
        h = C.c_void_p(None)
        _check(lib().fir_sharded_dem_build(sharded._h, C.byref(params), C.byref(h)))
        self._h = h
        a, b, t = C.c_int32(0), C.c_int32(0), C.c_float(0)
        _check(lib().fir_sharded_dem_info(self._h, C.byref(a), C.byref(b), C.byref(t)))
        self.n_pivots, self.chain_rows, self.threshold = a.value, b.value, np.float32(t.value)

    @property
    def pivots(self):
        out = np.empty(self.n_pivots, np.int32)
        _check(lib().fir_sharded_dem_get_pivots(self._h, _ptr(out)))
        return out

    def search(self, queries, count_to_check=0):
        q = np.ascontiguousarray(queries, dtype=np.float32)
        nq = q.shape[0]
        idx, dist = np.empty(nq, np.int32), np.empty(nq, np.float32)
        below, evals = np.empty(nq, np.uint8), np.empty(nq, np.int32)
        _check(lib().fir_sharded_dem_search(self._h, _ptr(q), nq, int(count_to_check), _ptr(idx), _ptr(dist), _ptr(below), _ptr(evals)))
        return idx, dist, below, evals

    def close(self):
        if getattr(self, "_h", None):
            lib().fir_sharded_dem_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
