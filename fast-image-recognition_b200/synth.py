"""Synthetic inputs of the shapes BASELINE.json names.

Clustered embeddings: C class centroids ~ N(0, I_D); sample = centroid[label] + sigma * N(0, I_D).
L2 configs keep the signed features; chi2/KL configs take ReLU ("EfficientNet-style" non-negative
features, about half exact zeros).  The caller then applies the reference's loader normalisation
(qt_cpp/db_features.cpp:79-101) — fir_b200.normalize_rows on the GPU.  Gallery rows are stored
class-major (stable sort by label) like getTrainingAndTestImages leaves them (db_features.cpp:129-160).
Counter-based Philox streams keyed by (seed, role) make every array reproducible independently.
"""
import numpy as np

SEED_GALLERY, SEED_QUERY, SEED_LABEL, SEED_CENTROID = 0x5EED0001, 0x5EED0002, 0x5EED0003, 0x5EED0004


def _rng(seed, stream=0):
    return np.random.Generator(np.random.Philox(key=[seed, stream]))


def centroids(n_classes, d, seed=SEED_CENTROID):
    return _rng(seed).standard_normal((n_classes, d), dtype=np.float32)


def make_split(n_gallery, n_query, d, n_classes, metric="l2", sigma=0.5, seed=0, shard=0):
    """Raw (un-normalised) fp32 features: gallery rows class-major, labels int32.  `shard` draws a different gallery
    (labels + noise) around the SAME class centroids and with the SAME queries — one shard per GPU of a sharded run."""
    cen = centroids(n_classes, d, SEED_CENTROID + seed)
    gl = np.sort(_rng(SEED_LABEL + seed, 2 * shard).integers(0, n_classes, n_gallery, dtype=np.int32), kind="stable")
    ql = _rng(SEED_LABEL + seed, 1).integers(0, n_classes, n_query, dtype=np.int32)
    g = cen[gl] + sigma * _rng(SEED_GALLERY + seed, shard).standard_normal((n_gallery, d), dtype=np.float32)
    q = cen[ql] + sigma * _rng(SEED_QUERY + seed).standard_normal((n_query, d), dtype=np.float32)
    if metric != "l2":
        np.maximum(g, 0, out=g)
        np.maximum(q, 0, out=q)
    return g.astype(np.float32), gl, q.astype(np.float32), ql


def make_split_device(n_gallery, n_query, d, n_classes, metric="l2", sigma=0.5, seed=0, device="cuda"):
    """Same distribution generated on the GPU with torch (plumbing) for galleries too large to build on
    the host in reasonable time (10M x 512).  Not bit-identical to make_split."""
    import torch
    gen = torch.Generator(device=device)
    gen.manual_seed(SEED_CENTROID + seed)
    cen = torch.randn((n_classes, d), generator=gen, device=device, dtype=torch.float32)
    gen.manual_seed(SEED_LABEL + seed)
    gl = torch.sort(torch.randint(0, n_classes, (n_gallery,), generator=gen, device=device, dtype=torch.int32)).values
    ql = torch.randint(0, n_classes, (n_query,), generator=gen, device=device, dtype=torch.int32)
    gen.manual_seed(SEED_GALLERY + seed)
    g = torch.empty((n_gallery, d), device=device, dtype=torch.float32)
    step = 1 << 20
    for lo in range(0, n_gallery, step):
        hi = min(n_gallery, lo + step)
        g[lo:hi] = cen[gl[lo:hi].long()] + sigma * torch.randn((hi - lo, d), generator=gen, device=device, dtype=torch.float32)
    gen.manual_seed(SEED_QUERY + seed)
    q = cen[ql.long()] + sigma * torch.randn((n_query, d), generator=gen, device=device, dtype=torch.float32)
    if metric != "l2":
        g.clamp_(min=0)
        q.clamp_(min=0)
    return g, gl, q, ql


def caltech_sizes():
    """Class sizes of a Caltech-101-shaped collection (qt_cpp/db.h:11,35,86): 101 classes, 8677 images, four classes above the
    loader's 400-image cap (db_features.cpp:119,147-149), the rest between 31 and 239 — the real list is not in the reference
    tree, so the tail is a fixed synthetic sequence with the same total."""
    big = [800, 798, 435, 435]
    rest = [31 + (i * 37) % 60 + (i % 7) * 5 for i in range(97)]
    rest[0] = 239
    fix = 8677 - sum(big) - sum(rest)
    i = 1
    while fix != 0:
        step = 1 if fix > 0 else -1
        if 31 <= rest[i] + step <= 239:
            rest[i] += step
            fix -= step
        i = 1 + i % 96
    return big + rest


def write_features_file(path, rows, class_names, file_names=None):
    """The reference's text format: 3 lines per image — file name, class name, D floats printed as
    '{:f} ' (qt_cpp/dnn_feature_extractor.py:58-64; parsed by qt_cpp/db_features.cpp:52-57)."""
    with open(path, "w") as f:
        for i, row in enumerate(rows):
            f.write((file_names[i] if file_names else "img_%06d.jpg" % i) + "\n")
            f.write(str(class_names[i]) + "\n")
            f.write("".join("{:f} ".format(float(v)) for v in row) + "\n")


def write_video_file(path, people):
    """The reference's video-features text format (qt_cpp/video.cpp:38-86): per person a name line and the number of
    videos; per video its frame count; per frame a file-name line and D floats.  people: {name: [frames x D array, ...]}."""
    with open(path, "w") as f:
        for name, videos in people.items():
            f.write(name + "\n%d\n" % len(videos))
            for v, frames in enumerate(videos):
                f.write("%d\n" % len(frames))
                for j, row in enumerate(frames):
                    f.write("%s_v%d_f%04d.jpg\n" % (name, v, j))
                    f.write("".join("{:f} ".format(float(x)) for x in row) + "\n")


# ---------------------------------------------------------------------------------------------------
# Counter-based generator (csrc/synth_common.h): every element is a pure function of (seed, row, column), so a rank of a
# sharded run can materialise rows [lo, hi) of the SAME gallery on its GPU (fir_synth_rows), and the host can regenerate
# identical bits (libfir_synth_host.so, or the numpy restatement below for small cases).  BASELINE configs 1-4 in
# bench.py are drawn from it.
# ---------------------------------------------------------------------------------------------------
SYNTH_KEY1 = 0x46495253
SYNTH_INV_STD = np.array([0x37ddb3d7], dtype=np.uint32).view(np.float32)[0]
ROLE_GALLERY, ROLE_QUERY, ROLE_LABEL, ROLE_CENTROID = 0, 1, 2, 3


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 on uint32 arrays (same rounds as fir_philox4x32_10)."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & 0xffffffff for c in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0 & 0xffffffff), np.uint64(k1 & 0xffffffff)
    m = np.uint64(0xffffffff)
    for _ in range(10):
        p0, p1 = np.uint64(0xD2511F53) * c0, np.uint64(0xCD9E8D57) * c2
        c0, c1, c2, c3 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & m, p1 & m, ((p0 >> np.uint64(32)) ^ c3 ^ k1) & m, p0 & m
        k0, k1 = (k0 + np.uint64(0x9E3779B9)) & m, (k1 + np.uint64(0xBB67AE85)) & m
    return c0, c1, c2, c3


def _z(seed, rows, d):
    """[len(rows), d] unit-variance values z(seed, row, col) — fir_synth_z2."""
    rows = np.asarray(rows, dtype=np.int64)
    pairs = (d + 1) // 2
    pr = np.broadcast_to(np.arange(pairs, dtype=np.uint64)[None, :], (len(rows), pairs))
    r = np.broadcast_to(rows.astype(np.uint64)[:, None], (len(rows), pairs))
    w = philox4x32_10(pr, r & np.uint64(0xffffffff), r >> np.uint64(32), np.zeros_like(pr), seed, SYNTH_KEY1)
    lo = np.uint64(0xffff)
    s0 = ((w[0] & lo) + (w[0] >> np.uint64(16)) + (w[1] & lo) + (w[1] >> np.uint64(16))).astype(np.int64) - 131070
    s1 = ((w[2] & lo) + (w[2] >> np.uint64(16)) + (w[3] & lo) + (w[3] >> np.uint64(16))).astype(np.int64) - 131070
    z = np.empty((len(rows), 2 * pairs), dtype=np.float32)
    z[:, 0::2] = s0.astype(np.float32) * SYNTH_INV_STD
    z[:, 1::2] = s1.astype(np.float32) * SYNTH_INV_STD
    return z[:, :d]


def synth_labels(role, row_lo, n_rows, n_total, n_classes, seed=0):
    rows = np.arange(row_lo, row_lo + n_rows, dtype=np.int64)
    if role == ROLE_GALLERY:
        return ((rows * n_classes) // n_total).astype(np.int32)
    r = rows.astype(np.uint64)
    w = philox4x32_10(np.zeros_like(r), r & np.uint64(0xffffffff), r >> np.uint64(32), np.ones_like(r), seed + ROLE_LABEL, SYNTH_KEY1)
    return (w[0] % np.uint64(n_classes)).astype(np.int32)


def synth_rows_numpy(role, row_lo, n_rows, n_total, d, n_classes, seed=0, sigma=0.5, relu=False):
    """numpy restatement of fir_synth_rows (small cases; the C twins are the fast paths) → (rows fp32, labels int32)."""
    lab = synth_labels(role, row_lo, n_rows, n_total, n_classes, seed)
    z = _z(seed + role, np.arange(row_lo, row_lo + n_rows, dtype=np.int64), d)
    cen = _z(seed + ROLE_CENTROID, np.arange(n_classes, dtype=np.int64), d)
    v = cen[lab] + (np.float32(sigma) * z).astype(np.float32)
    if relu:
        np.maximum(v, 0, out=v)
    return v.astype(np.float32), lab


_host_lib = None


def _host():
    global _host_lib
    if _host_lib is None:
        import ctypes as C
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libfir_synth_host.so")
        L = C.CDLL(path)
        L.fir_synth_rows_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                          C.c_uint32, C.c_float, C.c_int32, C.c_int32]
        _host_lib = L
    return _host_lib


def synth_rows_host(role, row_lo, n_rows, n_total, d, n_classes, seed=0, sigma=0.5, relu=False, out=None, threads=None):
    """Host twin (libfir_synth_host.so, threaded) → (rows, labels); `out` may be a preallocated [n_rows, d] fp32 array."""
    import os
    rows = out if out is not None else np.empty((n_rows, d), dtype=np.float32)
    assert rows.dtype == np.float32 and rows.flags["C_CONTIGUOUS"] and rows.shape == (n_rows, d)
    lab = np.empty(n_rows, dtype=np.int32)
    rc = _host().fir_synth_rows_host(rows.ctypes.data, lab.ctypes.data, row_lo, n_rows, n_total, d, n_classes, role, seed, sigma, int(relu),
                                     threads or os.cpu_count() or 1)
    if rc != 0:
        raise ValueError("fir_synth_rows_host: bad arguments")
    return rows, lab


def synth_rows_device(role, row_lo, n_rows, n_total, d, n_classes, seed=0, sigma=0.5, relu=False, device="cuda", out=None):
    """fir_synth_rows on the GPU → (rows, labels) CUDA tensors."""
    import ctypes as C
    import torch
    import fir_b200
    rows = out if out is not None else torch.empty((n_rows, d), dtype=torch.float32, device=device)
    lab = torch.empty((n_rows,), dtype=torch.int32, device=rows.device)
    fir_b200._check(fir_b200.lib().fir_synth_rows(C.c_void_p(rows.data_ptr()), C.c_void_p(lab.data_ptr()), row_lo, n_rows, n_total, d, n_classes,
                                                  role, seed, sigma, int(relu), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return rows, lab
