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
