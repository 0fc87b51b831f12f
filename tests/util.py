"""Shared helpers for the parity tests (test infrastructure)."""
import importlib

import numpy as np

synth = importlib.import_module("fast-image-recognition_b200.synth")


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a.view(np.uint64)


def make_data(port, metric, n, nq, d, n_classes, seed=0, sigma=0.5):
    """Synthetic split, normalised 'as loaded by db_features' with the oracle's loader restatement."""
    g, gl, q, ql = synth.make_split(n, nq, d, n_classes, metric, sigma=sigma, seed=seed)
    return port.normalize_rows(metric, g), gl, port.normalize_rows(metric, q), ql


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
