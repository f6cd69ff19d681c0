"""Seeded synthetic inputs shared by the CPU and GPU tests (no file I/O, no reference access)."""
import numpy as np


def sym_uniform(rng, B, n):
    D = rng.random((B, n, n))
    D = (D + D.transpose(0, 2, 1)) / 2
    for b in range(B):
        np.fill_diagonal(D[b], 0)
    return D.astype(np.float32)


def eeg_like(rng, B, n=47, win=250, k=8, noise=0.5):
    """correlation-distance matrices of low-rank + noise windows: d = sqrt(2(1-r))
    (the arithmetic of /root/reference/notebooks/2_graph_construction.ipynb:86-122)."""
    out = np.empty((B, n, n), np.float32)
    for b in range(B):
        A = rng.standard_normal((n, k)) / np.sqrt(k)
        x = A @ rng.standard_normal((k, win)) + noise * rng.standard_normal((n, win))
        r = np.clip(np.nan_to_num(np.corrcoef(x)), -1, 1)
        d = np.maximum(np.sqrt(2 * (1 - r)), 0)
        np.fill_diagonal(d, 0)
        out[b] = d
    return out


def circle_cloud(rng, B, n, noise=0.05):
    from sklearn.metrics import pairwise_distances
    out = np.empty((B, n, n), np.float32)
    for b in range(B):
        th = rng.random(n) * 2 * np.pi
        X = np.c_[np.cos(th), np.sin(th)] + noise * rng.standard_normal((n, 2))
        out[b] = pairwise_distances(X).astype(np.float32)
    return out
