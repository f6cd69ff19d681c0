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


def tiny_dataset(root, seed=5, seconds=8):
    """A one-recording data set in the reference's on-disk layout (data/<cond>/<rec>.mat,
    graphs/<cond>/<rec>/<band>_distances.npy), written with seeded numpy + the scipy restatement of
    notebooks 1-2 (oracle.signal_ref).  Used both by tests/golden/make_golden.py (which runs the
    reference's own script functions on it) and by the GPU driver tests (which must see the very
    same files)."""
    from pathlib import Path
    from scipy.io import savemat
    from oracle import signal_ref
    root = Path(root)
    rng = np.random.default_rng(seed)
    n = 250 * seconds
    A = rng.standard_normal((66, 8)) / np.sqrt(8)
    sub = (A @ rng.standard_normal((8, n)) + 0.5 * rng.standard_normal((66, n))).T     # (samples, electrodes)
    t = np.arange(44100 * seconds) / 44100.0
    y = (1 + 0.6 * np.sin(2 * np.pi * 3.1 * t + 0.4) + 0.3 * np.sin(2 * np.pi * 6.7 * t + 1.1)) * \
        rng.standard_normal(len(t))
    y = np.stack([y, y + 0.01 * rng.standard_normal(len(t))], axis=1)
    (root / "data" / "slow").mkdir(parents=True, exist_ok=True)
    mat = root / "data" / "slow" / "S01_trial1.mat"
    savemat(mat, {"subeeg": sub, "y": y, "Fs": np.array([[44100]])})
    good = [x - 1 for x in [2, 3, 4, 6, 7, 9, 11, 12, 13, 14, 15, 16, 18, 19, 20, 21, 22, 24, 25, 26, 27, 28, 30, 31, 33,
                            34, 36, 38, 40, 41, 42, 44, 45, 46, 48, 49, 50, 51, 52, 53, 54, 56, 57, 58, 59, 60, 65]]
    eeg = sub.T[good]
    gdir = root / "graphs" / "slow" / "S01_trial1"
    gdir.mkdir(parents=True, exist_ok=True)
    for band, (lo, hi) in signal_ref.FREQ_BANDS.items():
        filt = signal_ref.apply_bandpass_filter(eeg, lo, hi, 250)
        wins, _ = signal_ref.create_sliding_windows(filt, 1.0, 0.75, 250)
        corr = np.stack([signal_ref.compute_correlation_matrix(w) for w in wins])
        dist = np.stack([signal_ref.correlation_to_distance(c) for c in corr])
        np.save(gdir / f"{band}_correlations.npy", corr)
        np.save(gdir / f"{band}_distances.npy", dist)
    return mat, gdir
