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


def small_graph_dataset(root, seed=11):
    """graphs/<slow|fast>/<recording>/<band>_distances.npy for five recordings of different lengths
    (6-9 s: 21 to 33 windows at 75 % overlap) in the reference's layout, written with the scipy
    restatement of notebooks 1-2; one recording lacks the gamma band, one band file holds a slightly
    asymmetric first matrix (so the input checker has something to report).  Input of create_dataset:
    tests/golden/make_golden.py::dataset_golden runs the reference's own function on it."""
    from pathlib import Path
    from oracle import signal_ref
    root = Path(root)
    rng = np.random.default_rng(seed)
    spec = [("slow", "S01_trial1", 6), ("slow", "S02_trial4", 8), ("slow", "S02_trial9", 7),
            ("fast", "S01_trial2", 9), ("fast", "S03_trial1", 6)]
    for cond, name, seconds in spec:
        n = 250 * seconds
        A = rng.standard_normal((47, 8)) / np.sqrt(8)
        eeg = A @ rng.standard_normal((8, n)) + 0.5 * rng.standard_normal((47, n))
        gdir = root / "graphs" / cond / name
        gdir.mkdir(parents=True, exist_ok=True)
        for band, (lo, hi) in signal_ref.FREQ_BANDS.items():
            if name == "S03_trial1" and band == "gamma":
                continue
            filt = signal_ref.apply_bandpass_filter(eeg, lo, hi, 250)
            wins, _ = signal_ref.create_sliding_windows(filt, 1.0, 0.75, 250)
            dist = np.stack([signal_ref.correlation_to_distance(signal_ref.compute_correlation_matrix(w)) for w in wins])
            if name == "S02_trial4" and band == "theta":
                dist[0, 3, 7] += 1e-3
            np.save(gdir / f"{band}_distances.npy", dist)
    return root / "graphs" / "slow", root / "graphs" / "fast"


# ------------------------------------------------------------------------------------------------
# Known answers from theory (no code of ours, no third-party library involved)
# ------------------------------------------------------------------------------------------------
def cycle_metric(n, geometry="graph"):
    """n evenly spaced points on a circle.  geometry="graph": hop distance on the n-cycle
    (integers); "chord": Euclidean chords 2 sin(pi k / n), rounded to float32 per hop count k so
    that equal chords are exactly equal.  By Adamaszek & Adams, "The Vietoris-Rips complexes of a
    circle" (Pacific J. Math. 2017, Thm 7.4 / Cor. 6.7 for finite evenly spaced subsets) the Rips
    complex at the scale of k hops is homotopy equivalent to S^1 while k/n < 1/3 and to a sphere
    of dimension >= 2 or a wedge of 2-spheres from k/n >= 1/3 on, so H1 is ONE bar
    [d(1 hop), d(ceil(n/3) hops)) and nothing else of non-zero persistence; H0 is n-1 bars
    [0, d(1 hop)) and one essential class.  Returns (D float32 (n,n), birth, death)."""
    i = np.arange(n)
    hops = np.abs(i[:, None] - i[None, :])
    hops = np.minimum(hops, n - hops)
    if geometry == "graph":
        val = np.arange(n, dtype=np.float32)
    else:
        val = (2.0 * np.sin(np.pi * np.arange(n) / n)).astype(np.float32)
    D = val[hops]
    k = -(-n // 3)
    return D.astype(np.float32), val[1], val[k]


def known_answer_cases():
    """(name, D float32, thresh, expected H0 finite deaths (sorted), expected H1 bars (list of (b, d)))"""
    s2, s3 = np.float32(np.sqrt(2.0)), np.float32(np.sqrt(3.0))
    cases = []
    # unit square: the 4-cycle closes at 1 and is filled by the diagonals at sqrt 2
    sq = np.array([[0, 1, s2, 1], [1, 0, 1, s2], [s2, 1, 0, 1], [1, s2, 1, 0]], np.float32)
    cases.append(("square", sq, np.inf, [1, 1, 1], [(np.float32(1), s2)]))
    # the same below the diagonals: the cycle never dies
    cases.append(("square_thresh", sq, 1.2, [1, 1, 1], [(np.float32(1), np.float32(np.inf))]))
    # two unit squares 10 apart: two independent cycles; H0 joins the squares at distance 10, where the
    # two facing sides and the two gaps also close a 10 x 1 rectangle, filled by its diagonal sqrt 101
    pts = np.array([[0, 0], [1, 0], [1, 1], [0, 1], [11, 0], [12, 0], [12, 1], [11, 1]], float)
    d2 = np.sqrt(((pts[:, None] - pts[None]) ** 2).sum(-1)).astype(np.float32)
    cases.append(("two_squares", d2, np.inf, [1] * 6 + [10], [(np.float32(1), s2), (np.float32(1), s2),
                                                          (np.float32(10), np.float32(np.sqrt(101.0)))]))
    # points on a line: a tree at every scale, no H1
    line = np.abs(np.arange(9)[:, None] - np.arange(9)[None]).astype(np.float32)
    cases.append(("line", line, np.inf, [1] * 8, []))
    # octahedron (6 points, antipodes at 2, all others at sqrt 2): its Rips complex at sqrt 2 is the
    # 2-sphere, simply connected
    octa = np.full((6, 6), s2, np.float32)
    for a in range(3):
        octa[2 * a, 2 * a + 1] = octa[2 * a + 1, 2 * a] = 2
    np.fill_diagonal(octa, 0)
    cases.append(("octahedron", octa, np.inf, [s2] * 5, []))
    # regular hexagon with unit side: filled by the short diagonals sqrt 3 (the octahedron again)
    D, b, d = cycle_metric(6, "chord")
    assert b == np.float32(1) and d == s3
    cases.append(("hexagon", D, np.inf, [1] * 5, [(b, d)]))
    return cases
