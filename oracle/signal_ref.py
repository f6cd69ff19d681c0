"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's signal stages on top of the same
scipy / numpy calls the reference makes.  The notebook functions cannot be imported (ipynb), so
they are restated with verbatim semantics and their cell positions cited; the utils.py functions
are additionally pinned against the reference's own code through tests/golden/signal.npz
(tests/golden/make_golden.py imports /root/reference/scripts/utils.py where it lies)."""
from __future__ import annotations

import numpy as np
from scipy import signal

FREQ_BANDS = {"delta": (0.5, 4), "theta": (4, 8), "alpha": (8, 13), "beta": (13, 30), "gamma": (30, 50)}


def design_bandpass_filter(lowcut, highcut, fs, order=4):
    """/root/reference/notebooks/1_preprocesamiento.ipynb:209-233"""
    nyquist = 0.5 * fs
    return signal.butter(order, [lowcut / nyquist, highcut / nyquist], btype="band", output="sos")


def apply_bandpass_filter(data, lowcut, highcut, fs, order=4):
    """/root/reference/notebooks/1_preprocesamiento.ipynb:236-265 (per-channel sosfiltfilt)"""
    sos = design_bandpass_filter(lowcut, highcut, fs, order)
    out = np.zeros_like(data)
    for i in range(data.shape[0]):
        out[i, :] = signal.sosfiltfilt(sos, data[i, :])
    return out


def create_sliding_windows(data, window_size, overlap, fs):
    """/root/reference/notebooks/1_preprocesamiento.ipynb:314-364"""
    n_channels, n_samples = data.shape
    window_samples = int(window_size * fs)
    step_samples = int(window_samples * (1 - overlap))
    n_windows = (n_samples - window_samples) // step_samples + 1
    if n_windows < 1:
        return np.array([]), np.array([])
    windows = np.zeros((n_windows, n_channels, window_samples))
    times = np.zeros(n_windows)
    for i in range(n_windows):
        s = i * step_samples
        windows[i] = data[:, s:s + window_samples]
        times[i] = (s + window_samples // 2) / fs
    return windows, times


def compute_correlation_matrix(window_data):
    """/root/reference/notebooks/2_graph_construction.ipynb:86-97"""
    with np.errstate(invalid="ignore", divide="ignore"):
        c = np.corrcoef(window_data)
    return np.nan_to_num(c, nan=0.0)


def correlation_to_distance(corr_matrix, method="euclidean"):
    """/root/reference/notebooks/2_graph_construction.ipynb:100-122"""
    c = np.clip(corr_matrix, -1, 1)
    if method == "euclidean":
        d = np.sqrt(2 * (1 - c))
    elif method == "abs":
        d = 1 - np.abs(c)
    elif method == "standard":
        d = 1 - c
    elif method == "sqrt":
        d = np.sqrt(1 - c ** 2)
    else:
        raise ValueError(f"Unknown method: {method}")
    d = np.maximum(d, 0)
    np.fill_diagonal(d, 0)
    return d


def bandpass_filter(s, fs, low, high):
    """/root/reference/scripts/utils.py:66-74 (ba form + filtfilt)"""
    nyq = fs / 2
    lo = max(low / nyq, 0.001)
    hi = min(high / nyq, 0.999)
    if lo >= hi:
        return s
    b, a = signal.butter(4, [lo, hi], btype="band")
    return signal.filtfilt(b, a, s)


def create_windows(s, win_samples, step_samples):
    """/root/reference/scripts/utils.py:82-89"""
    w = []
    start = 0
    while start + win_samples <= len(s):
        w.append(s[start:start + win_samples])
        start += step_samples
    return np.array(w) if w else np.array([]).reshape(0, win_samples)


def compute_tau(s, max_lag=None):
    """/root/reference/scripts/utils.py:92-104"""
    if max_lag is None:
        max_lag = len(s) // 4
    max_lag = min(max_lag, len(s) - 1)
    sc = s - np.mean(s)
    ac = np.correlate(sc, sc, mode="full")
    ac = ac[len(ac) // 2:]
    ac = ac / (ac[0] + 1e-10)
    for i in range(1, min(max_lag, len(ac))):
        if ac[i] <= 0:
            return max(i, 1)
    return max(max_lag // 10, 1)


def takens_embedding(s, dim, tau, subsample=1):
    """/root/reference/scripts/utils.py:107-116"""
    n = len(s) - (dim - 1) * tau
    if n <= 0:
        return np.array([]).reshape(0, dim)
    idx = np.arange(n)[:, None] + np.arange(dim)[None, :] * tau
    pc = s[idx]
    if subsample > 1:
        pc = pc[::subsample]
    return pc


def normalise_cloud(pc):
    """the min-max step of compute_audio_persistence, /root/reference/scripts/utils.py:127-130"""
    mn = pc.min(axis=0)
    rg = pc.max(axis=0) - mn
    rg[rg == 0] = 1
    return (pc - mn) / rg


def eeg_like_recording(rng, C=47, T=15000, k=8, noise=0.5):
    """SURVEY.md §8(d) config (a) generator."""
    S = rng.standard_normal((k, T))
    A = rng.standard_normal((C, k)) / np.sqrt(k)
    E = rng.standard_normal((C, T))
    return A @ S + noise * E


def eeg_distances(x, fs=250, window_size=1.0, overlap=0.75, bands=FREQ_BANDS):
    """notebook 1 + notebook 2 for one recording: (n_bands, W, C, C) float64 distances."""
    out = []
    for b, (lo, hi) in bands.items():
        y = apply_bandpass_filter(x, lo, hi, fs)
        wins, _ = create_sliding_windows(y, window_size, overlap, fs)
        out.append(np.stack([correlation_to_distance(compute_correlation_matrix(w)) for w in wins]))
    return np.stack(out)


def resample_audio(audio, fs_audio=44100, fs_target=250):
    """/root/reference/scripts/utils.py:77-79"""
    return signal.resample_poly(audio, fs_target, fs_audio)


def compute_envelope(s, fs):
    """/root/reference/scripts/utils.py:56-63"""
    analytic = signal.hilbert(s)
    env = np.abs(analytic)
    nyq = fs / 2
    cutoff = min(50, nyq * 0.9)
    b, a = signal.butter(4, cutoff / nyq, btype="low")
    return signal.filtfilt(b, a, env)
