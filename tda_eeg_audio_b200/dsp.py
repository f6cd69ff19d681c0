"""Signal stages of the hot path: zero-phase band-pass filtering, windowing, correlation distance.

Batched CUDA entry points plus drop-ins with the reference's own names and signatures:
  design_bandpass_filter / apply_bandpass_filter   /root/reference/notebooks/1_preprocesamiento.ipynb:209-265
  create_sliding_windows                           /root/reference/notebooks/1_preprocesamiento.ipynb:314-364
  compute_correlation_matrix / correlation_to_distance  /root/reference/notebooks/2_graph_construction.ipynb:86-122
  bandpass_filter, create_windows                  /root/reference/scripts/utils.py:66-89

Filter DESIGN (a few dozen Butterworth coefficients and the zi steady states) is done once on the
host with scipy.signal.butter / sosfilt_zi / lfilter_zi, as the reference does; every sample of
every recording goes through the CUDA recursion (csrc/iir.cu)."""
from __future__ import annotations

import numpy as np

from . import _lib

FREQ_BANDS = {"delta": (0.5, 4), "theta": (4, 8), "alpha": (8, 13), "beta": (13, 30), "gamma": (30, 50)}
_METHODS = {"euclidean": 0, "abs": 1, "standard": 2, "sqrt": 3}


# ----------------------------------------------------------------------------- design (host)
def design_bandpass_filter(lowcut, highcut, fs, order=4):
    from scipy import signal
    nyquist = 0.5 * fs
    return signal.butter(order, [lowcut / nyquist, highcut / nyquist], btype="band", output="sos")


def _sos_padlen(sos):
    ntaps = 2 * sos.shape[0] + 1
    ntaps -= min(int((sos[:, 2] == 0).sum()), int((sos[:, 5] == 0).sum()))
    return 3 * ntaps


def design_bandpass_ba(low, high, fs):
    """utils.bandpass_filter's design; returns None when lo >= hi (signal passes through)."""
    from scipy import signal
    nyq = fs / 2
    lo = max(low / nyq, 0.001)
    hi = min(high / nyq, 0.999)
    if lo >= hi:
        return None
    return signal.butter(4, [lo, hi], btype="band")


# ----------------------------------------------------------------------------- batched CUDA
def sosfiltfilt_batched(x, sos_list, out=None, ws=None):
    """x: CUDA float64 (n_seq, T) (row stride free); sos_list: (n_bands, n_sections, 6) host.
    Returns (n_bands, n_seq, T) CUDA float64 = scipy.signal.sosfiltfilt per band and row."""
    import torch
    from scipy import signal
    _lib.require_cuda()
    sos = np.ascontiguousarray(np.asarray(sos_list, dtype=np.float64))
    assert sos.ndim == 3 and sos.shape[2] == 6
    nb, ns, _ = sos.shape
    pad = {_sos_padlen(s) for s in sos}
    assert len(pad) == 1, "bands with different padlen must be filtered in separate calls"
    zi = np.ascontiguousarray(np.stack([signal.sosfilt_zi(s) for s in sos]))
    return _filtfilt(x, 0, nb, ns, sos, zi, pad.pop(), out, ws)


def filtfilt_batched(x, ba_list, out=None, ws=None):
    """x: CUDA float64 (n_seq, T); ba_list: list of (b, a) with equal lengths (host).
    Returns (n_bands, n_seq, T) = scipy.signal.filtfilt(b, a, row) per band and row."""
    from scipy import signal
    _lib.require_cuda()
    nb = len(ba_list)
    n = max(max(len(b), len(a)) for b, a in ba_list)
    coef = np.zeros((nb, 2, n))
    zi = np.zeros((nb, n - 1))
    for k, (b, a) in enumerate(ba_list):
        assert len(b) == n and len(a) == n
        coef[k, 0], coef[k, 1] = b, a
        zi[k] = signal.lfilter_zi(b, a)
    return _filtfilt(x, 1, nb, n, coef, zi, 3 * n, out, ws)


def _filtfilt(x, form, nb, n, coef, zi, padlen, out, ws):
    import torch
    lib = _lib.load()
    assert x.is_cuda and x.dtype == torch.float64 and x.dim() == 2 and x.stride(1) == 1
    n_seq, T = x.shape
    if T <= padlen:
        raise ValueError(f"The length of the input vector x must be greater than padlen, which is {padlen}.")
    if out is None:
        out = torch.empty((nb, n_seq, T), dtype=torch.float64, device=x.device)
    wsb = int(lib.tda_filtfilt_workspace_bytes(n_seq, nb, T, padlen))
    if ws is None or ws.numel() < wsb:
        ws = torch.empty((max(wsb, 8),), dtype=torch.uint8, device=x.device)
    coef = np.ascontiguousarray(coef, dtype=np.float64)
    zi = np.ascontiguousarray(zi, dtype=np.float64)
    with torch.cuda.device(x.device):
        rc = lib.tda_filtfilt_f64(x.data_ptr(), n_seq, T, x.stride(0), form, nb, n, coef.ctypes.data, zi.ctypes.data,
                                  padlen, out.data_ptr(), ws.data_ptr(), wsb, torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "tda_filtfilt_f64")
    return out


def n_windows(T, win, step):
    return 0 if T < win else (T - win) // step + 1


def corrdist_windows(x, win, step, method="euclidean", out=None, want_corr=False, out_rec_stride=0):
    """x: CUDA float64 (R, C, T) contiguous in (C, T).  Returns D (R, W, C, C) float32
    (and the float64 correlations when want_corr)."""
    import torch
    _lib.require_cuda()
    assert x.is_cuda and x.dtype == torch.float64 and x.dim() == 3 and x.stride(2) == 1 and x.stride(1) == x.shape[2]
    R, C, T = x.shape
    W = n_windows(T, win, step)
    if out is None:
        out = torch.empty((R, W, C, C), dtype=torch.float32, device=x.device)
    corr = torch.empty((R, W, C, C), dtype=torch.float64, device=x.device) if want_corr else None
    if W > 0 and R > 0:
        with torch.cuda.device(x.device):
            rc = _lib.load().tda_corrdist_windows(x.data_ptr(), R, C, T, x.stride(0), win, step, _METHODS[method],
                                                  out.data_ptr(), corr.data_ptr() if want_corr else None,
                                                  out_rec_stride, torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "tda_corrdist_windows")
    return (out, corr) if want_corr else out


def eeg_distances_from_raw(x, fs=250, bands=FREQ_BANDS, window_size=1.0, overlap=0.75, order=4,
                           rec_chunk=None, out=None):
    """Raw EEG (R, C, T) CUDA float64 -> correlation-distance matrices (R, n_bands, W, C, C) float32:
    notebooks 1 + 2 of the reference for a whole dataset (band-pass sos filtfilt, 1 s windows,
    corrcoef, sqrt(2(1-r))).

    rec_chunk: recordings filtered per launch (None: as many as fit half of the free device memory --
    the filtered bands and the padded intermediate take 2 x n_bands x 8 bytes per sample; the filter
    kernels hand out their work dynamically, so one big launch beats several small ones)."""
    import torch
    R, C, T = x.shape
    if rec_chunk is None:
        free, _ = torch.cuda.mem_get_info(x.device)
        per_rec = 2 * len(bands) * C * (T + 64) * 8
        rec_chunk = int(max(1, min(R, (free // 2) // per_rec)))
    win = int(window_size * fs)
    step = int(win * (1 - overlap))
    W = n_windows(T, win, step)
    names = list(bands)
    sos = np.stack([design_bandpass_filter(*bands[b], fs, order) for b in names])
    nb = len(names)
    if out is None:
        out = torch.empty((R, nb, W, C, C), dtype=torch.float32, device=x.device)
    filt = ws = None
    for r0 in range(0, R, rec_chunk):
        rc = min(rec_chunk, R - r0)
        xc = x[r0:r0 + rc].reshape(rc * C, T)
        if filt is None or filt.shape[1] != rc * C:
            filt = torch.empty((nb, rc * C, T), dtype=torch.float64, device=x.device)
            ws = torch.empty((int(_lib.load().tda_filtfilt_workspace_bytes(rc * C, nb, T, _sos_padlen(sos[0]))),),
                             dtype=torch.uint8, device=x.device)
        sosfiltfilt_batched(xc, sos, out=filt, ws=ws)
        for b in range(nb):
            # window (rec, w) of band b lands at out[r0+rec, b, w]
            corrdist_windows(filt[b].view(rc, C, T), win, step, out=out[r0:r0 + rc, b],
                             out_rec_stride=out.stride(0))
    return out


# ----------------------------------------------------------------------------- drop-ins (numpy in/out)
def apply_bandpass_filter(data, lowcut, highcut, fs, order=4):
    import torch
    sos = design_bandpass_filter(lowcut, highcut, fs, order)
    x = torch.from_numpy(np.ascontiguousarray(data, dtype=np.float64)).cuda()
    return sosfiltfilt_batched(x, sos[None])[0].cpu().numpy()


def bandpass_filter(s, fs, low, high):
    import torch
    ba = design_bandpass_ba(low, high, fs)
    if ba is None:
        return s
    x = torch.from_numpy(np.ascontiguousarray(s, dtype=np.float64)).cuda()[None]
    return filtfilt_batched(x, [ba])[0, 0].cpu().numpy()


def create_sliding_windows(data, window_size, overlap, fs):
    data = np.asarray(data)
    n_channels, n_samples = data.shape
    window_samples = int(window_size * fs)
    step_samples = int(window_samples * (1 - overlap))
    nw = (n_samples - window_samples) // step_samples + 1
    if nw < 1:
        print(f"Warning: Recording too short for {window_size}s windows (only {n_samples / fs:.2f}s)")
        return np.array([]), np.array([])
    idx = np.arange(nw)[:, None] * step_samples + np.arange(window_samples)[None, :]
    windows = np.ascontiguousarray(data[:, idx].transpose(1, 0, 2)).astype(np.float64)
    times = (np.arange(nw) * step_samples + window_samples // 2) / fs
    return windows, times


def create_windows(s, win_samples, step_samples):
    s = np.asarray(s)
    nw = n_windows(len(s), win_samples, step_samples)
    if nw == 0:
        return np.array([]).reshape(0, win_samples)
    idx = np.arange(nw)[:, None] * step_samples + np.arange(win_samples)[None, :]
    return s[idx]


def compute_correlation_matrix(window_data):
    import torch
    w = torch.from_numpy(np.ascontiguousarray(window_data, dtype=np.float64)).cuda()[None]
    _, corr = corrdist_windows(w, w.shape[2], w.shape[2], want_corr=True)
    return corr[0, 0].cpu().numpy()


def correlation_to_distance_batched(corr, method="euclidean"):
    """corr: CUDA float64 (W, n, n) -> float64 distances (W, n, n), one correlation_to_distance per matrix."""
    import torch
    if method not in _METHODS:
        raise ValueError(f"Unknown method: {method}")
    _lib.require_cuda()
    corr = corr.contiguous()
    d = torch.empty_like(corr)
    st = torch.cuda.current_stream().cuda_stream
    lib = _lib.load()
    for w in range(corr.shape[0]):
        _lib.check(lib.tda_corr_to_dist_f64(corr[w].data_ptr(), corr.shape[1], _METHODS[method], d[w].data_ptr(), st),
                   "tda_corr_to_dist_f64")
    return d


def correlation_to_distance(corr_matrix, method="euclidean"):
    import torch
    if method not in _METHODS:
        raise ValueError(f"Unknown method: {method}")
    _lib.require_cuda()
    c = torch.from_numpy(np.ascontiguousarray(corr_matrix, dtype=np.float64)).cuda()
    d = torch.empty_like(c)
    rc = _lib.load().tda_corr_to_dist_f64(c.data_ptr(), c.shape[0], _METHODS[method], d.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "tda_corr_to_dist_f64")
    return d.cpu().numpy()
