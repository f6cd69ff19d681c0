"""Audio front end of the hot path: 44.1 kHz audio -> 250 Hz (polyphase FIR, only the kept outputs
are computed) -> Hilbert amplitude envelope -> 50 Hz zero-phase low-pass.  Mirrors
resample_audio / compute_envelope of the reference (/root/reference/scripts/utils.py:56-63,77-79),
i.e. scipy.signal.resample_poly, scipy.signal.hilbert and butter(4)+filtfilt.

Filter DESIGN (the 17,641-tap Kaiser FIR, the Butterworth coefficients) is a host-side constant
computed once with scipy.signal.firwin / butter, as the reference does; every sample goes through
the CUDA kernels (csrc/audio.cu, csrc/iir.cu)."""
from __future__ import annotations

import math

import numpy as np

from . import _lib
from .dsp import filtfilt_batched

_DESIGNS = {}


def _output_len(len_h, in_len, up, down):
    """length of scipy.signal.upfirdn's output"""
    return (((in_len - 1) * up + len_h) - 1) // down + 1


def design_resample(n_in, up, down, window=("kaiser", 5.0)):
    """scipy.signal.resample_poly's filter and alignment for an input of n_in samples:
    (up, down, hpoly (up, qmax) float64, n_pre_remove, n_out).  hpoly[p, q] = h_padded[p + q*up]."""
    from scipy.signal import firwin
    g = math.gcd(int(up), int(down))
    up, down = int(up) // g, int(down) // g
    key = (n_in, up, down, window)
    if key in _DESIGNS:
        return _DESIGNS[key]
    n_out = n_in * up
    n_out = n_out // down + bool(n_out % down)
    max_rate = max(up, down)
    half_len = 10 * max_rate
    h = firwin(2 * half_len + 1, 1.0 / max_rate, window=window).astype(np.float64)
    h *= up
    n_pre_pad = down - half_len % down
    n_post_pad = 0
    n_pre_remove = (half_len + n_pre_pad) // down
    while _output_len(len(h) + n_pre_pad + n_post_pad, n_in, up, down) < n_out + n_pre_remove:
        n_post_pad += 1
    hp = np.concatenate([np.zeros(n_pre_pad), h, np.zeros(n_post_pad)])
    qmax = -(-len(hp) // up)
    hpoly = np.zeros((up, qmax))
    for p in range(up):
        col = hp[p::up]
        hpoly[p, :len(col)] = col
    _DESIGNS[key] = (up, down, hpoly, n_pre_remove, n_out)
    return _DESIGNS[key]


def resample_poly_batched(x, up, down, out=None):
    """x: CUDA float64 (n_seq, n_in) -> (n_seq, n_out) = scipy.signal.resample_poly(row, up, down)."""
    import torch
    _lib.require_cuda()
    assert x.is_cuda and x.dtype == torch.float64 and x.dim() == 2 and x.stride(1) == 1
    n_seq, n_in = x.shape
    up_r, down_r, hpoly, n_pre, n_out = design_resample(n_in, up, down)
    if up_r == down_r == 1:
        return x.clone()
    hd = torch.from_numpy(hpoly).to(x.device)
    if out is None:
        out = torch.empty((n_seq, n_out), dtype=torch.float64, device=x.device)
    for s0 in range(0, n_seq, 65535):
        ns = min(65535, n_seq - s0)
        with torch.cuda.device(x.device):
            rc = _lib.load().tda_resample_poly_f64(
                x[s0:].data_ptr(), ns, n_in, x.stride(0), up_r, down_r, hd.data_ptr(), hpoly.shape[1], n_pre, n_out,
                out[s0:].data_ptr(), out.stride(0), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "tda_resample_poly_f64")
    return out


def hilbert_envelope_batched(x, out=None):
    """x: CUDA float64 (n_seq, T) -> abs(scipy.signal.hilbert(row)) (n_seq, T)."""
    import torch
    _lib.require_cuda()
    lib = _lib.load()
    assert x.is_cuda and x.dtype == torch.float64 and x.dim() == 2 and x.stride(1) == 1
    n_seq, T = x.shape
    if out is None:
        out = torch.empty((n_seq, T), dtype=torch.float64, device=x.device)
    wsb = int(lib.tda_hilbert_envelope_workspace_bytes(n_seq, T))
    ws = torch.empty((max(wsb, 16),), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.tda_hilbert_envelope_f64(x.data_ptr(), n_seq, T, x.stride(0), out.data_ptr(), out.stride(0),
                                          ws.data_ptr(), wsb, torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "tda_hilbert_envelope_f64")
    return out


def compute_envelope_batched(s, fs):
    """utils.compute_envelope for every row of s (CUDA float64 (n_seq, T))."""
    from scipy import signal
    nyq = fs / 2
    cutoff = min(50, nyq * 0.9)
    b, a = signal.butter(4, cutoff / nyq, btype="low")
    env = hilbert_envelope_batched(s)
    return filtfilt_batched(env, [(b, a)])[0]


def audio_envelope_from_raw(audio, fs_audio=44100, fs_target=250):
    """Raw audio (R, n_samples) CUDA float64 -> 250 Hz amplitude envelopes (R, n_out):
    resample_audio followed by compute_envelope, as tda_eeg_audio_comparison.process_recording does
    (/root/reference/scripts/tda_eeg_audio_comparison.py:52-60)."""
    return compute_envelope_batched(resample_poly_batched(audio, fs_target, fs_audio), fs_target)


# ----------------------------------------------------------------------------- drop-ins (numpy in/out)
def resample_audio(audio, fs_audio=44100, fs_target=250):
    import torch
    x = torch.from_numpy(np.ascontiguousarray(audio, dtype=np.float64)).cuda()[None]
    return resample_poly_batched(x, fs_target, fs_audio)[0].cpu().numpy()


def compute_envelope(s, fs):
    import torch
    x = torch.from_numpy(np.ascontiguousarray(s, dtype=np.float64)).cuda()[None]
    return compute_envelope_batched(x, fs)[0].cpu().numpy()


def load_audio(mat_path):
    """utils.load_audio (/root/reference/scripts/utils.py:47-53): mono mean of mat['y'] as float64."""
    import scipy.io as sio
    y = sio.loadmat(str(mat_path))["y"]
    if y.ndim == 2:
        y = y.mean(axis=1)
    return y.astype(np.float64)
