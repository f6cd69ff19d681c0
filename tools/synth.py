"""Synthetic workloads of the benches and parity tests (setup only, never timed).

raw_eeg / raw_eeg_batch are BASELINE.md §5 (a)/(b): recording `rec` is
    rng = np.random.default_rng(20261018 + rec)
    A = rng.standard_normal((47, 8)) / sqrt(8); S = rng.standard_normal((8, T)); E = rng.standard_normal((47, T))
    x = A @ S + 0.5 * E                                   (float64, 47 channels x T = 15,000 samples at 250 Hz)
one mixing matrix per recording; the band-pass / windows / correlation distance that follow are the
repo's own chain (tda_eeg_audio_b200.dsp.eeg_distances_from_raw) or, on the CPU, the oracle's
(oracle.signal_ref.eeg_distances).

eeg_like_distance_matrices is the quick stand-in of round 1 (a new mixing matrix per WINDOW, no
band-pass): kept for smoke tests and profiling workloads that need matrices without a filter pass."""
from __future__ import annotations

import numpy as np

SEED0 = 20261018
N_CH, N_SRC, T_EEG = 47, 8, 15000


def raw_eeg(rec: int, T: int = T_EEG, seed0: int = SEED0) -> np.ndarray:
    rng = np.random.default_rng(seed0 + rec)
    A = rng.standard_normal((N_CH, N_SRC)) / np.sqrt(N_SRC)
    S = rng.standard_normal((N_SRC, T))
    E = rng.standard_normal((N_CH, T))
    return A @ S + 0.5 * E


def raw_eeg_batch(rec0: int, n: int, T: int = T_EEG, out=None, threads: int = 0):
    """(n, 47, T) float64 host array of recordings rec0 .. rec0+n-1 (numpy's generators release the GIL)."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    if out is None:
        out = np.empty((n, N_CH, T), dtype=np.float64)
    threads = threads or min(32, os.cpu_count() or 1)

    def one(k):
        out[k] = raw_eeg(rec0 + k, T)

    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(one, range(n)))
    return out


def raw_eeg_to_device(rec0: int, n: int, device, T: int = T_EEG, chunk: int = 64):
    """(n, 47, T) float64 CUDA tensor of the stated generator, staged through pinned chunks."""
    import torch
    x = torch.empty((n, N_CH, T), dtype=torch.float64, device=device)
    stage = torch.empty((min(chunk, n), N_CH, T), dtype=torch.float64, pin_memory=True)
    for r0 in range(0, n, chunk):
        m = min(chunk, n - r0)
        raw_eeg_batch(rec0 + r0, m, T, out=stage[:m].numpy())
        x[r0:r0 + m].copy_(stage[:m], non_blocking=False)
    return x


def eeg_distance_matrices(rec0: int, n: int, device, step: int = 250, rec_chunk: int = 256, keep_raw: bool = False):
    """BASELINE §5(b) input: (n, 5, W, 47, 47) float32 correlation-distance matrices of recordings
    rec0 .. rec0+n-1 (5 bands, 1 s windows, W = 60 at step 250), built by the repo's own signal chain
    on the device.  Returns (D, x) with x the raw recordings when keep_raw."""
    import torch
    from tda_eeg_audio_b200 import dsp
    overlap = 1.0 - step / 250.0
    W = dsp.n_windows(T_EEG, 250, int(250 * (1 - overlap)))
    D = torch.empty((n, 5, W, N_CH, N_CH), dtype=torch.float32, device=device)
    xall = torch.empty((n, N_CH, T_EEG), dtype=torch.float64, device=device) if keep_raw else None
    for r0 in range(0, n, rec_chunk):
        m = min(rec_chunk, n - r0)
        x = raw_eeg_to_device(rec0 + r0, m, device)
        dsp.eeg_distances_from_raw(x, overlap=overlap, rec_chunk=m, out=D[r0:r0 + m])
        if keep_raw:
            xall[r0:r0 + m] = x
    return D, xall


def eeg_like_distance_matrices(B, n=47, win=250, k=8, noise=0.5, seed=SEED0, device="cuda", chunk=8192):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((B, n, n), dtype=torch.float32, device=device)
    for b0 in range(0, B, chunk):
        nb = min(chunk, B - b0)
        A = torch.randn((nb, n, k), generator=g, device=device, dtype=torch.float64) / k ** 0.5
        S = torch.randn((nb, k, win), generator=g, device=device, dtype=torch.float64)
        E = torch.randn((nb, n, win), generator=g, device=device, dtype=torch.float64)
        x = A @ S + noise * E
        x = x - x.mean(dim=2, keepdim=True)
        c = x @ x.transpose(1, 2) / (win - 1)
        sd = torch.sqrt(torch.diagonal(c, dim1=1, dim2=2))
        r = (c / sd[:, :, None] / sd[:, None, :]).clamp_(-1, 1)
        d = torch.sqrt(2 * (1 - r)).clamp_min_(0)
        d.diagonal(dim1=1, dim2=2).zero_()
        out[b0:b0 + nb] = d.float()
    return out
