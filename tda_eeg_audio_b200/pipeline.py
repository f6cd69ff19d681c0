"""Batched drivers of the hot path (what the reference does with Python loops over windows).

eeg_features_from_distances = process_file_features
(/root/reference/scripts/tda_eeg_classification_v2.py:338-442) for a whole dataset at once."""
from __future__ import annotations

from . import features as _features
from .rips import rips_h01_batched, rips_h01_checked


def eeg_features_from_distances(D, thresh=2.0, cap1=128, state=None, want_pairs=False, n_points=None, check=True):
    """D: CUDA float32 (R, Bd, Wn, N, N), or with `n_points=N` the condensed (R, Bd, Wn, N(N-1)/2)
    of rips.condense.  Returns dict with
    table (R, Bd*44) float64, feats (R, Bd, Wn, 2, 11) float64 and the raw diagram tensors.
    `state` (a dict) keeps every buffer alive between calls so a steady-state step allocates nothing.
    `cap1` is the INITIAL row capacity of the H1 output: with `check` (default; one device->host sync
    per call) a batch holding a window with more bars is run again with the capacity it needs (kept in
    `state` for the following calls), and an exhausted engine capacity raises TdaError -- features are
    never computed on truncated diagrams.  check=False leaves the status bits to the caller."""
    import torch
    R, Bd, Wn = D.shape[:3]
    B = R * Bd * Wn
    if state is None:
        state = {}
    cap1 = max(cap1, state.get("cap1", 0))
    run = rips_h01_checked if check else rips_h01_batched
    if n_points is None:
        N = D.shape[3]
        rips = run(D.reshape(B, N, N), thresh=thresh, cap1=cap1, want_pairs=want_pairs,
                   out=state.setdefault("rips", {}))
    else:
        rips = run(D.reshape(B, D.shape[3]), thresh=thresh, cap1=cap1, want_pairs=want_pairs,
                   out=state.setdefault("rips", {}), n_points=n_points)
    state["cap1"] = rips["bd1"].shape[1]
    feats = state.get("feats")
    if feats is None or feats.shape[0] != B:
        feats = state["feats"] = torch.empty((B, 2, 11), dtype=torch.float64, device=D.device)
    _features.diagram_features(rips, out=feats)
    table = _features.aggregate_windows(feats.view(R, Bd, Wn, 2, 11), out=state.get("table"))
    state["table"] = table
    return {"table": table, "feats": feats.view(R, Bd, Wn, 2, 11), "rips": rips, "state": state}


# ------------------------------------------------------------------------------------------------
# audio side and the EEG-audio coupling (process_recording / matched_vs_mismatched)
# ------------------------------------------------------------------------------------------------
def select_windows(n_win, max_windows):
    """tda_eeg_audio_comparison.py:77-80 / matched_vs_mismatched.py:52-55: evenly spaced subset."""
    import numpy as np
    if max_windows is not None and n_win > max_windows:
        return np.linspace(0, n_win - 1, max_windows, dtype=int)
    return np.arange(n_win)


def audio_diagrams_from_envelope(env, fs=250, bands=None, window_sec=1.0, overlap=0.75, takens_dim=3,
                                 subsample=2, max_windows=15, thresh=2.0, cap1=256, want_pairs=False,
                                 window_idx=None):
    """Envelope (R, T) CUDA float64 -> per band Takens/Rips diagrams of the selected windows.

    The audio chain of process_recording (/root/reference/scripts/tda_eeg_audio_comparison.py:57-92)
    ≡ get_audio_diagrams (/root/reference/scripts/matched_vs_mismatched.py:35-63) for R recordings:
    ba band-pass filtfilt, 1 s windows (step int(win*(1-overlap))), evenly spaced window subset,
    ONE tau per (recording, band) from the first selected window (max_lag = win//2), Takens(dim 3,
    subsample), min-max normalisation, Rips H0/H1.
    Returns dict: rips (diagram tensors over B = R*n_bands*n_sel items, item = (rec, band, sel)),
    tau (R, n_bands) int32, idx (selected window indices), npts (B,), shape (R, n_bands, n_sel)."""
    import numpy as np
    import torch
    from . import dsp, takens
    from .rips import rips_h01_checked
    bands = bands or dsp.FREQ_BANDS
    names = list(bands)
    R, T = env.shape
    win = int(window_sec * fs)
    step = int(win * (1 - overlap))
    n_win = dsp.n_windows(T, win, step)
    idx = select_windows(n_win, max_windows) if window_idx is None else np.asarray(window_idx, dtype=np.int64)
    n_sel = len(idx)
    ba = [dsp.design_bandpass_ba(*bands[b], fs) for b in names]
    assert all(x is not None for x in ba), "degenerate band (lo >= hi)"
    filt = dsp.filtfilt_batched(env, ba)                              # (n_bands, R, T)
    starts = torch.from_numpy(idx * step).to(env.device)
    gather = starts[:, None] + torch.arange(win, device=env.device)[None, :]        # (n_sel, win)
    wins = filt[:, :, gather].permute(1, 0, 2, 3).contiguous()       # (R, n_bands, n_sel, win)
    nb = len(names)
    first = wins[:, :, 0, :].reshape(R * nb, win)
    tau = takens.compute_tau_batched(first, max_lag=win // 2)        # one tau per (recording, band)
    tau_items = tau.view(R, nb, 1).expand(R, nb, n_sel).reshape(-1).contiguous()
    flat = wins.view(R * nb * n_sel, win)
    ldp = (win + subsample - 1) // subsample
    pts, npts = takens.takens_cloud_batched(flat, tau_items, takens_dim, subsample, normalise=True, ldp=ldp)
    nmax = int(npts.max().item()) if npts.numel() else 0
    nmax = max(nmax, 2)
    D = takens.pairwise_distance_f32(pts[:, :nmax].contiguous(), npts, ld=nmax)
    rips = rips_h01_checked(D, thresh=thresh, cap1=cap1, want_pairs=want_pairs, npts=npts, engine="auto")
    return {"rips": rips, "tau": tau.view(R, nb), "idx": idx, "npts": npts, "shape": (R, nb, n_sel), "D": D}


def cross_wasserstein(eeg_rips, audio_rips, idx_eeg=None, idx_audio=None):
    """Per-item W_H0 and W_H1 between EEG and audio diagrams (safe_wasserstein semantics).
    idx_* select which EEG / audio item forms pair k (None => same position)."""
    from .wasserstein import wasserstein_batched
    w0 = wasserstein_batched(eeg_rips["bd0"], eeg_rips["counts"][:, 0], audio_rips["bd0"],
                             audio_rips["counts"][:, 0], idx_eeg, idx_audio)
    w1 = wasserstein_batched(eeg_rips["bd1"], eeg_rips["counts"][:, 1], audio_rips["bd1"],
                             audio_rips["counts"][:, 1], idx_eeg, idx_audio)
    return w0, w1


def mismatch_reference_recording(n_rec, n_subjects=45):
    """recording r belongs to subject r % 45, condition (r // 45) % 2;
    its mismatched audio is the subject's FIRST recording of the opposite condition
    (/root/reference/scripts/matched_vs_mismatched.py:117-121).  Returns int array (n_rec,), -1 if
    the subject has no recording in the other condition."""
    import numpy as np
    r = np.arange(n_rec)
    subj, cond = r % n_subjects, (r // n_subjects) % 2
    first = {}
    for k in range(n_rec):
        first.setdefault((int(subj[k]), int(cond[k])), k)
    return np.array([first.get((int(subj[k]), 1 - int(cond[k])), -1) for k in range(n_rec)])
