"""Batched drivers of the hot path (what the reference does with Python loops over windows).

eeg_features_from_distances = process_file_features
(/root/reference/scripts/tda_eeg_classification_v2.py:338-442) for a whole dataset at once."""
from __future__ import annotations

from . import features as _features
from .rips import rips_h01_batched


def eeg_features_from_distances(D, thresh=2.0, cap1=128, state=None, want_pairs=False):
    """D: CUDA float32 (R, Bd, Wn, N, N).  Returns dict with
    table (R, Bd*44) float64, feats (R, Bd, Wn, 2, 11) float64 and the raw diagram tensors.
    `state` (a dict) keeps every buffer alive between calls so a steady-state step allocates nothing."""
    import torch
    R, Bd, Wn, N, _ = D.shape
    B = R * Bd * Wn
    if state is None:
        state = {}
    rips = rips_h01_batched(D.reshape(B, N, N), thresh=thresh, cap1=cap1, want_pairs=want_pairs,
                            out=state.setdefault("rips", {}))
    feats = state.get("feats")
    if feats is None or feats.shape[0] != B:
        feats = state["feats"] = torch.empty((B, 2, 11), dtype=torch.float64, device=D.device)
    _features.diagram_features(rips, out=feats)
    table = _features.aggregate_windows(feats.view(R, Bd, Wn, 2, 11), out=state.get("table"))
    state["table"] = table
    return {"table": table, "feats": feats.view(R, Bd, Wn, 2, 11), "rips": rips, "state": state}


def check_truncation(result):
    """True if any window had more H1 bars than cap1 (one device->host sync)."""
    return bool((result["rips"]["status"] & 1).any().item())
