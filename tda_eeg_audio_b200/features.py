"""Persistence features (the 11 scalars per diagram) and the per-recording window aggregation.

Mirrors extract_features (/root/reference/scripts/utils.py:144-177) ≡ extract_persistence_features
(/root/reference/scripts/tda_eeg_classification_v2.py:179-250) and the aggregation loop
(/root/reference/scripts/tda_eeg_classification_v2.py:429-436)."""
from __future__ import annotations

import numpy as np

from . import _lib

FEATURE_NAMES = ["n_features", "n_essential", "mean_birth", "std_birth", "mean_death", "std_death",
                 "mean_persistence", "std_persistence", "max_persistence", "total_persistence",
                 "persistence_entropy"]


def pers_features_batched(bd, counts, out=None):
    """bd (B, cap, 2) CUDA float32; counts: CUDA int32 1-D view (may be strided, e.g. counts[:, 1]).
    out: optional CUDA float64 view (B, 11) (may be a strided slice of a (B, 2, 11) tensor)."""
    import torch
    _lib.require_cuda()
    lib = _lib.load()
    B, cap, _ = bd.shape
    assert bd.is_cuda and bd.dtype == torch.float32 and bd.is_contiguous()
    assert counts.dtype == torch.int32 and counts.dim() == 1 and counts.shape[0] == B
    if out is None:
        out = torch.empty((B, 11), dtype=torch.float64, device=bd.device)
    assert out.dtype == torch.float64 and out.shape == (B, 11) and out.stride(1) == 1
    with torch.cuda.device(bd.device):
        rc = lib.tda_pers_features(bd.data_ptr(), cap, counts.data_ptr(), counts.stride(0) if B > 1 else 1, B,
                                   out.data_ptr(), out.stride(0) if B > 1 else 11,
                                   torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "tda_pers_features")
    return out


def diagram_features(rips_out, out=None):
    """(B, 2, 11) float64 features of the H0 and H1 diagrams returned by rips_h01_batched."""
    import torch
    B = rips_out["counts"].shape[0]
    if out is None:
        out = torch.empty((B, 2, 11), dtype=torch.float64, device=rips_out["counts"].device)
    pers_features_batched(rips_out["bd0"], rips_out["counts"][:, 0], out[:, 0])
    pers_features_batched(rips_out["bd1"], rips_out["counts"][:, 1], out[:, 1])
    return out


def aggregate_windows(feats, out=None):
    """feats (R, Bd, Wn, 2, 11) CUDA float64 -> (R, Bd*44) in feature_names.txt column order."""
    import torch
    _lib.require_cuda()
    R, Bd, Wn, two, eleven = feats.shape
    assert (two, eleven) == (2, 11) and feats.is_cuda and feats.dtype == torch.float64
    feats = feats.contiguous()
    if out is None:
        out = torch.empty((R, Bd * 44), dtype=torch.float64, device=feats.device)
    with torch.cuda.device(feats.device):
        rc = _lib.load().tda_aggregate_windows(feats.data_ptr(), R, Bd, Wn, out.data_ptr(),
                                               torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "tda_aggregate_windows")
    return out


def extract_features(diagram):
    """Drop-in for utils.extract_features: numpy (k,2) diagram -> dict of 11 python scalars."""
    import torch
    _lib.require_cuda()
    d = np.asarray(diagram, dtype=np.float64).reshape(-1, 2)
    k = d.shape[0]
    bd = torch.from_numpy(d.astype(np.float32)).cuda().reshape(1, max(k, 0), 2)
    if k == 0:
        bd = torch.zeros((1, 1, 2), dtype=torch.float32, device="cuda")
    cnt = torch.tensor([k], dtype=torch.int32, device="cuda")
    v = pers_features_batched(bd.contiguous(), cnt)[0].cpu().numpy()
    out = {}
    for name, x in zip(FEATURE_NAMES, v):
        out[name] = int(x) if name in ("n_features", "n_essential") else float(x)
    return out
