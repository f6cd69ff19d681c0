"""Exact diagram Wasserstein distances.  Mirrors persim.wasserstein as used through
utils.safe_wasserstein (/root/reference/scripts/utils.py:180-191)."""
from __future__ import annotations

import numpy as np

from . import _lib


def wasserstein_batched(bdA, nA, bdB, nB, idxA=None, idxB=None, out=None, limA=None, limB=None):
    """bdA (BA, capA, 2), bdB (BB, capB, 2) CUDA padded diagrams, both float32 (the Rips engines'
    output) or both float64; nA / nB CUDA int32 1-D (strided views such as counts[:, 1] are fine).
    Returns CUDA float64 (K,), K = len(idxA) or BA.

    Capacity (include/tda_b200.h): shared memory is linear in the sizes of the largest pair (about 57 bytes
    per point): up to ~4,000 points per pair, beyond that the call raises TdaError (TDA_E_SIZE)."""
    import torch
    _lib.require_cuda()
    assert bdA.is_cuda and bdB.is_cuda and bdA.dtype == bdB.dtype and bdA.dtype in (torch.float32, torch.float64)
    fn = _lib.load().tda_wasserstein_batched if bdA.dtype == torch.float32 else _lib.load().tda_wasserstein_batched_f64
    bdA, bdB = bdA.contiguous(), bdB.contiguous()
    K = bdA.shape[0] if idxA is None else idxA.shape[0]
    if idxA is not None:
        idxA = idxA.to(torch.int32).contiguous()
    if idxB is not None:
        idxB = idxB.to(torch.int32).contiguous()
        assert idxB.shape[0] == K
    elif idxA is None:
        assert bdB.shape[0] == K
    if out is None:
        out = torch.empty((K,), dtype=torch.float64, device=bdA.device)
    # shared memory is sized by the largest diagrams present, not by the padded capacity
    if limA is None:
        limA = max(int(nA.max().item()), 1) if nA.numel() else 1
    if limB is None:
        limB = max(int(nB.max().item()), 1) if nB.numel() else 1
    limA, limB = min(limA, bdA.shape[1]), min(limB, bdB.shape[1])
    sa = nA.stride(0) if nA.numel() > 1 else 1
    sb = nB.stride(0) if nB.numel() > 1 else 1
    with torch.cuda.device(bdA.device):
        rc = fn(
            bdA.data_ptr(), nA.data_ptr(), sa, bdA.shape[1], limA, bdB.data_ptr(), nB.data_ptr(), sb, bdB.shape[1],
            limB,
            None if idxA is None else idxA.data_ptr(), None if idxB is None else idxB.data_ptr(), K,
            out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "tda_wasserstein_batched")
    return out


def _as_batch(d):
    """one diagram -> (1, k, 2) float64 CUDA tensor + its row count (persim works in float64)"""
    import torch
    d = np.asarray(d, dtype=np.float64)
    if d.ndim != 2 or d.shape[0] == 0:
        d = np.zeros((0, 2))
    elif d.shape[1] != 2:
        raise ValueError(f"a persistence diagram has shape (k, 2), got {d.shape}")
    bd = torch.zeros((1, max(len(d), 1), 2), dtype=torch.float64, device="cuda")
    if len(d):
        bd[0, : len(d)] = torch.from_numpy(np.ascontiguousarray(d)).cuda()
    return bd, torch.tensor([len(d)], dtype=torch.int32, device="cuda")


def wasserstein(dgm1, dgm2, matching=False):
    """Drop-in for persim.wasserstein (matching=False, which is all the reference uses)."""
    if matching:
        raise NotImplementedError("matching=True is not used by the reference path")
    a, na = _as_batch(dgm1)
    b, nb = _as_batch(dgm2)
    return float(wasserstein_batched(a, na, b, nb)[0].item())


def safe_wasserstein(dgm1, dgm2):
    """Drop-in for utils.safe_wasserstein (/root/reference/scripts/utils.py:180-191): non-finite rows
    are dropped, an empty diagram counts as [[0, 0]], and MALFORMED INPUT (what makes persim raise:
    ragged / non-numeric / wrong-shaped diagrams) gives nan.  Failures of the engine itself -- a pair
    beyond its capacity (TdaError TDA_E_SIZE), a CUDA error, a missing library -- are NOT turned into
    nan: persim would have returned a value there, and np.nanmean downstream would hide the hole."""
    try:
        a, na = _as_batch(dgm1)
        b, nb = _as_batch(dgm2)
    except (ValueError, TypeError):
        return np.nan
    return float(wasserstein_batched(a, na, b, nb)[0].item())
