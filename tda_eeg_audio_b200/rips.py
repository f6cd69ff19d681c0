"""Rips H0+H1 entry points.

`ripser(...)` mirrors the third-party call the reference makes
(/root/reference/scripts/utils.py:131,140; tda_eeg_classification_v2.py:170-175):
`ripser(X, maxdim=1, thresh=..., distance_matrix=...)["dgms"]`.  `rips_h01_batched` is the
batched tensor form (the fast path): one call for hundreds of thousands of windows.
"""
from __future__ import annotations

import numpy as np

from . import _lib


def _ptr(t):
    return None if t is None else t.data_ptr()


def condense(D):
    """(B, N, N) -> (B, N(N-1)/2): the upper triangle in row-major order, i.e. the vector `DParam`
    that ripser.py builds (`dm[I > J]` on a meshgrid) and hands its C++ core.
    Works on torch tensors and numpy arrays."""
    import torch
    n = D.shape[-1]
    if isinstance(D, torch.Tensor):
        iu = torch.triu_indices(n, n, 1, device=D.device)
        return D[..., iu[0], iu[1]].contiguous()
    iu = np.triu_indices(n, 1)
    return np.ascontiguousarray(D[..., iu[0], iu[1]])


def rips_h01_batched(D, thresh=float("inf"), cap1=None, want_pairs=True, out=None, npts=None, engine="auto",
                     n_points=None):
    """D: CUDA float32 tensor (B, N, N) (any row stride; upper triangle is read), 2 <= N <= 2048.

    With `n_points=N` (N <= 64), D is the condensed form (B, N(N-1)/2) of `condense` -- ripser's own
    FFI format, half the bytes.

    engine="auto": N <= 64 runs the warp-per-window engine (rips_small), larger N (or ragged
    batches) the grid-cooperative engine (rips_large).  `npts` (CUDA int32 (B,), large engine)
    gives per-item point counts for padded batches.  Returns dict of CUDA tensors: bd0 (B,N,2) f32, pr0 (B,N,2) i64,
    bd1 (B,cap1,2) f32, pr1 (B,cap1,2) i64, counts (B,2) i32, status (B,) i32 — layout of
    include/tda_b200.h.
    """
    import torch
    _lib.require_cuda()
    lib = _lib.load()
    condensed = n_points is not None
    if condensed:
        N = int(n_points)
        if not (isinstance(D, torch.Tensor) and D.is_cuda and D.dtype == torch.float32 and D.dim() == 2
                and D.shape[1] == N * (N - 1) // 2):
            raise TypeError("condensed D must be a CUDA float32 tensor of shape (B, N(N-1)/2)")
        if N > 64 or npts is not None or engine not in ("auto", "small"):
            raise NotImplementedError("condensed input is served by the N <= 64 engine only")
        B = D.shape[0]
        if D.stride(1) != 1:
            D = D.contiguous()
        engine = "small"
    else:
        if not (isinstance(D, torch.Tensor) and D.is_cuda and D.dtype == torch.float32 and D.dim() == 3):
            raise TypeError("D must be a CUDA float32 tensor of shape (B, N, N)")
        B, N, N2 = D.shape
        if N != N2:
            raise Exception("Distance matrix is not square")
        if D.stride(2) != 1 or D.stride(1) < N:
            D = D.contiguous()
    if engine == "auto":
        engine = "small" if (N <= 64 and npts is None) else "large"
    if cap1 is None:
        cap1 = max(N * (N - 1) // 2 - (N - 1), 1)
    dev = D.device
    if out is None:
        out = {}
    def buf(name, shape, dtype):
        t = out.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, device=dev)
            out[name] = t
        return t
    bd0 = buf("bd0", (B, N, 2), torch.float32)
    bd1 = buf("bd1", (B, cap1, 2), torch.float32)
    pr0 = buf("pr0", (B, N, 2), torch.int64) if want_pairs else None
    pr1 = buf("pr1", (B, cap1, 2), torch.int64) if want_pairs else None
    counts = buf("counts", (B, 2), torch.int32)
    status = buf("status", (B,), torch.int32)
    if engine == "small":
        wsb = int(lib.tda_rips_h01_workspace_bytes(B, N))
    elif engine == "large":
        wsb = int(lib.tda_rips_h01_large_workspace_bytes(B, N))
    else:
        raise ValueError(f"unknown engine {engine!r}")
    if wsb == 0 and B > 0:
        raise _lib.TdaError(f"rips_h01_batched: unsupported size N={N} for engine {engine}")
    ws = buf("ws", (max(wsb, 16),), torch.uint8)
    sB = D.stride(0) if B > 0 else 0
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        if engine == "small":
            rc = lib.tda_rips_h01_batched(
                D.data_ptr(), B, N, 0 if condensed else D.stride(1), sB, float(thresh), bd0.data_ptr(),
                _ptr(pr0), bd1.data_ptr(), _ptr(pr1), counts.data_ptr(), cap1, status.data_ptr(),
                ws.data_ptr(), wsb, stream)
        else:
            if npts is not None:
                assert npts.is_cuda and npts.dtype == torch.int32 and npts.shape == (B,)
                npts = npts.contiguous()
            rc = lib.tda_rips_h01_large(
                D.data_ptr(), _ptr(npts), B, N, D.stride(1), sB, float(thresh), bd0.data_ptr(), _ptr(pr0), N,
                bd1.data_ptr(), _ptr(pr1), cap1, counts.data_ptr(), status.data_ptr(), ws.data_ptr(), wsb, stream)
    _lib.check(rc, "tda_rips_h01_" + engine)
    return out


ST_H1_TRUNCATED, ST_NAN_INPUT, ST_INTERNAL = 1, 2, 4   # include/tda_b200.h: TDA_ST_*


def rips_h01_checked(D, thresh=float("inf"), cap1=None, out=None, **kw):
    """rips_h01_batched, then the per-item status reduced on the host (one device->host sync):

      * TDA_ST_INTERNAL (an engine capacity exhausted; the item's diagrams are NOT valid) raises TdaError;
      * TDA_ST_H1_TRUNCATED (more H1 bars than `cap1` rows): the batch is run again with cap1 = the
        largest true count -- ripser never truncates, so neither do the drop-ins and drivers;
      * TDA_ST_NAN_INPUT (NaN distances: those edges are absent, ripser's own comparison semantics)
        warns once per call.
    Returns the same dict (with the enlarged bd1 / pr1 after a re-run)."""
    import torch
    import warnings
    out = rips_h01_batched(D, thresh=thresh, cap1=cap1, out=out, **kw)
    st = out["status"]
    if st.numel() == 0 or not bool((st != 0).any().item()):
        return out
    flags = 0
    for v in torch.unique(st).tolist():
        flags |= int(v)
    if flags & ST_INTERNAL:
        bad = torch.nonzero(st & ST_INTERNAL).flatten()[:8].tolist()
        raise _lib.TdaError(f"rips_h01_batched: internal capacity exhausted (TDA_ST_INTERNAL) for items {bad}: "
                            "more simultaneously alive H1 classes than the last tier holds")
    if flags & ST_NAN_INPUT:
        warnings.warn("rips_h01_batched: NaN distances in the input; those edges were left out "
                      "(TDA_ST_NAN_INPUT)", RuntimeWarning, stacklevel=2)
    if flags & ST_H1_TRUNCATED:
        need = int(out["counts"][:, 1].max().item())
        out = rips_h01_batched(D, thresh=thresh, cap1=need, out=out, **kw)
        assert not bool((out["status"] & ST_H1_TRUNCATED).any().item())
    return out


def tier_counts(out, n_points):
    """How many windows of the last rips_h01_batched call on `out` (N <= 64 engine) each capacity
    tier finished: dict W -> count for the 1 / 2 / 4 / 64-word tiers (the one-word tier exists for
    47-point windows only).  Reads the hand-over counters the tiers leave at the head of the
    workspace (csrc/rips_small.cu: counters[2] -> two-word tier, [0] -> four-word, [1] -> 64-word);
    one device->host sync."""
    B = int(out["counts"].shape[0])
    c = out["ws"][:12].view(__import__("torch").int32).tolist()
    to_w2, to_w4, to_w64 = (c[2] if n_points == 47 else B), c[0], c[1]
    return {1: B - to_w2, 2: to_w2 - to_w4, 4: to_w4 - to_w64, 64: to_w64}


def ripser(X, maxdim=1, thresh=np.inf, coeff=2, distance_matrix=False, do_cocycles=False,
           metric="euclidean", n_perm=None):
    """Single-matrix drop-in for ripser.ripser (maxdim<=1, coeff=2 — all the reference uses).

    numpy in, numpy out: {"dgms": [H0 (k,2) float64, H1 (k,2) float64], "pairs": [...]}."""
    import torch
    _lib.require_cuda()
    if coeff != 2 or maxdim > 1 or do_cocycles or n_perm is not None or metric != "euclidean":
        raise NotImplementedError("tda_eeg_audio_b200.ripser supports maxdim<=1, coeff=2, euclidean")
    X = np.asarray(X)
    if distance_matrix:
        if X.ndim != 2 or X.shape[0] != X.shape[1]:
            raise Exception("Distance matrix is not square")
        dm = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32)).cuda()
    else:
        from .takens import pairwise_distance_f32  # device Gram-trick distance, f64 -> f32
        dm = pairwise_distance_f32(torch.from_numpy(np.ascontiguousarray(X, dtype=np.float64)).cuda()[None])[0]
    r = rips_h01_checked(dm[None], thresh=float(thresh))
    n0, n1 = (int(x) for x in r["counts"][0].tolist())
    dg0 = r["bd0"][0, :n0].double().cpu().numpy().reshape(-1, 2)
    dg1 = r["bd1"][0, :n1].double().cpu().numpy().reshape(-1, 2)
    pr0 = r["pr0"][0, :n0].cpu().numpy()
    pr1 = r["pr1"][0, :n1].cpu().numpy()
    dgms = [dg0] + ([dg1] if maxdim >= 1 else [])
    return {"dgms": dgms, "pairs": [pr0] + ([pr1] if maxdim >= 1 else []), "cocycles": [[], []],
            "num_edges": None, "dperm2all": None, "idx_perm": np.arange(dm.shape[0]), "r_cover": 0.0}
