"""Audio side before Rips: delay tau, Takens embedding (+ min-max normalisation), pairwise
distances.  Mirrors compute_tau / takens_embedding / compute_audio_persistence's normalisation
(/root/reference/scripts/utils.py:92-130) and the sklearn pairwise_distances call inside
ripser(point_cloud)."""
from __future__ import annotations

import numpy as np

from . import _lib


def _stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


def compute_tau_batched(wins, max_lag=None):
    """wins: CUDA float64 (B, L) (row stride free) -> CUDA int32 (B,)"""
    import torch
    _lib.require_cuda()
    assert wins.is_cuda and wins.dtype == torch.float64 and wins.dim() == 2 and wins.stride(1) == 1
    B, L = wins.shape
    tau = torch.empty((B,), dtype=torch.int32, device=wins.device)
    with torch.cuda.device(wins.device):
        rc = _lib.load().tda_compute_tau(wins.data_ptr(), B, L, wins.stride(0) if B > 1 else L,
                                         -1 if max_lag is None else int(max_lag), tau.data_ptr(), _stream())
    _lib.check(rc, "tda_compute_tau")
    return tau


def takens_cloud_batched(wins, tau, dim=3, subsample=1, normalise=True, ldp=None):
    """wins (B, L) CUDA float64; tau: CUDA int32 (B,) or a python int shared by all windows.
    Returns pts (B, ldp, dim) float64 (rows >= npts[b] undefined) and npts (B,) int32."""
    import torch
    _lib.require_cuda()
    B, L = wins.shape
    if isinstance(tau, int):
        tau_t = torch.tensor([tau], dtype=torch.int32, device=wins.device)
        tstride = 0
    else:
        tau_t = tau.to(torch.int32).contiguous()
        tstride = 1
    if ldp is None:
        ldp = (L + subsample - 1) // subsample
    pts = torch.zeros((B, ldp, dim), dtype=torch.float64, device=wins.device)
    npts = torch.empty((B,), dtype=torch.int32, device=wins.device)
    with torch.cuda.device(wins.device):
        rc = _lib.load().tda_takens_cloud(wins.data_ptr(), B, L, wins.stride(0) if B > 1 else L, tau_t.data_ptr(),
                                          tstride, dim, subsample, 1 if normalise else 0, ldp, pts.data_ptr(),
                                          npts.data_ptr(), _stream())
    _lib.check(rc, "tda_takens_cloud")
    return pts, npts


def pairwise_distance_f32(pts, npts=None, ld=None):
    """pts (B, n, dim) CUDA float64 -> D (B, ld, ld) float32 with sklearn's Gram-trick arithmetic."""
    import torch
    _lib.require_cuda()
    assert pts.is_cuda and pts.dtype == torch.float64 and pts.dim() == 3
    pts = pts.contiguous()
    B, n, dim = pts.shape
    if ld is None:
        ld = n
    D = torch.zeros((B, ld, ld), dtype=torch.float32, device=pts.device)
    with torch.cuda.device(pts.device):
        rc = _lib.load().tda_pairwise_dist_f32(pts.data_ptr(), None if npts is None else npts.data_ptr(), B, n, dim,
                                               ld, D.data_ptr(), _stream())
    _lib.check(rc, "tda_pairwise_dist_f32")
    return D


# ----------------------------------------------------------------------------- drop-ins
def compute_tau(s, max_lag=None):
    import torch
    w = torch.from_numpy(np.ascontiguousarray(s, dtype=np.float64)).cuda()[None]
    return int(compute_tau_batched(w, max_lag)[0].item())


def takens_embedding(s, dim, tau, subsample=1):
    import torch
    s = np.ascontiguousarray(s, dtype=np.float64)
    n = len(s) - (dim - 1) * tau
    if n <= 0:
        return np.array([]).reshape(0, dim)
    w = torch.from_numpy(s).cuda()[None]
    pts, npts = takens_cloud_batched(w, int(tau), dim, subsample, normalise=False)
    return pts[0, : int(npts[0].item())].cpu().numpy()
