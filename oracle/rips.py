"""TEST INFRASTRUCTURE ONLY — ctypes front-end of the CPU oracle (oracle/rips_cpu.cpp).

Mirrors the third-party `ripser.ripser(X, maxdim, thresh, coeff, distance_matrix)` call the
reference makes (/root/reference/scripts/utils.py:131,140;
/root/reference/scripts/tda_eeg_classification_v2.py:170-175) per SURVEY.md Appendix A.1.
parity status: "parity unpinned" (no golden vectors exist in the reference; pinned against
oracle/rips_naive.py instead).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle_rips.so")
    src = os.path.join(_HERE, "rips_cpu.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle_rips.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.oracle_rips_h01_batched.restype = ctypes.c_int
        _LIB.oracle_rips_h01_batched.argtypes = [
            ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        _LIB.oracle_max_threads.restype = ctypes.c_int
    return _LIB


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def rips_h01_batched(D, thresh=np.inf, cap1=None, nthreads=0, return_stats=False):
    """D: (B, n, n) float32 (upper triangle read).  Returns padded arrays like the C-ABI."""
    D = np.ascontiguousarray(D, dtype=np.float32)
    assert D.ndim == 3 and D.shape[1] == D.shape[2]
    B, n, _ = D.shape
    if cap1 is None:
        cap1 = max(n * (n - 1) // 2 - (n - 1), 1)
    bd0 = np.zeros((B, n, 2), np.float32)
    pr0 = np.full((B, n, 2), -1, np.int64)
    bd1 = np.zeros((B, cap1, 2), np.float32)
    pr1 = np.full((B, cap1, 2), -1, np.int64)
    counts = np.zeros((B, 2), np.int32)
    status = np.zeros((B,), np.int32)
    stats = np.zeros(3, np.int64)
    thr = np.float32(thresh)
    rc = lib().oracle_rips_h01_batched(
        D.ctypes.data, B, n, n, ctypes.c_float(float(thr)), bd0.ctypes.data, pr0.ctypes.data,
        bd1.ctypes.data, pr1.ctypes.data, counts.ctypes.data, cap1, status.ctypes.data, nthreads,
        stats.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"oracle_rips_h01_batched rc={rc}")
    out = dict(bd0=bd0, pr0=pr0, bd1=bd1, pr1=pr1, counts=counts, status=status)
    if return_stats:
        out["stats"] = dict(columns=int(stats[0]), emergent=int(stats[1]), additions=int(stats[2]))
    return out


def pairwise_f32(X):
    """What ripser.py feeds its C++ core for a point cloud: sklearn pairwise_distances (f64
    Gram-trick Euclidean) cast to float32."""
    from sklearn.metrics import pairwise_distances
    return pairwise_distances(np.asarray(X), metric="euclidean").astype(np.float32)


def ripser(X, maxdim=1, thresh=np.inf, coeff=2, distance_matrix=False, **_):
    """Drop-in shaped like ripser.ripser for maxdim<=1, coeff=2.  Adds "pairs"."""
    if coeff != 2 or maxdim > 1:
        raise NotImplementedError("oracle covers maxdim<=1, coeff=2 (all the reference uses)")
    X = np.asarray(X)
    if distance_matrix:
        if X.shape[0] != X.shape[1]:
            raise Exception("Distance matrix is not square")
        dm = X.astype(np.float32)
    else:
        dm = pairwise_f32(X)
    r = rips_h01_batched(dm[None], thresh=thresh)
    n0, n1 = r["counts"][0]
    dg0 = r["bd0"][0, :n0].astype(np.float64)
    dg1 = r["bd1"][0, :n1].astype(np.float64)
    dgms = [dg0.reshape(-1, 2)] + ([dg1.reshape(-1, 2)] if maxdim >= 1 else [])
    pairs = [r["pr0"][0, :n0].copy()] + ([r["pr1"][0, :n1].copy()] if maxdim >= 1 else [])
    return {"dgms": dgms, "pairs": pairs, "num_edges": None, "cocycles": [[], []],
            "dperm2all": dm, "idx_perm": np.arange(dm.shape[0]), "r_cover": 0.0}
