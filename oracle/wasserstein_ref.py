"""TEST INFRASTRUCTURE ONLY — CPU restatement of persim.wasserstein (persim>=0.3,
/root/reference/requirements.txt:6; call site /root/reference/scripts/utils.py:180-191 via the
alias at utils.py:12).  persim is absent from this image; semantics per SURVEY.md Appendix A.2.
parity status: "parity unpinned" (no golden vectors in the reference); the optimum of a linear
assignment problem is unique in value, so scipy's exact LSAP on the persim cost matrix is the
oracle for the CUDA solver.
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import linear_sum_assignment


def cost_matrix(dgm1, dgm2):
    S = np.array(dgm1, dtype=np.float64).reshape(-1, 2)
    S = S[np.isfinite(S[:, 1])]
    T = np.array(dgm2, dtype=np.float64).reshape(-1, 2)
    T = T[np.isfinite(T[:, 1])]
    if len(S) == 0:
        S = np.array([[0.0, 0.0]])
    if len(T) == 0:
        T = np.array([[0.0, 0.0]])
    M, N = len(S), len(T)
    # L2 ground metric via the Gram trick, as sklearn.metrics.pairwise_distances does
    from sklearn.metrics import pairwise_distances
    DUL = pairwise_distances(S, T)
    c = np.cos(np.pi / 4)
    s = np.sin(np.pi / 4)
    R = np.array([[c, -s], [s, c]])
    Sr = S @ R
    Tr = T @ R
    D = np.zeros((M + N, M + N))
    np.fill_diagonal(D, 0)
    D[0:M, 0:N] = DUL
    UR = np.inf * np.ones((M, M))
    np.fill_diagonal(UR, Sr[:, 1])
    D[0:M, N:N + M] = UR
    UL = np.inf * np.ones((N, N))
    np.fill_diagonal(UL, Tr[:, 1])
    D[M:N + M, 0:N] = UL
    return D, M, N


def wasserstein(dgm1, dgm2, matching=False):
    D, M, N = cost_matrix(dgm1, dgm2)
    ri, ci = linear_sum_assignment(D)
    total = float(np.sum(D[ri, ci]))
    if matching:
        return total, np.c_[ri, ci]
    return total


def safe_wasserstein(d1, d2):
    """utils.safe_wasserstein (/root/reference/scripts/utils.py:180-191)."""
    def clean(d):
        d = np.asarray(d)
        if d.ndim != 2 or d.shape[0] == 0:
            return np.array([[0, 0]])
        d = d[np.isfinite(d).all(axis=1)]
        return d if len(d) > 0 else np.array([[0, 0]])
    try:
        return wasserstein(clean(d1), clean(d2))
    except Exception:
        return np.nan
