"""TEST INFRASTRUCTURE ONLY — ctypes front-end of oracle/pcoh_large_model.cpp, the executable CPU
model of the algorithm tda_eeg_audio_b200/csrc/rips_large.cu runs (classification of every edge
as merging / apparent / birth, then a cocycle sweep over the edges a live class can see).
Checked against oracle/rips_cpu.cpp in tests/test_oracle_rips.py; never used by the product."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libmodel_large.so")
        src = os.path.join(_HERE, "pcoh_large_model.cpp")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "-B", "libmodel_large.so"], stdout=subprocess.DEVNULL)
        _LIB = ctypes.CDLL(so)
        _LIB.model_large_rips_h01.restype = ctypes.c_int
    return _LIB


STAT_NAMES = ("edges", "visited_runs", "visited_edges", "births", "deaths", "absorbs", "max_live", "loops",
              "scrubs", "groups_evaluated")


def rips_h01(D, thresh=np.inf, cap1=None):
    """D (n, n) float32 (upper triangle read) -> dict like oracle.rips.ripser plus "stats"."""
    D = np.ascontiguousarray(D, dtype=np.float32)
    n = D.shape[0]
    if cap1 is None:
        cap1 = max(n * (n - 1) // 2, 1)
    bd0 = np.zeros((n, 2), np.float32); pr0 = np.full((n, 2), -1, np.int64)
    bd1 = np.zeros((cap1, 2), np.float32); pr1 = np.full((cap1, 2), -1, np.int64)
    counts = np.zeros(2, np.int32); stats = np.zeros(16, np.int64)
    vp = ctypes.c_void_p
    rc = lib().model_large_rips_h01(vp(D.ctypes.data), n, n, ctypes.c_float(float(np.float32(thresh))),
                                    vp(bd0.ctypes.data), vp(pr0.ctypes.data), vp(bd1.ctypes.data),
                                    vp(pr1.ctypes.data), cap1, vp(counts.ctypes.data), vp(stats.ctypes.data))
    if rc != 0:
        raise RuntimeError(f"model_large_rips_h01 rc={rc}")
    n0, n1 = (int(x) for x in counts)
    n1 = min(n1, cap1)
    return {"dgms": [bd0[:n0].astype(np.float64), bd1[:n1].astype(np.float64)], "pairs": [pr0[:n0].copy(), pr1[:n1].copy()],
            "stats": dict(zip(STAT_NAMES, (int(x) for x in stats)))}
