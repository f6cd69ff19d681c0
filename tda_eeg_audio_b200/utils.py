"""Drop-in mirror of the reference's scripts/utils.py (same names, signatures, defaults and return
conventions; /root/reference/scripts/utils.py:24-191) on top of the CUDA engine.

    from tda_eeg_audio_b200.utils import *          # instead of: from utils import *

Single-call drop-ins move one window at a time and exist for API parity; the throughput path is
tda_eeg_audio_b200.pipeline (whole datasets per call).  Out of scope here, as in SURVEY.md §2:
load_audio (.mat I/O) and permute_labels_by_subject (statistics)."""
from __future__ import annotations

import numpy as np

from .dsp import FREQ_BANDS, bandpass_filter, create_windows  # noqa: F401
from .features import extract_features  # noqa: F401
from .rips import ripser
from .takens import compute_tau, takens_embedding  # noqa: F401
from .wasserstein import safe_wasserstein, wasserstein as wasserstein_distance  # noqa: F401

MAX_DIM = 1
MAX_EDGE_LENGTH = 2.0
TAKENS_DIM = 3
TAKENS_SUBSAMPLE = 2
FS_AUDIO = 44100
FS_EEG = 250


def compute_audio_persistence(point_cloud, max_dim=MAX_DIM, max_edge_length=MAX_EDGE_LENGTH):
    """utils.py:123-132 — min-max normalise the cloud, then ripser(pc, maxdim, thresh)."""
    import torch
    from . import takens as _t
    from .rips import rips_h01_checked
    point_cloud = np.asarray(point_cloud, dtype=np.float64)
    if len(point_cloud) < 3:
        return [np.array([[0, 0]]), np.array([[0, 0]])]
    pc = torch.from_numpy(np.ascontiguousarray(point_cloud)).cuda()
    mn = pc.min(dim=0).values
    rg = pc.max(dim=0).values - mn
    rg[rg == 0] = 1
    pcn = ((pc - mn) / rg)[None]
    D = _t.pairwise_distance_f32(pcn)
    r = rips_h01_checked(D, thresh=float(max_edge_length))
    n0, n1 = (int(x) for x in r["counts"][0].tolist())
    dg = [r["bd0"][0, :n0].double().cpu().numpy().reshape(-1, 2)]
    if max_dim >= 1:
        dg.append(r["bd1"][0, :n1].double().cpu().numpy().reshape(-1, 2))
    return dg


def compute_eeg_persistence(dist_matrix, max_dim=MAX_DIM, max_edge_length=MAX_EDGE_LENGTH):
    """utils.py:135-141 — (D + D^T)/2, zero diagonal, clamp at 0 (float64), then
    ripser(dm, maxdim, thresh, distance_matrix=True)."""
    import torch
    from . import _lib
    from .rips import rips_h01_checked
    dm = np.ascontiguousarray(dist_matrix, dtype=np.float64)
    if dm.ndim != 2 or dm.shape[0] != dm.shape[1]:
        raise Exception("Distance matrix is not square")
    n = dm.shape[0]
    d64 = torch.from_numpy(dm).cuda()
    d32 = torch.empty((1, n, n), dtype=torch.float32, device="cuda")
    _lib.check(_lib.load().tda_symmetrize_f64_to_f32(d64.data_ptr(), 1, n, d32.data_ptr(),
                                                    torch.cuda.current_stream().cuda_stream),
               "tda_symmetrize_f64_to_f32")
    r = rips_h01_checked(d32, thresh=float(max_edge_length))
    n0, n1 = (int(x) for x in r["counts"][0].tolist())
    dg = [r["bd0"][0, :n0].double().cpu().numpy().reshape(-1, 2)]
    if max_dim >= 1:
        dg.append(r["bd1"][0, :n1].double().cpu().numpy().reshape(-1, 2))
    return dg


def install_shims():
    """Make `from ripser import ripser` / `from persim import wasserstein` resolve to this engine,
    so the reference's scripts run unmodified."""
    import sys
    import types
    from . import wasserstein as _w
    m_r = types.ModuleType("ripser")
    m_r.ripser = ripser
    m_p = types.ModuleType("persim")
    m_p.wasserstein = _w.wasserstein
    sys.modules["ripser"], sys.modules["persim"] = m_r, m_p
    return m_r, m_p
