"""The reference-named drop-ins (tda_eeg_audio_b200.utils) against the reference's own utils.py
fixtures and the CPU oracle; and the sys.modules shims."""
import os

import numpy as np
import pytest

from oracle import signal_ref
from tests import inputs

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "signal.npz"))


def test_constants_and_names():
    from tda_eeg_audio_b200 import utils as u
    assert (u.MAX_DIM, u.MAX_EDGE_LENGTH, u.TAKENS_DIM, u.TAKENS_SUBSAMPLE, u.FS_AUDIO, u.FS_EEG) == \
        (1, 2.0, 3, 2, 44100, 250)
    assert list(u.FREQ_BANDS.items()) == list(signal_ref.FREQ_BANDS.items())
    for name in ("bandpass_filter", "create_windows", "compute_tau", "takens_embedding",
                 "compute_audio_persistence", "compute_eeg_persistence", "extract_features", "safe_wasserstein"):
        assert callable(getattr(u, name))


def test_compute_eeg_persistence(cuda):
    from oracle import rips as orips
    from tda_eeg_audio_b200 import utils as u
    D = inputs.eeg_like(np.random.default_rng(3), 1)[0].astype(np.float64)
    D[3, 7] += 1e-3                                   # asymmetric input: symmetrised in float64 first
    dm = np.maximum((D + D.T) / 2, 0); np.fill_diagonal(dm, 0)
    ref = orips.ripser(dm, maxdim=1, thresh=2.0, distance_matrix=True)["dgms"]
    got = u.compute_eeg_persistence(D)
    assert len(got) == 2 and got[0].dtype == np.float64
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])


def test_compute_audio_persistence(cuda):
    from oracle import rips as orips
    from tda_eeg_audio_b200 import utils as u
    pc = G["takens_tau7_sub2"]
    ref = orips.ripser(signal_ref.normalise_cloud(pc), maxdim=1, thresh=2.0)["dgms"]
    got = u.compute_audio_persistence(pc)
    assert got[0].shape == ref[0].shape and got[1].shape == ref[1].shape
    np.testing.assert_allclose(got[0], ref[0], rtol=1e-5)
    np.testing.assert_allclose(got[1], ref[1], rtol=1e-5)
    small = u.compute_audio_persistence(pc[:2])
    assert np.array_equal(small[0], [[0, 0]]) and np.array_equal(small[1], [[0, 0]])


def test_shims(cuda):
    import sys
    from tda_eeg_audio_b200 import utils as u
    saved = {k: sys.modules.get(k) for k in ("ripser", "persim")}
    try:
        u.install_shims()
        from persim import wasserstein
        from ripser import ripser
        D = inputs.eeg_like(np.random.default_rng(4), 1)[0]
        dg = ripser(D, maxdim=1, thresh=2.0, distance_matrix=True)["dgms"]
        assert dg[0].shape == (47, 2) and wasserstein(dg[1], dg[1]) == 0.0
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
