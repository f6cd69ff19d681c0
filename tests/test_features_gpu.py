"""GPU parity of the feature kernels: against the numpy restatement of the reference
(oracle/features_ref.py, itself pinned to the reference's own extract_features by
tests/golden/features.npz) — tolerance 1e-9 relative (north_star allows 1e-5)."""
import numpy as np
import pytest

from tests import inputs

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-9, 1e-12


def test_features_and_table_match_oracle(cuda):
    import torch
    from oracle import features_ref, rips as orips
    from tda_eeg_audio_b200 import pipeline
    R, Bd, Wn, N = 3, 5, 7, 47
    D = inputs.eeg_like(np.random.default_rng(21), R * Bd * Wn)
    res = pipeline.eeg_features_from_distances(torch.from_numpy(D).cuda().view(R, Bd, Wn, N, N), cap1=128)
    c = orips.rips_h01_batched(D, 2.0, cap1=128)
    feats = np.zeros((len(D), 2, 11))
    for b in range(len(D)):
        n0, n1 = c["counts"][b]
        feats[b, 0] = features_ref.extract_features_vec(c["bd0"][b, :n0])
        feats[b, 1] = features_ref.extract_features_vec(c["bd1"][b, :n1])
    np.testing.assert_allclose(res["feats"].cpu().numpy().reshape(-1, 2, 11), feats, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(res["table"].cpu().numpy(),
                               features_ref.aggregate_windows(feats.reshape(R, Bd, Wn, 2, 11)), rtol=RTOL, atol=ATOL)


def test_extract_features_dropin_edge_cases(cuda):
    from oracle import features_ref
    from tda_eeg_audio_b200.features import extract_features, FEATURE_NAMES
    cases = [
        np.zeros((0, 2)),                                   # empty diagram
        np.array([[0.0, np.inf]]),                          # only an essential class
        np.array([[0.1, 0.4]]),                             # single finite bar: std := 0, entropy := 0
        np.array([[0.0, 0.0], [0.0, 0.0]]),                 # zero total persistence
        np.array([[0.0, 0.5], [0.0, 0.5], [0.0, np.inf]]),
        np.array([[0.2, 0.9], [0.1, 0.3], [0.25, 0.26], [0.3, np.inf], [0.05, 0.8]]),
    ]
    for d in cases:
        got = extract_features(d)
        want = features_ref.extract_features_vec(d)
        assert list(got.keys()) == FEATURE_NAMES
        assert isinstance(got["n_features"], int) and isinstance(got["mean_birth"], float)
        # the reference holds float32 values in float64 arrays; feed the same float32 values
        want32 = features_ref.extract_features_vec(d.astype(np.float32).astype(np.float64))
        np.testing.assert_allclose(np.array(list(got.values()), float), want32, rtol=RTOL, atol=ATOL)
        assert want.shape == (11,)


def test_golden_features(cuda):
    """Fixtures produced by the reference's own extract_features (tests/golden/make_golden.py)."""
    import os
    from tda_eeg_audio_b200.features import extract_features
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "features.npz"))
    for k in range(int(g["n"])):
        d = g[f"dgm{k}"]
        got = np.array(list(extract_features(d).values()), float)
        np.testing.assert_allclose(got, g[f"feat{k}"], rtol=1e-6, atol=1e-9)


def test_host_e2e_entry(cuda):
    import torch
    from tda_eeg_audio_b200 import _lib, pipeline
    R, Bd, Wn, N, cap1 = 2, 5, 6, 47, 128
    D = inputs.eeg_like(np.random.default_rng(22), R * Bd * Wn)
    table = np.zeros((R, Bd * 44)); feats = np.zeros((R * Bd * Wn, 2, 11))
    counts = np.zeros((R * Bd * Wn, 2), np.int32)
    rc = _lib.load().tda_eeg_features_host(D.ctypes.data, R, Bd, Wn, N, 2.0, cap1, None, None, counts.ctypes.data,
                                           None, feats.ctypes.data, table.ctypes.data, 0)
    assert rc == 0
    res = pipeline.eeg_features_from_distances(torch.from_numpy(D).cuda().view(R, Bd, Wn, N, N), cap1=cap1)
    assert np.array_equal(table, res["table"].cpu().numpy())
    assert np.array_equal(feats, res["feats"].cpu().numpy().reshape(-1, 2, 11))
    assert np.array_equal(counts, res["rips"]["counts"].cpu().numpy())
