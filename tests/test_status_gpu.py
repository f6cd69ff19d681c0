"""Every per-item status bit of the Rips engines, forced, and what the drop-ins / drivers do with it
(no caller used to read them).  TDA_ST_H1_TRUNCATED -> the batch is re-run with the
capacity it needs (ripser never truncates); TDA_ST_INTERNAL -> TdaError; TDA_ST_NAN_INPUT -> warning.
Also: the host entry points keep their staging state per device ordinal, float64 diagrams reach the
Wasserstein kernel unrounded, and safe_wasserstein returns nan only for malformed input."""
import warnings

import numpy as np
import pytest

from tests import inputs

pytestmark = pytest.mark.gpu


def test_truncation_is_rerun_not_returned(cuda):
    import torch
    from tda_eeg_audio_b200 import pipeline
    from tda_eeg_audio_b200.rips import rips_h01_batched, rips_h01_checked
    D = torch.from_numpy(inputs.eeg_like(np.random.default_rng(1), 120)).cuda()
    raw = rips_h01_batched(D, thresh=2.0, cap1=4, want_pairs=False)
    assert bool((raw["status"] & 1).any())                       # the bit is raised ...
    full = rips_h01_checked(D, thresh=2.0, cap1=4, want_pairs=False)
    assert not bool((full["status"] & 1).any())                  # ... and acted on
    assert full["bd1"].shape[1] == int(full["counts"][:, 1].max())
    a = pipeline.eeg_features_from_distances(D.view(2, 5, 12, 47, 47), cap1=4)
    b = pipeline.eeg_features_from_distances(D.view(2, 5, 12, 47, 47), cap1=256)
    assert torch.equal(a["table"], b["table"]) and torch.equal(a["feats"], b["feats"])
    c = pipeline.eeg_features_from_distances(D.view(2, 5, 12, 47, 47), cap1=4, check=False)
    assert bool((c["rips"]["status"] & 1).any()) and not torch.equal(c["table"], b["table"])


def test_internal_capacity_raises(cuda):
    """a random metric on 500 points holds more than 1,024 H1 classes at once: beyond the last tier"""
    import torch
    from tda_eeg_audio_b200 import _lib
    from tda_eeg_audio_b200.rips import rips_h01_batched, rips_h01_checked
    D = torch.from_numpy(inputs.sym_uniform(np.random.default_rng(2), 1, 500)).cuda()
    raw = rips_h01_batched(D, thresh=2.0, cap1=64, want_pairs=False, engine="large")
    assert int(raw["status"][0]) & 4
    with pytest.raises(_lib.TdaError, match="TDA_ST_INTERNAL"):
        rips_h01_checked(D, thresh=2.0, cap1=64, want_pairs=False, engine="large")


def test_nan_input_warns(cuda):
    from tda_eeg_audio_b200 import ripser
    D = inputs.eeg_like(np.random.default_rng(3), 1)[0].astype(np.float64)
    D[3, 9] = D[9, 3] = np.nan
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        r = ripser(D, maxdim=1, thresh=2.0, distance_matrix=True)
    assert any("NaN" in str(x.message) for x in w) and len(r["dgms"]) == 2


def test_wasserstein_float64_and_safe_semantics(cuda):
    from oracle import wasserstein_ref
    from tda_eeg_audio_b200 import _lib
    from tda_eeg_audio_b200.wasserstein import safe_wasserstein, wasserstein
    rng = np.random.default_rng(4)
    a = np.sort(rng.random((9, 2)), axis=1) + 1e-9 * rng.random((9, 2))      # not representable in float32
    b = np.sort(rng.random((14, 2)), axis=1)
    ref = wasserstein_ref.wasserstein(a, b)
    assert abs(wasserstein(a, b) - ref) <= 1e-12 * max(ref, 1.0)
    assert abs(safe_wasserstein(a, np.vstack([b, [[0.2, np.inf]]])) - ref) <= 1e-12 * max(ref, 1.0)
    assert np.isnan(safe_wasserstein(np.zeros((3, 3)), b))                    # malformed input -> nan
    assert np.isnan(safe_wasserstein([[0.1, 0.2], [0.3]], b))
    big = np.sort(rng.random((2100, 2)), axis=1)                              # beyond the engine's capacity (~4,000 points per pair)
    with pytest.raises(_lib.TdaError):
        safe_wasserstein(big, big)


def test_host_entry_on_two_devices_in_one_process(cuda):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from tda_eeg_audio_b200 import _lib
    lib = _lib.load()
    D = inputs.eeg_like(np.random.default_rng(5), 64)
    outs = []
    for dev in (0, 1, 0):
        bd0 = np.zeros((64, 47, 2), np.float32); bd1 = np.zeros((64, 128, 2), np.float32)
        cnt = np.zeros((64, 2), np.int32); st = np.zeros(64, np.int32)
        rc = lib.tda_rips_h01_host(D.ctypes.data, 64, 47, 2.0, bd0.ctypes.data, None, bd1.ctypes.data, None,
                                   cnt.ctypes.data, 128, st.ctypes.data, dev)
        assert rc == 0
        outs.append((bd0, bd1, cnt))
    for o in outs[1:]:
        assert all(np.array_equal(x, y) for x, y in zip(o, outs[0]))
    torch.cuda.set_device(0)
