"""GPU parity of the audio front end (polyphase resampler, Hilbert envelope + LP50) against the
reference's own outputs (tests/golden/audio.npz) and the scipy restatement at full size."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

A = np.load(os.path.join(os.path.dirname(__file__), "golden", "audio.npz"))


def test_resample_and_envelope_match_reference_fixture(cuda):
    from tda_eeg_audio_b200 import audio
    rs = audio.resample_audio(A["audio"].astype(np.float64))
    assert rs.shape == A["resampled"].shape
    np.testing.assert_allclose(rs, A["resampled"], rtol=0, atol=1e-12 * np.abs(A["resampled"]).max())
    env = audio.compute_envelope(A["resampled"], 250)
    np.testing.assert_allclose(env, A["envelope"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(audio.compute_envelope(A["odd_input"], 250), A["odd_envelope"], rtol=0, atol=1e-11)


def test_resample_full_length_batch(cuda):
    """60 s of 44.1 kHz audio -> 15,000 samples at 250 Hz (BASELINE config c), 3 recordings"""
    import torch
    from oracle import signal_ref
    from tda_eeg_audio_b200 import audio
    rng = np.random.default_rng(3)
    n = 2646000
    t = np.arange(n) / 44100.0
    x = np.stack([(1 + 0.6 * np.sin(2 * np.pi * 3.1 * t + p)) * rng.standard_normal(n) for p in (0.1, 1.0, 2.0)])
    y = audio.resample_poly_batched(torch.from_numpy(x).cuda(), 250, 44100).cpu().numpy()
    assert y.shape == (3, 15000)
    for k in range(3):
        ref = signal_ref.resample_audio(x[k])
        np.testing.assert_allclose(y[k], ref, rtol=0, atol=1e-12 * np.abs(ref).max())
    env = audio.compute_envelope_batched(torch.from_numpy(y).cuda(), 250).cpu().numpy()
    for k in range(3):
        np.testing.assert_allclose(env[k], signal_ref.compute_envelope(y[k], 250), rtol=0, atol=1e-11)


@pytest.mark.parametrize("up,down,n", [(3, 7, 1000), (2, 1, 513), (5, 3, 777), (1, 4, 4001), (8, 5, 100)])
def test_resample_other_ratios(cuda, up, down, n):
    import torch
    from scipy import signal
    from tda_eeg_audio_b200 import audio
    x = np.random.default_rng(n).standard_normal((2, n))
    y = audio.resample_poly_batched(torch.from_numpy(x).cuda(), up, down).cpu().numpy()
    for k in range(2):
        np.testing.assert_allclose(y[k], signal.resample_poly(x[k], up, down), rtol=0, atol=1e-12)


def test_raw_audio_to_envelope_chain(cuda):
    import torch
    from oracle import signal_ref
    from tda_eeg_audio_b200 import audio
    x = np.random.default_rng(5).standard_normal((2, 44100 * 4))
    env = audio.audio_envelope_from_raw(torch.from_numpy(x).cuda()).cpu().numpy()
    for k in range(2):
        ref = signal_ref.compute_envelope(signal_ref.resample_audio(x[k]), 250)
        np.testing.assert_allclose(env[k], ref, rtol=0, atol=1e-11)
