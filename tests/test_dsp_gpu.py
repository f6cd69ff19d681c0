"""GPU parity of the signal stages (FP64): batched zero-phase IIR (sos and ba forms), fused
window -> correlation -> distance, and the drop-in helpers.  Oracle = the scipy/numpy calls the
reference makes (oracle/signal_ref.py).  Tolerances are written per test; north_star asks 1e-5
relative, the kernels land many orders tighter because they replay scipy's recursion exactly."""
import os

import numpy as np
import pytest

from oracle import signal_ref

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def test_sosfiltfilt_all_bands(cuda):
    import torch
    from scipy import signal
    from tda_eeg_audio_b200 import dsp
    rng = np.random.default_rng(0)
    x = signal_ref.eeg_like_recording(rng, C=47, T=3000) * 20 + 5.0
    sos = np.stack([signal_ref.design_bandpass_filter(lo, hi, 250) for lo, hi in signal_ref.FREQ_BANDS.values()])
    y = dsp.sosfiltfilt_batched(torch.from_numpy(x).cuda(), sos).cpu().numpy()
    for b in range(5):
        ref = signal.sosfiltfilt(sos[b], x, axis=1)
        assert _rel(y[b], ref) < 1e-12, (b, _rel(y[b], ref))


@pytest.mark.parametrize("T", [28, 29, 64, 250, 1001])
def test_sosfiltfilt_ragged_lengths(cuda, T):
    import torch
    from scipy import signal
    from tda_eeg_audio_b200 import dsp
    x = np.random.default_rng(T).standard_normal((5, T))
    sos = signal_ref.design_bandpass_filter(8, 13, 250)
    y = dsp.sosfiltfilt_batched(torch.from_numpy(x).cuda(), sos[None]).cpu().numpy()[0]
    assert _rel(y, signal.sosfiltfilt(sos, x, axis=1)) < 1e-12


def test_filtfilt_too_short_raises_like_scipy(cuda):
    import torch
    from tda_eeg_audio_b200 import dsp
    sos = signal_ref.design_bandpass_filter(8, 13, 250)
    with pytest.raises(ValueError, match="padlen"):
        dsp.sosfiltfilt_batched(torch.zeros((2, 27), dtype=torch.float64, device="cuda"), sos[None])


def test_ba_filtfilt_bands_and_lowpass(cuda):
    """the ill-conditioned ba-form delta band is the hard case"""
    import torch
    from scipy import signal
    from tda_eeg_audio_b200 import dsp
    rng = np.random.default_rng(1)
    env = np.abs(rng.standard_normal((3, 15000))) + 0.3 * np.sin(np.arange(15000) / 40.0)
    ba = [signal.butter(4, [max(lo / 125, 0.001), min(hi / 125, 0.999)], btype="band")
          for lo, hi in signal_ref.FREQ_BANDS.values()]
    y = dsp.filtfilt_batched(torch.from_numpy(env).cuda(), ba).cpu().numpy()
    for b in range(5):
        for r in range(3):
            ref = signal_ref.bandpass_filter(env[r], 250, *list(signal_ref.FREQ_BANDS.values())[b])
            assert _rel(y[b, r], ref) < 1e-9, (b, r, _rel(y[b, r], ref))
    bl, al = signal.butter(4, 50 / 125, btype="low")
    yl = dsp.filtfilt_batched(torch.from_numpy(env).cuda(), [(bl, al)]).cpu().numpy()[0]
    assert _rel(yl, signal.filtfilt(bl, al, env, axis=1)) < 1e-12
    # drop-in with the reference's signature
    got = dsp.bandpass_filter(env[0], 250, 0.5, 4)
    assert _rel(got, signal_ref.bandpass_filter(env[0], 250, 0.5, 4)) < 1e-9
    s0 = env[0]
    assert dsp.bandpass_filter(s0, 250, 200, 100) is s0                # lo >= hi: input returned untouched


@pytest.mark.parametrize("step", [250, 62])
def test_corrdist_windows(cuda, step):
    import torch
    from tda_eeg_audio_b200 import dsp
    rng = np.random.default_rng(2)
    x = signal_ref.eeg_like_recording(rng, C=47, T=2000)
    x = signal_ref.apply_bandpass_filter(x, 8, 13, 250)
    x2 = np.stack([x, x[::-1] * 3 + 1.0])
    D, corr = dsp.corrdist_windows(torch.from_numpy(x2).cuda(), 250, step, want_corr=True)
    D, corr = D.cpu().numpy(), corr.cpu().numpy()
    for r in range(2):
        wins = signal_ref.create_windows(x2[r].T, 250, step)  # (W, 250, C)
        assert D.shape[1] == len(wins)
        for w in range(len(wins)):
            c = signal_ref.compute_correlation_matrix(wins[w].T)
            d = signal_ref.correlation_to_distance(c)
            np.testing.assert_allclose(corr[r, w], c, rtol=0, atol=1e-13)
            # 1e-5 relative is the contract; off-diagonal distances here are >= 0.05
            np.testing.assert_allclose(D[r, w], d.astype(np.float32), rtol=2e-7, atol=0)
            assert np.array_equal(D[r, w], D[r, w].T) and (np.diag(D[r, w]) == 0).all()


def test_corrdist_constant_channel_and_methods(cuda):
    import torch
    from tda_eeg_audio_b200 import dsp
    rng = np.random.default_rng(3)
    w = rng.standard_normal((47, 250))
    w[5] = 2.5                       # zero variance -> NaN correlations -> 0 -> d = sqrt(2)
    w[9] = w[8]                      # duplicate channel -> r = 1 -> d = 0
    c = dsp.compute_correlation_matrix(w)
    cref = signal_ref.compute_correlation_matrix(w)
    np.testing.assert_allclose(c, cref, rtol=0, atol=1e-13)
    assert (c[5] == 0).all() and c[8, 9] == 1.0
    for m in ("euclidean", "abs", "standard", "sqrt"):
        np.testing.assert_allclose(dsp.correlation_to_distance(c, m), signal_ref.correlation_to_distance(cref, m),
                                   rtol=1e-12, atol=1e-7 if m in ("euclidean", "sqrt") else 1e-13)
    with pytest.raises(ValueError, match="Unknown method"):
        dsp.correlation_to_distance(c, "nope")


def test_windows_dropins():
    from tda_eeg_audio_b200 import dsp
    rng = np.random.default_rng(4)
    x = rng.standard_normal((3, 1000))
    a, ta = dsp.create_sliding_windows(x, 1.0, 0.75, 250)
    b, tb = signal_ref.create_sliding_windows(x, 1.0, 0.75, 250)
    assert np.array_equal(a, b) and np.array_equal(ta, tb)
    assert np.array_equal(dsp.create_windows(x[0], 250, 62), signal_ref.create_windows(x[0], 250, 62))
    assert dsp.create_windows(x[0][:100], 250, 62).shape == (0, 250)
    e, _ = dsp.create_sliding_windows(x[:, :100], 1.0, 0.75, 250)
    assert e.size == 0


def test_raw_eeg_to_distances_end_to_end(cuda):
    """config (a): one synthetic recording, 5 bands, reference default 75 % overlap."""
    import torch
    from tda_eeg_audio_b200 import dsp
    rng = np.random.default_rng(20261018)
    x = signal_ref.eeg_like_recording(rng, T=2500)
    ref = signal_ref.eeg_distances(x)                    # (5, W, 47, 47) float64
    got = dsp.eeg_distances_from_raw(torch.from_numpy(np.stack([x, x])).cuda(), rec_chunk=1).cpu().numpy()
    assert got.shape == (2, 5) + ref.shape[1:]
    for r in range(2):
        err = np.abs(got[r] - ref) / np.maximum(ref, 1e-30)
        off = ~np.eye(47, dtype=bool)
        assert err[..., off].max() < 1e-5, err[..., off].max()      # north_star tolerance
        assert err[..., off].max() < 5e-7                            # what we actually reach (f32 rounding)
