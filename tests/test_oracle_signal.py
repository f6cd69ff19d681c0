"""CPU-only: the signal oracle against the fixtures generated from the reference's own utils.py
(tests/golden/signal.npz), and the host-side drop-ins that need no GPU."""
import os

import numpy as np

from oracle import signal_ref

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "signal.npz"))


def test_bandpass_ba_matches_reference_fixture():
    for name, (lo, hi) in signal_ref.FREQ_BANDS.items():
        got = signal_ref.bandpass_filter(G["env"], 250, lo, hi)
        np.testing.assert_allclose(got, G[f"bp_{name}"], rtol=0, atol=1e-12 * np.abs(G[f"bp_{name}"]).max())


def test_windows_tau_takens_match_reference_fixture():
    assert np.array_equal(signal_ref.create_windows(G["bp_alpha"], 250, 62), G["windows_alpha"])
    for b, name in enumerate(signal_ref.FREQ_BANDS):
        wins = signal_ref.create_windows(G[f"bp_{name}"], 250, 62)
        assert [signal_ref.compute_tau(w, max_lag=125) for w in wins] == list(G["taus"][b])
    w0 = G["windows_alpha"][0]
    assert np.array_equal(signal_ref.takens_embedding(w0, 3, 7, 2), G["takens_tau7_sub2"])
    assert np.array_equal(signal_ref.takens_embedding(w0, 3, 12, 1), G["takens_tau12_sub1"])


def test_window_dropins_are_pure_slicing():
    from tda_eeg_audio_b200 import dsp
    x = np.random.default_rng(4).standard_normal((3, 1000))
    a, ta = dsp.create_sliding_windows(x, 1.0, 0.75, 250)
    b, tb = signal_ref.create_sliding_windows(x, 1.0, 0.75, 250)
    assert np.array_equal(a, b) and np.array_equal(ta, tb)
    assert np.array_equal(dsp.create_windows(x[0], 250, 62), signal_ref.create_windows(x[0], 250, 62))
    assert dsp.create_windows(x[0][:100], 250, 62).shape == (0, 250)


def test_audio_front_end_matches_reference_fixture():
    """oracle restatement of resample_audio / compute_envelope vs the reference's own outputs."""
    A = np.load(os.path.join(os.path.dirname(__file__), "golden", "audio.npz"))
    rs = signal_ref.resample_audio(A["audio"].astype(np.float64))
    np.testing.assert_allclose(rs, A["resampled"], rtol=0, atol=1e-13 * np.abs(A["resampled"]).max())
    np.testing.assert_allclose(signal_ref.compute_envelope(rs, 250), A["envelope"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(signal_ref.compute_envelope(A["odd_input"], 250), A["odd_envelope"], rtol=0, atol=1e-12)


def test_resample_design_matches_scipy():
    """the host-side filter / alignment the CUDA resampler is fed reproduces scipy's upfirdn call"""
    from scipy import signal
    from tda_eeg_audio_b200 import audio
    x = np.random.default_rng(2).standard_normal(5000)
    for up, down in ((250, 44100), (3, 7), (2, 1), (5, 3)):
        u, d, hpoly, n_pre, n_out = audio.design_resample(len(x), up, down)
        h = hpoly.T.reshape(-1)              # h_padded (zero tail)
        y = signal.upfirdn(h, x, u, d)[n_pre:n_pre + n_out]
        ref = signal.resample_poly(x, up, down)
        assert len(ref) == n_out
        np.testing.assert_allclose(y, ref, rtol=0, atol=1e-13)


def test_notebook_functions_match_reference_fixture():
    """tests/golden/notebooks.npz holds the outputs of the reference's OWN notebook functions
    (notebooks/1_preprocesamiento.ipynb: design/apply_bandpass_filter, create_sliding_windows;
    notebooks/2_graph_construction.ipynb: compute_correlation_matrix, correlation_to_distance),
    executed from the .ipynb where it lies (tests/golden/make_golden.py::notebooks_golden).  The
    restatements the GPU parity tests compare against must reproduce them exactly."""
    N = np.load(os.path.join(os.path.dirname(__file__), "golden", "notebooks.npz"))
    x = N["x"]
    for name, (lo, hi) in signal_ref.FREQ_BANDS.items():
        assert np.array_equal(signal_ref.design_bandpass_filter(lo, hi, 250, 4), N[f"sos_{name}"])
        assert np.array_equal(signal_ref.apply_bandpass_filter(x, lo, hi, 250), N[f"filt_{name}"])
    wins, times = signal_ref.create_sliding_windows(N["filt_alpha"], 1.0, 0.75, 250)
    assert np.array_equal(wins, N["windows_alpha"]) and np.array_equal(times, N["window_times"])
    c = signal_ref.compute_correlation_matrix(N["window47"])
    assert np.array_equal(c, N["corr47"])
    assert (c[5] == 0).all() and c[8, 9] == 1.0          # zero-variance row -> 0, duplicate channel -> 1
    for m in ("euclidean", "abs", "standard", "sqrt"):
        assert np.array_equal(signal_ref.correlation_to_distance(N["corr47"], m), N[f"dist47_{m}"])
    # the host-side design the CUDA filter is fed is the same call
    from tda_eeg_audio_b200 import dsp
    for name, (lo, hi) in signal_ref.FREQ_BANDS.items():
        assert np.array_equal(dsp.design_bandpass_filter(lo, hi, 250, 4), N[f"sos_{name}"])
