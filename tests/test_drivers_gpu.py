"""GPU parity of the script-level drivers against the outputs of the REFERENCE's own script
functions (tests/golden/drivers.json: process_file_features of tda_eeg_classification_v2.py,
process_recording of tda_eeg_audio_comparison.py and get_*_diagrams / compute_cross_wasserstein of
matched_vs_mismatched.py, executed in the build container on tests.inputs.tiny_dataset with
ripser / persim replaced by the CPU oracle)."""
import json
import os

import numpy as np
import pytest

from tests import inputs

pytestmark = pytest.mark.gpu
G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "drivers.json")))
TOL = 1e-5      # north_star: features and Wasserstein distances within 1e-5 relative
# The delta band of the AUDIO chain goes through utils.bandpass_filter's ba-form Butterworth
# (0.5-4 Hz at 250 Hz), which is ill-conditioned: scipy's own result (lfilter_zi solves a linear
# system with LAPACK) moves by ~1e-5 between the build container's CPU and the GPU box's CPU.  The
# fixture is therefore compared at 1e-4 for that band, and the same band is compared per window
# against the scipy/oracle chain run on the SAME box (test_delta_band_against_same_box_oracle, 1e-5 from the same envelope).
TOL_DELTA_AUDIO = 1e-4


@pytest.fixture(scope="module")
def dataset(tmp_path_factory):
    root = tmp_path_factory.mktemp("tiny")
    mat, gdir = inputs.tiny_dataset(root)
    return root, mat, gdir


def test_process_file_features(cuda, dataset):
    from tda_eeg_audio_b200 import drivers, dsp
    _, _, gdir = dataset
    for tag, kw in (("all", {}), ("rand10", {"max_windows_per_band": 10, "window_sampling": "random"}),
                    ("first7", {"max_windows_per_band": {"alpha": 7}, "window_sampling": "first"})):
        feats, meta = drivers.process_file_features(gdir, dsp.FREQ_BANDS, **kw)
        ref = G[tag]
        assert list(feats) == list(ref["features"]) or sorted(feats) == sorted(ref["features"])
        assert {k: int(v) for k, v in meta["n_windows_used"].items()} == ref["n_windows_used"]
        for k, v in ref["features"].items():
            assert feats[k] == pytest.approx(v, rel=TOL, abs=1e-9), (tag, k)


def test_feature_dict_order_is_the_table_order(cuda, dataset):
    from tda_eeg_audio_b200 import drivers, dsp, storage
    feats, _ = drivers.process_file_features(dataset[2], dsp.FREQ_BANDS)
    assert list(feats) == storage.feature_names()


def test_process_recording(cuda, dataset):
    from tda_eeg_audio_b200 import drivers
    _, mat, gdir = dataset
    got = drivers.process_recording(mat, gdir)
    ref = G["process_recording"]
    assert got["filename"] == ref["filename"] and got["subject"] == ref["subject"] and got["condition"] == "slow"
    assert sorted(got["bands"]) == sorted(ref["bands"])
    for b, rb in ref["bands"].items():
        gb = got["bands"][b]
        assert gb["tau"] == rb["tau"] and gb["n_windows"] == rb["n_windows"]
        tol = TOL_DELTA_AUDIO if b == "delta" else TOL
        assert gb["wasserstein_h0"] == pytest.approx(rb["wasserstein_h0"], rel=tol)
        assert gb["wasserstein_h1"] == pytest.approx(rb["wasserstein_h1"], rel=tol)
        for f, c in rb["feature_correlations"].items():
            if b == "delta":
                continue   # rank statistics of a series that itself moves between machines (see above)
            assert gb["feature_correlations"][f]["r"] == pytest.approx(c["r"], abs=1e-9), (b, f)
            assert gb["feature_correlations"][f]["p"] == pytest.approx(c["p"], abs=1e-9), (b, f)


def test_delta_band_against_same_box_oracle(cuda, dataset):
    """process_recording's chain for the ill-conditioned delta band, restated with scipy + the CPU
    oracle on this very machine (tda_eeg_audio_comparison.py:57-100), per window, FROM THE SAME
    ENVELOPE.  The GPU envelope differs from scipy's by 7e-16 (FFT rounding, tests/test_audio_gpu.py);
    the ill-conditioned delta recursion followed by the min-max normalisation of nearly flat
    windows amplifies even that to ~1e-4 in W — the reference moves as much against itself between
    two CPUs — so the chain after the envelope is what can be, and is, held to 1e-5 here."""
    from oracle import rips as orips, signal_ref, wasserstein_ref
    from tda_eeg_audio_b200 import audio as _audio, dsp, pipeline
    from tda_eeg_audio_b200.drivers import _cuda, _eeg_rips
    _, mat, gdir = dataset
    a = _audio.load_audio(mat)
    env_ref = signal_ref.compute_envelope(signal_ref.resample_audio(a), 250)
    env = _cuda(env_ref)
    lo, hi = dsp.FREQ_BANDS["delta"]
    wins = signal_ref.create_windows(signal_ref.bandpass_filter(env_ref, 250, lo, hi), 250, 62)
    dm = np.load(gdir / "delta_distances.npy")
    idx = pipeline.select_windows(min(len(wins), dm.shape[0]), 15)
    tau = signal_ref.compute_tau(wins[idx[0]], max_lag=125)
    ares = pipeline.audio_diagrams_from_envelope(env[None], bands={"delta": (lo, hi)}, max_windows=None, window_idx=idx)
    w0, w1 = pipeline.cross_wasserstein(_eeg_rips(dm[idx], 2.0), ares["rips"])
    assert int(ares["tau"][0, 0]) == tau

    def clean(x):
        x = x[np.isfinite(x).all(1)]
        return x if len(x) else np.array([[0.0, 0.0]])
    for k, w in enumerate(idx):
        pc = signal_ref.takens_embedding(wins[w], 3, tau, 2)
        mn = pc.min(0); rg = pc.max(0) - mn; rg[rg == 0] = 1
        ra = orips.ripser((pc - mn) / rg, thresh=2.0)
        d = dm[w]; d = (d + d.T) / 2; np.fill_diagonal(d, 0); d = np.maximum(d, 0)
        re = orips.ripser(d, thresh=2.0, distance_matrix=True)
        for dim, got in ((0, w0), (1, w1)):
            ref = wasserstein_ref.wasserstein(clean(re["dgms"][dim]), clean(ra["dgms"][dim]))
            assert float(got[k]) == pytest.approx(ref, rel=1e-5), (k, dim)


def test_matched_mismatched_helpers(cuda, dataset):
    from tda_eeg_audio_b200 import drivers
    _, mat, gdir = dataset
    a = drivers.get_audio_diagrams(mat)
    e = drivers.get_eeg_diagrams(gdir)
    assert {b: len(v) for b, v in a.items()} == G["n_audio_diagrams"]
    for b, v in G["cross_wasserstein_h1"].items():
        assert drivers.compute_cross_wasserstein(e[b], a[b]) == pytest.approx(v, rel=TOL_DELTA_AUDIO if b == "delta" else TOL)
    assert drivers.get_audio_diagrams(str(mat) + ".missing") is None


def test_preprocess_and_graphs_reproduce_the_layout(cuda, dataset):
    """notebook 1 + 2 drivers: the .npy files they write equal the scipy restatement's"""
    from tda_eeg_audio_b200 import drivers
    root, mat, gdir = dataset
    meta = drivers.preprocess_file(mat, root / "preprocessed" / "slow")
    assert meta["n_electrodes"] == 47 and meta["fs_eeg"] == 250 and meta["n_windows"] == 29
    drivers.build_graphs_for_file(root / "preprocessed" / "slow" / "S01_trial1", root / "graphs_gpu" / "slow")
    for band in ("delta", "alpha", "gamma"):
        ref_d = np.load(gdir / f"{band}_distances.npy")
        ref_c = np.load(gdir / f"{band}_correlations.npy")
        got_d = np.load(root / "graphs_gpu" / "slow" / "S01_trial1" / f"{band}_distances.npy")
        got_c = np.load(root / "graphs_gpu" / "slow" / "S01_trial1" / f"{band}_correlations.npy")
        assert got_d.shape == ref_d.shape == (29, 47, 47) and got_d.dtype == np.float64
        np.testing.assert_allclose(got_c, ref_c, rtol=0, atol=1e-10)
        np.testing.assert_allclose(got_d, ref_d, rtol=0, atol=1e-8)
