"""CPU-only: on-disk layouts and host-side driver logic (no GPU compute)."""
import hashlib
import os

import numpy as np

from tda_eeg_audio_b200 import storage

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_feature_names_match_reference_file():
    ref = open(os.path.join(GOLD, "feature_names.txt")).read().split()
    assert storage.feature_names() == ref and len(ref) == 220


def test_mat_loading_and_layout_roundtrip(tmp_path):
    from scipy.io import savemat
    rng = np.random.default_rng(0)
    n = 500
    sub = rng.standard_normal((n, 66))                      # stored (samples, electrodes) -> transposed
    y = rng.standard_normal((n * 176, 2))
    savemat(tmp_path / "S01_t1.mat", {"subeeg": sub, "y": y, "Fs": np.array([[44000]])})
    eeg, audio, fs_eeg, fs_audio = storage.load_eeg_file(tmp_path / "S01_t1.mat")
    assert eeg.shape == (47, n) and fs_audio == 44000 and fs_eeg == 250
    assert np.array_equal(eeg, sub.T[storage.GOOD_ELECTRODES]) and np.allclose(audio, y.mean(axis=1))
    w = {"alpha": rng.standard_normal((3, 47, 250))}
    d = storage.save_preprocessed(tmp_path / "pre", "S01_t1", w, np.arange(3.0), audio)
    assert np.array_equal(storage.load_preprocessed(d)["alpha"], w["alpha"])
    g = storage.save_graphs(tmp_path / "graphs", "S01_t1", "alpha", np.eye(47)[None], np.zeros((1, 47, 47)))
    assert storage.load_distances(g, "alpha").shape == (1, 47, 47) and storage.load_distances(g, "beta") is None
    storage.save_feature_dataset(tmp_path / "features", np.zeros((2, 220)), [0, 1], ["S01", "S02"], ["a.mat", "b.mat"])
    X, yv, subj, names, files = storage.load_feature_dataset(tmp_path / "features")
    assert X.shape == (2, 220) and names == storage.feature_names() and files == ["a.mat", "b.mat"]


def test_window_selection_is_the_reference_rule():
    from tda_eeg_audio_b200 import drivers, pipeline
    seed = int(hashlib.md5(b"S01_t1-alpha-42").hexdigest()[:8], 16)
    want = np.random.default_rng(seed).choice(238, size=60, replace=False)
    assert np.array_equal(drivers.select_window_indices("S01_t1", "alpha", 238, 60, "random", 42), want)
    assert np.array_equal(drivers.select_window_indices("S01_t1", "alpha", 238, {"alpha": 10}, "first", 42), np.arange(10))
    assert np.array_equal(drivers.select_window_indices("x", "alpha", 7, None, "random", 42), np.arange(7))
    assert np.array_equal(pipeline.select_windows(238, 15), np.linspace(0, 237, 15, dtype=int))
    assert np.array_equal(pipeline.select_windows(9, 15), np.arange(9))
