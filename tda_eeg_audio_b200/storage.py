"""On-disk layouts of the reference pipeline, so that someone holding the
private data can point this engine at the same directories:

  data/<cond>/<rec>.mat                       subeeg, y, Fs          (1_preprocesamiento.ipynb:132-156)
  preprocessed/<cond>/<rec>/<band>.npy        (W, 47, win) float64   (1_preprocesamiento.ipynb:404-423)
  preprocessed/<cond>/<rec>/window_times.npy, audio.npy
  graphs/<cond>/<rec>/<band>_correlations.npy (W, 47, 47) float64    (2_graph_construction.ipynb:137-138)
  graphs/<cond>/<rec>/<band>_distances.npy    (W, 47, 47) float64
  features/X.npy, y.npy, subjects.npy, feature_names.txt, filenames.txt
                                              (tda_eeg_classification_v2.py: create_dataset / main)
Host-side file handling only (numpy / scipy.io); no compute happens here."""
from __future__ import annotations

from pathlib import Path

import numpy as np

from .dsp import FREQ_BANDS
from .features import FEATURE_NAMES

# 1-based MATLAB indices of the 47 good electrodes (/root/reference/notebooks/1_preprocesamiento.ipynb:66-115)
GOOD_ELECTRODES_MATLAB = [2, 3, 4, 6, 7, 9, 11, 12, 13, 14, 15, 16, 18, 19, 20, 21, 22, 24, 25, 26, 27, 28, 30, 31,
                          33, 34, 36, 38, 40, 41, 42, 44, 45, 46, 48, 49, 50, 51, 52, 53, 54, 56, 57, 58, 59, 60, 65]
GOOD_ELECTRODES = [x - 1 for x in GOOD_ELECTRODES_MATLAB]
N_ELECTRODES = len(GOOD_ELECTRODES)


def load_eeg_file(filepath):
    """1_preprocesamiento.ipynb:117-156: (eeg (47, n) good electrodes, mono audio, fs_eeg, fs_audio).
    The EEG matrix is transposed when stored (samples, electrodes); fs_eeg is inferred from the
    audio duration."""
    from scipy.io import loadmat
    data = loadmat(str(filepath))
    eeg_all = data["subeeg"]
    audio = data["y"]
    fs_audio = int(data["Fs"][0, 0])
    if eeg_all.shape[0] > eeg_all.shape[1]:
        eeg_all = eeg_all.T
    eeg = eeg_all[GOOD_ELECTRODES, :]
    audio_duration = audio.shape[0] / fs_audio
    fs_eeg = int(round(eeg.shape[1] / audio_duration))
    if audio.ndim > 1:
        audio = audio.mean(axis=1)
    return eeg, audio, fs_eeg, fs_audio


def feature_names(bands=None):
    """The 220 column names of X.npy in the order of features/feature_names.txt:
    {band}_{h0|h1}_{feat}_{mean|std}, feature-major, (h0_mean, h0_std, h1_mean, h1_std) per feature
    (/root/reference/scripts/tda_eeg_classification_v2.py:429-436)."""
    out = []
    for band in (bands or FREQ_BANDS):
        for feat in FEATURE_NAMES:
            out += [f"{band}_h0_{feat}_mean", f"{band}_h0_{feat}_std", f"{band}_h1_{feat}_mean", f"{band}_h1_{feat}_std"]
    return out


def save_preprocessed(out_dir, stem, band_windows, window_times, audio=None):
    d = Path(out_dir) / stem
    d.mkdir(parents=True, exist_ok=True)
    for band, w in band_windows.items():
        np.save(d / f"{band}.npy", np.asarray(w, dtype=np.float64))
    np.save(d / "window_times.npy", np.asarray(window_times))
    if audio is not None:
        np.save(d / "audio.npy", np.asarray(audio))
    return d


def load_preprocessed(file_dir, bands=None):
    d = Path(file_dir)
    return {b: np.load(d / f"{b}.npy") for b in (bands or FREQ_BANDS) if (d / f"{b}.npy").exists()}


def save_graphs(out_dir, stem, band, correlations, distances):
    d = Path(out_dir) / stem
    d.mkdir(parents=True, exist_ok=True)
    np.save(d / f"{band}_correlations.npy", np.asarray(correlations, dtype=np.float64))
    np.save(d / f"{band}_distances.npy", np.asarray(distances, dtype=np.float64))
    return d


def load_distances(file_dir, band):
    f = Path(file_dir) / f"{band}_distances.npy"
    return np.load(f) if f.exists() else None


def save_feature_dataset(out_dir, X, y, subjects, filenames, names=None):
    d = Path(out_dir)
    d.mkdir(parents=True, exist_ok=True)
    np.save(d / "X.npy", np.asarray(X, dtype=np.float64))
    np.save(d / "y.npy", np.asarray(y))
    np.save(d / "subjects.npy", np.asarray(subjects))
    (d / "feature_names.txt").write_text("\n".join(names or feature_names()) + "\n")
    (d / "filenames.txt").write_text("\n".join(filenames) + "\n")
    return d


def load_feature_dataset(in_dir):
    d = Path(in_dir)
    names = (d / "feature_names.txt").read_text().split()
    files = (d / "filenames.txt").read_text().split()
    return np.load(d / "X.npy"), np.load(d / "y.npy"), np.load(d / "subjects.npy", allow_pickle=True), names, files
