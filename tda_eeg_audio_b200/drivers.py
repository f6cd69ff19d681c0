"""Script-level drivers with the reference's names and return conventions,
running every numeric stage on the GPU in one batch per call instead of a Python loop per window:

  preprocess_file            notebooks/1_preprocesamiento.ipynb:340-436
  build_graphs_for_file      notebooks/2_graph_construction.ipynb:100-150
  validate_distance_matrix   scripts/tda_eeg_classification_v2.py:110-140   (device-side checker)
  process_file_features      scripts/tda_eeg_classification_v2.py:338-442   (md5-seeded window choice)
  compute_min_windows_per_band / create_dataset
                             scripts/tda_eeg_classification_v2.py:445-474, 499-606 (a directory tree ->
                             X, y, subjects; ONE Rips launch per band over all recordings)
  get_audio_diagrams / get_eeg_diagrams / compute_cross_wasserstein
                             scripts/matched_vs_mismatched.py:35-95
  process_recording          scripts/tda_eeg_audio_comparison.py:45-127     (linspace window choice,
                             matched W_H0 / W_H1 and the Spearman feature series)
Paths are passed in (the reference hard-codes them relative to its project root)."""
from __future__ import annotations

import hashlib
from pathlib import Path

import numpy as np

from . import audio as _audio
from . import dsp, pipeline, storage
from .features import FEATURE_NAMES, aggregate_windows, diagram_features
from .rips import rips_h01_checked

WINDOW_SEC = 1.0
OVERLAP = 0.75
MAX_WINDOWS = 15
FS_AUDIO = 44100
FS_EEG = 250
TAKENS_DIM = 3
TAKENS_SUBSAMPLE = 2


def _cuda(a, dtype=np.float64):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).cuda()


# ------------------------------------------------------------------------------------------ notebook 1
def preprocess_file(filepath, output_dir, freq_bands=None, window_size=1.0, overlap=0.75, fs_eeg=250,
                    filter_order=4):
    filepath = Path(filepath)
    freq_bands = freq_bands or dsp.FREQ_BANDS
    eeg, audio, fs_file, _ = storage.load_eeg_file(filepath)
    if abs(fs_file - fs_eeg) > 1:
        print(f"Warning: EEG sampling frequency mismatch in {filepath.name}: {fs_file} Hz vs expected {fs_eeg} Hz")
        fs_eeg = fs_file
    x = _cuda(eeg)
    names = list(freq_bands)
    sos = np.stack([dsp.design_bandpass_filter(*freq_bands[b], fs_eeg, filter_order) for b in names])
    filt = dsp.sosfiltfilt_batched(x, sos).cpu().numpy()               # (bands, 47, n)
    band_windows, band_meta, times = {}, {}, None
    for k, b in enumerate(names):
        w, t = dsp.create_sliding_windows(filt[k], window_size, overlap, fs_eeg)
        if w.shape[0] == 0:
            continue
        band_windows[b], times = w, t
        band_meta[b] = {"n_windows": w.shape[0], "window_shape": w.shape, "freq_range": tuple(freq_bands[b])}
    if not band_windows:
        return None
    storage.save_preprocessed(output_dir, filepath.stem, band_windows, times, audio)
    nw = next(iter(band_windows.values())).shape[0]
    return {"filename": filepath.name, "n_electrodes": eeg.shape[0], "n_samples": eeg.shape[1],
            "duration_sec": eeg.shape[1] / fs_eeg, "fs_eeg": fs_eeg, "bands": band_meta, "n_windows": nw}


# ------------------------------------------------------------------------------------------ notebook 2
def build_graphs_for_file(file_dir, output_dir, freq_bands=None, distance_method="euclidean"):
    file_dir = Path(file_dir)
    meta = {"filename": file_dir.name, "n_electrodes": storage.N_ELECTRODES, "bands": {}}
    for band, windows in storage.load_preprocessed(file_dir, freq_bands).items():
        W, C, L = windows.shape
        x = _cuda(windows.transpose(1, 0, 2).reshape(1, C, W * L))     # windows laid end to end
        D, corr = dsp.corrdist_windows(x, L, L, method=distance_method, want_corr=True)
        c64 = corr[0].cpu().numpy()
        d64 = dsp.correlation_to_distance_batched(corr[0], distance_method).cpu().numpy()
        storage.save_graphs(output_dir, file_dir.name, band, c64, d64)
        meta["bands"][band] = {"n_windows": int(W)}
    return meta


# ------------------------------------------------------------------------------------------ features
def select_window_indices(dir_name, band, n_windows, max_windows_per_band, window_sampling, random_state):
    """tda_eeg_classification_v2.py:385-400"""
    if max_windows_per_band is None:
        return np.arange(n_windows)
    max_n = max_windows_per_band.get(band, n_windows) if isinstance(max_windows_per_band, dict) \
        else int(max_windows_per_band)
    max_n = min(max_n, n_windows)
    if window_sampling == "random":
        seed = int(hashlib.md5(f"{dir_name}-{band}-{random_state}".encode()).hexdigest()[:8], 16)
        return np.random.default_rng(seed).choice(n_windows, size=max_n, replace=False)
    return np.arange(max_n)


def _eeg_rips(dist_matrices, thresh, cap1=None, want_pairs=False):
    """compute_persistence_diagram's preparation ((D + D^T)/2, zero diagonal, clamp, f32) + Rips"""
    import torch
    from . import _lib
    d64 = _cuda(dist_matrices)
    B, n, _ = d64.shape
    d32 = torch.empty((B, n, n), dtype=torch.float32, device="cuda")
    _lib.check(_lib.load().tda_symmetrize_f64_to_f32(d64.data_ptr(), B, n, d32.data_ptr(),
                                                    torch.cuda.current_stream().cuda_stream),
               "tda_symmetrize_f64_to_f32")
    return rips_h01_checked(d32, thresh=float(thresh), cap1=cap1, want_pairs=want_pairs)


# the reference's messages (tda_eeg_classification_v2.py:110-140), in its order
def _validation_issues(flags, stats):
    issues = []
    if flags & 1:
        issues.append(f"No simétrica: asimetría máxima={stats[0]:.6f}")
    if flags & 2:
        issues.append(f"Valores negativos presentes: min={stats[1]:.6f}")
    if flags & 4:
        issues.append(f"Diagonal no cero: max={stats[2]:.6f}")
    if flags & 8:
        issues.append("Contiene valores NaN")
    if flags & 16:
        issues.append("Contiene valores Inf")
    return issues


def validate_distance_matrices(D64):
    """Batched checker: D64 CUDA float64 (B, n, n) -> (flags (B,) int32, stats (B, 3) float64) on the
    device (TDA_DM_* bits of include/tda_b200.h)."""
    import torch
    from . import _lib
    D64 = D64.contiguous()
    B, n, _ = D64.shape
    flags = torch.zeros((B,), dtype=torch.int32, device=D64.device)
    stats = torch.zeros((B, 3), dtype=torch.float64, device=D64.device)
    _lib.check(_lib.load().tda_validate_distance_f64(D64.data_ptr(), B, n, flags.data_ptr(), stats.data_ptr(),
                                                    torch.cuda.current_stream().cuda_stream),
               "tda_validate_distance_f64")
    return flags, stats


def validate_distance_matrix(distance_matrix, name=""):
    """tda_eeg_classification_v2.validate_distance_matrix: (is_valid, issues) with the reference's own
    messages; the checks themselves run on the device."""
    distance_matrix = np.asarray(distance_matrix)
    if distance_matrix.ndim != 2:
        return False, [f"No es 2D: forma={distance_matrix.shape}"]
    n, m = distance_matrix.shape
    if n != m:
        return False, [f"No es cuadrada: forma=({n}, {m})"]
    flags, stats = validate_distance_matrices(_cuda(distance_matrix)[None])
    issues = _validation_issues(int(flags[0].item()), stats[0].tolist())
    return len(issues) == 0, issues


def process_file_features(file_dir, freq_bands, max_dim=1, max_edge_length=2.0, max_windows_per_band=None,
                          window_sampling="random", random_state=42, verbose=False):
    file_dir = Path(file_dir)
    file_features = {}
    metadata = {"n_windows": {}, "n_windows_used": {}, "validation_issues": [],
                "window_sampling": window_sampling, "max_windows_per_band": max_windows_per_band}
    for band in freq_bands:
        dm = storage.load_distances(file_dir, band)
        if dm is None:
            if verbose:
                print(f"  Warning: {band}_distances.npy not found")
            metadata["n_windows"][band] = 0
            continue
        n_windows = dm.shape[0]
        metadata["n_windows"][band] = n_windows
        if n_windows == 0:
            continue
        is_valid, issues = validate_distance_matrix(dm[0], f"{band}[0]")
        if not is_valid:
            metadata["validation_issues"].extend([f"{band}: {i}" for i in issues])
        use = select_window_indices(file_dir.name, band, n_windows, max_windows_per_band, window_sampling,
                                    random_state)
        metadata["n_windows_used"][band] = len(use)
        if len(use) == 0:
            continue
        r = _eeg_rips(dm[use], max_edge_length)
        f = diagram_features(r)                                        # (n_used, 2, 11) on the device
        row = aggregate_windows(f.view(1, 1, len(use), 2, 11))[0].cpu().numpy()   # 44 values, feature-major
        for k, feat in enumerate(FEATURE_NAMES):
            file_features[f"{band}_h0_{feat}_mean"] = row[4 * k]
            file_features[f"{band}_h0_{feat}_std"] = row[4 * k + 1]
            file_features[f"{band}_h1_{feat}_mean"] = row[4 * k + 2]
            file_features[f"{band}_h1_{feat}_std"] = row[4 * k + 3]
    metadata["n_windows_total"] = int(sum(metadata["n_windows"].values()))
    metadata["n_windows_used_total"] = int(sum(metadata["n_windows_used"].values()))
    return file_features, metadata


def compute_min_windows_per_band(graphs_dirs, freq_bands):
    """tda_eeg_classification_v2.py:445-474: smallest window count per band over the whole data set
    (array headers only: np.load(mmap_mode="r"))."""
    min_windows = {band: np.inf for band in freq_bands}
    for graphs_dir in graphs_dirs:
        graphs_dir = Path(graphs_dir)
        if not graphs_dir.exists():
            continue
        for file_dir in (d for d in graphs_dir.iterdir() if d.is_dir()):
            for band in freq_bands:
                dist_file = file_dir / f"{band}_distances.npy"
                if not dist_file.exists():
                    continue
                try:
                    n_windows = np.load(dist_file, mmap_mode="r").shape[0]
                except Exception:
                    continue
                if n_windows > 0:
                    min_windows[band] = min(min_windows[band], n_windows)
    return {band: (0 if val == np.inf else int(val)) for band, val in min_windows.items()}


def create_dataset(graphs_dir_slow, graphs_dir_fast, freq_bands, max_dim=1, max_edge_length=2.0,
                   equalize_windows=True, window_sampling="random", max_windows_per_band="min", random_state=42,
                   batch_start=0, batch_end=None, verbose=True):
    """tda_eeg_classification_v2.create_dataset (:499-606): graphs/<slow|fast>/<recording>/<band>_distances.npy
    -> (X, y, subjects, feature_names, filenames, metadata), same entry order (sorted slow then sorted
    fast), same window choice, same column order, same skipping of recordings without features.

    Where the reference calls process_file_features per recording (a ripser call per window), all the
    selected windows of ALL recordings of a band go through one symmetrise + Rips + feature launch; the
    mean / std over a recording's windows is one aggregation launch per distinct window count (one,
    when the windows are equalised)."""
    import torch
    graphs_dir_slow, graphs_dir_fast = Path(graphs_dir_slow), Path(graphs_dir_fast)
    say = print if verbose else (lambda *a, **k: None)
    if equalize_windows:
        if max_windows_per_band == "min":
            max_windows_per_band = compute_min_windows_per_band([graphs_dir_slow, graphs_dir_fast], freq_bands)
            say("\nEqualizando ventanas por banda (min global):")
            for band, nmin in max_windows_per_band.items():
                say(f"  {band}: {nmin} ventanas")
        else:
            say(f"\nEqualizando ventanas por banda (máx fijo): {max_windows_per_band}")
    slow_dirs = sorted(d for d in graphs_dir_slow.iterdir() if d.is_dir())
    fast_dirs = sorted(d for d in graphs_dir_fast.iterdir() if d.is_dir())
    entries = [(d, 0) for d in slow_dirs] + [(d, 1) for d in fast_dirs]
    total_entries = len(entries)
    if batch_end is None or batch_end < 0:
        batch_end = total_entries
    batch_start = max(0, batch_start)
    batch_end = min(batch_end, total_entries)
    entries = entries[batch_start:batch_end]
    say(f"\nProcesando {len(entries)} archivos (batch {batch_start}:{batch_end}) de total {total_entries}...")

    E = len(entries)
    features = [dict() for _ in range(E)]
    metas = [{"n_windows": {}, "n_windows_used": {}, "validation_issues": [], "window_sampling": window_sampling,
              "max_windows_per_band": max_windows_per_band} for _ in range(E)]
    failed = [False] * E
    for band in freq_bands:
        mats, owner, first_of = [], [], []       # selected windows of every recording, their entry, first windows
        for k, (file_dir, _) in enumerate(entries):
            if failed[k]:
                continue
            try:
                dm = storage.load_distances(file_dir, band)
            except Exception as exc:              # the reference: a load error is noted, the band skipped
                metas[k]["validation_issues"].append(f"{band}: error de carga - {exc}")
                continue
            if dm is None:
                metas[k]["n_windows"][band] = 0
                continue
            n_windows = dm.shape[0]
            metas[k]["n_windows"][band] = n_windows
            if n_windows == 0:
                continue
            try:
                use = select_window_indices(file_dir.name, band, n_windows, max_windows_per_band, window_sampling,
                                            random_state)
            except Exception as exc:              # the reference: any failure drops the recording
                say(f"Error procesando {file_dir.name}: {exc}")
                failed[k] = True
                continue
            metas[k]["n_windows_used"][band] = len(use)
            first_of.append((k, dm[0]))
            if len(use) == 0:
                continue
            mats.append(dm[use])
            owner.append((k, len(use)))
        if first_of:
            fl, stt = validate_distance_matrices(_cuda(np.stack([m for _, m in first_of])))
            fl, stt = fl.tolist(), stt.tolist()
            for (k, _), f, sv in zip(first_of, fl, stt):
                metas[k]["validation_issues"].extend([f"{band}: {i}" for i in _validation_issues(f, sv)])
        if not mats:
            continue
        r = _eeg_rips(np.concatenate(mats), max_edge_length)          # one launch for the whole band
        f = diagram_features(r)                                       # (n_total, 2, 11) on the device
        rows = {}
        starts = np.concatenate([[0], np.cumsum([n for _, n in owner])])
        for cnt in sorted({n for _, n in owner}):
            grp = [i for i, (_, n) in enumerate(owner) if n == cnt]
            idx = torch.from_numpy(np.concatenate([np.arange(starts[i], starts[i] + cnt) for i in grp])).cuda()
            agg = aggregate_windows(f[idx].view(len(grp), 1, cnt, 2, 11)).cpu().numpy()      # (len(grp), 44)
            for row, i in zip(agg, grp):
                rows[owner[i][0]] = row
        for k, row in rows.items():
            for q, feat in enumerate(FEATURE_NAMES):
                features[k][f"{band}_h0_{feat}_mean"] = row[4 * q]
                features[k][f"{band}_h0_{feat}_std"] = row[4 * q + 1]
                features[k][f"{band}_h1_{feat}_mean"] = row[4 * q + 2]
                features[k][f"{band}_h1_{feat}_std"] = row[4 * q + 3]

    all_features, all_labels, all_subjects, all_filenames, all_metadata = [], [], [], [], []
    for k, (file_dir, label) in enumerate(entries):
        if failed[k] or len(features[k]) == 0:
            continue
        meta = metas[k]
        meta["n_windows_total"] = int(sum(meta["n_windows"].values()))
        meta["n_windows_used_total"] = int(sum(meta["n_windows_used"].values()))
        filename = file_dir.name
        parts = filename.split("_")
        meta["filename"], meta["subject"], meta["label"] = filename, (parts[0] if parts else filename), label
        all_features.append(features[k])
        all_labels.append(label)
        all_subjects.append(meta["subject"])
        all_filenames.append(filename)
        all_metadata.append(meta)
    # pd.DataFrame(list of dicts): columns in order of first appearance, missing entries NaN
    feature_names = []
    for fd in all_features:
        for name in fd:
            if name not in feature_names:
                feature_names.append(name)
    X = np.array([[fd.get(name, np.nan) for name in feature_names] for fd in all_features], dtype=np.float64) \
        if all_features else np.zeros((0, 0))
    y = np.array(all_labels)
    subjects = np.array(all_subjects)
    say("Resumen del Dataset")
    say("-" * 60)
    say(f"Total de muestras: {X.shape[0]}")
    say(f"Total de características: {X.shape[1] if X.ndim == 2 else 0}")
    return X, y, subjects, feature_names, all_filenames, all_metadata


# ------------------------------------------------------------------------------------------ EEG-audio
def _diagram_lists(r, max_dim=1):
    """padded diagram tensors -> list (per item) of [H0 (k,2), H1 (k,2)] float64 arrays"""
    cnt = r["counts"].cpu().numpy()
    bd0, bd1 = r["bd0"].double().cpu().numpy(), r["bd1"].double().cpu().numpy()
    assert (cnt[:, 1] <= bd1.shape[1]).all(), "truncated H1 diagrams (use rips_h01_checked)"
    return [[bd0[i, :cnt[i, 0]], bd1[i, :cnt[i, 1]]][: max_dim + 1] for i in range(len(cnt))]


def _audio_band_diagrams(envelope, want_tensors=False):
    """envelope (T,) float64 -> per band result of pipeline.audio_diagrams_from_envelope for one recording"""
    env = _cuda(envelope)[None]
    return pipeline.audio_diagrams_from_envelope(env, fs=FS_EEG, window_sec=WINDOW_SEC, overlap=OVERLAP,
                                                 takens_dim=TAKENS_DIM, subsample=TAKENS_SUBSAMPLE,
                                                 max_windows=MAX_WINDOWS)


def get_audio_diagrams(mat_path):
    """matched_vs_mismatched.get_audio_diagrams: {band: [dgms per selected window]}"""
    mat_path = Path(mat_path)
    if not mat_path.exists():
        return None
    a = _audio.load_audio(mat_path)
    env = _audio.audio_envelope_from_raw(_cuda(a)[None], FS_AUDIO, FS_EEG)
    res = pipeline.audio_diagrams_from_envelope(env, fs=FS_EEG, window_sec=WINDOW_SEC, overlap=OVERLAP,
                                                takens_dim=TAKENS_DIM, subsample=TAKENS_SUBSAMPLE,
                                                max_windows=MAX_WINDOWS)
    R, nb, n_sel = res["shape"]
    if n_sel == 0:
        return {}
    lists = _diagram_lists(res["rips"])
    npts = res["npts"].cpu().numpy()
    out = {}
    for b, band in enumerate(dsp.FREQ_BANDS):
        items = range(b * n_sel, (b + 1) * n_sel)
        out[band] = [lists[i] for i in items if npts[i] >= 3]
    return out


def get_eeg_diagrams(graph_dir):
    """matched_vs_mismatched.get_eeg_diagrams: {band: [dgms per selected window]}"""
    graph_dir = Path(graph_dir)
    if not graph_dir.exists():
        return None
    out = {}
    for band in dsp.FREQ_BANDS:
        dm = storage.load_distances(graph_dir, band)
        if dm is None or dm.shape[0] == 0:
            continue
        idx = pipeline.select_windows(dm.shape[0], MAX_WINDOWS)
        out[band] = _diagram_lists(_eeg_rips(dm[idx], 2.0))
    return out


def compute_cross_wasserstein(eeg_dgms_band, audio_dgms_band):
    """matched_vs_mismatched.compute_cross_wasserstein: nanmean of W_H1 over windows paired by position"""
    from .wasserstein import safe_wasserstein
    n = min(len(eeg_dgms_band), len(audio_dgms_band))
    if n == 0:
        return np.nan
    return np.nanmean([safe_wasserstein(eeg_dgms_band[i][1], audio_dgms_band[i][1]) for i in range(n)])


def process_recording(mat_path, graph_dir):
    """tda_eeg_audio_comparison.process_recording for one recording (paths passed explicitly)."""
    import torch
    from scipy.stats import spearmanr
    mat_path, graph_dir = Path(mat_path), Path(graph_dir)
    if not mat_path.exists() or not graph_dir.exists():
        return None
    filename = mat_path.name
    results = {"filename": filename, "condition": mat_path.parent.name, "subject": filename.split("_")[0], "bands": {}}
    a = _audio.load_audio(mat_path)
    env = _audio.audio_envelope_from_raw(_cuda(a)[None], FS_AUDIO, FS_EEG)[0]
    win = int(WINDOW_SEC * FS_EEG)
    step = int(win * (1 - OVERLAP))
    n_audio_win = dsp.n_windows(env.shape[0], win, step)
    for b, (bname, (lo, hi)) in enumerate(dsp.FREQ_BANDS.items()):
        dm = storage.load_distances(graph_dir, bname)
        if dm is None:
            continue
        n_win = min(n_audio_win, dm.shape[0])
        if n_win == 0:
            continue
        idx = pipeline.select_windows(n_win, MAX_WINDOWS)
        ares = pipeline.audio_diagrams_from_envelope(env[None], fs=FS_EEG, bands={bname: (lo, hi)}, window_sec=WINDOW_SEC,
                                                     overlap=OVERLAP, takens_dim=TAKENS_DIM, subsample=TAKENS_SUBSAMPLE,
                                                     max_windows=None, window_idx=idx)
        keep = (ares["npts"] >= 3).cpu().numpy()
        if not keep.any():
            continue
        eres = _eeg_rips(dm[idx], 2.0)
        w0, w1 = pipeline.cross_wasserstein(eres, ares["rips"])
        w0, w1 = w0.cpu().numpy()[keep], w1.cpu().numpy()[keep]
        fa = diagram_features(ares["rips"])[:, 1].cpu().numpy()[keep]
        fe = diagram_features(eres)[:, 1].cpu().numpy()[keep]
        feat_corrs = {}
        for feat in ["mean_persistence", "total_persistence", "persistence_entropy", "max_persistence", "n_features"]:
            k = FEATURE_NAMES.index(feat)
            a_ts, e_ts = fa[:, k], fe[:, k]
            if len(a_ts) >= 5 and np.std(a_ts) > 1e-10 and np.std(e_ts) > 1e-10:
                r, p = spearmanr(a_ts, e_ts)
                feat_corrs[feat] = {"r": float(r), "p": float(p)}
            else:
                feat_corrs[feat] = {"r": 0.0, "p": 1.0}
        results["bands"][bname] = {"wasserstein_h0": float(np.nanmean(w0)), "wasserstein_h1": float(np.nanmean(w1)),
                                   "n_windows": len(idx), "tau": int(ares["tau"][0, 0].item()),
                                   "feature_correlations": feat_corrs}
    return results if results["bands"] else None
