"""Generates tests/golden/*.npz by running the REFERENCE's own functions in the build container
(/root/reference/scripts/utils.py imported where it lies, with ripser/persim replaced by the CPU
oracle — oracle/reference_import.py).  The fixtures travel to the GPU box; the reference does not.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import reference_import  # noqa: E402
from tests import inputs  # noqa: E402


def main():
    u = reference_import.load_utils()
    rng = np.random.default_rng(20261018)
    # ---- features: reference extract_features on diagrams of the reference's own persistence calls
    out = {}
    dgms = []
    for D in inputs.eeg_like(rng, 4):
        dgms += u.compute_eeg_persistence(D.astype(np.float64))
    for D in inputs.sym_uniform(rng, 2, 47):
        dgms += u.compute_eeg_persistence(D.astype(np.float64))
    dgms += [np.zeros((0, 2)), np.array([[0.0, np.inf]]), np.array([[0.1, 0.4]]),
             np.array([[0.0, 0.0], [0.0, 0.0]])]
    for k, d in enumerate(dgms):
        f = u.extract_features(d)
        out[f"dgm{k}"] = np.asarray(d, np.float64)
        out[f"feat{k}"] = np.array([float(v) for v in f.values()])
    out["n"] = len(dgms)
    out["names"] = np.array(list(u.extract_features(dgms[0]).keys()))
    np.savez_compressed(os.path.join(HERE, "features.npz"), **out)
    print("features.npz:", len(dgms), "diagrams")


if __name__ == "__main__":
    main()


def signal_golden():
    """utils.bandpass_filter / create_windows / compute_tau / takens_embedding from the reference's
    own module, on seeded inputs."""
    u = reference_import.load_utils()
    rng = np.random.default_rng(7)
    out = {}
    env = np.abs(rng.standard_normal(3000)) + 0.3 * np.sin(np.arange(3000) / 40.0)
    out["env"] = env
    for name, (lo, hi) in u.FREQ_BANDS.items():
        out[f"bp_{name}"] = u.bandpass_filter(env, 250, lo, hi)
    out["windows_alpha"] = u.create_windows(out["bp_alpha"], 250, 62)
    taus = []
    for name in u.FREQ_BANDS:
        wins = u.create_windows(out[f"bp_{name}"], 250, 62)
        taus.append([u.compute_tau(w, max_lag=125) for w in wins])
    out["taus"] = np.array(taus)
    w0 = out["windows_alpha"][0]
    out["takens_tau7_sub2"] = u.takens_embedding(w0, 3, 7, 2)
    out["takens_tau12_sub1"] = u.takens_embedding(w0, 3, 12, 1)
    np.savez_compressed(os.path.join(HERE, "signal.npz"), **out)
    print("signal.npz:", sorted(out))


if __name__ == "__main__":
    signal_golden()


def audio_golden():
    """utils.resample_audio / compute_envelope from the reference's own module on a seeded 0.6 s
    amplitude-modulated noise burst (stored as float32 so the fixture stays small)."""
    u = reference_import.load_utils()
    rng = np.random.default_rng(11)
    n = 26460
    t = np.arange(n) / 44100.0
    audio = ((1 + 0.6 * np.sin(2 * np.pi * 3.1 * t + 0.3)) * rng.standard_normal(n)).astype(np.float32)
    a64 = audio.astype(np.float64)
    rs = u.resample_audio(a64)
    env = u.compute_envelope(rs, 250)
    odd_in = np.abs(rng.standard_normal(1501)) + 0.2                              # odd length
    np.savez_compressed(os.path.join(HERE, "audio.npz"), audio=audio, resampled=rs, envelope=env,
                        odd_input=odd_in, odd_envelope=u.compute_envelope(odd_in, 250))
    print("audio.npz:", rs.shape, env.shape)


if __name__ == "__main__":
    audio_golden()


def names_golden():
    """the reference's own features/feature_names.txt (column order of X.npy)"""
    import shutil
    shutil.copyfile(os.path.join(reference_import.REFERENCE_ROOT, "features", "feature_names.txt"),
                    os.path.join(HERE, "feature_names.txt"))
    print("feature_names.txt copied")


if __name__ == "__main__":
    names_golden()


def _reference_functions(script, names, namespace):
    """exec the named top-level function definitions of a reference script where it lies (the
    scripts themselves cannot be imported: they need seaborn/matplotlib and run their analysis at
    import time)."""
    import ast
    path = os.path.join(reference_import.REFERENCE_ROOT, "scripts", script)
    tree = ast.parse(open(path).read(), filename=path)
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert {n.name for n in body} == set(names), (script, names)
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), namespace)
    return namespace


def drivers_golden():
    """process_file_features (tda_eeg_classification_v2.py) and process_recording
    (tda_eeg_audio_comparison.py) of the reference, run on tests.inputs.tiny_dataset with
    ripser/persim replaced by the CPU oracle."""
    import hashlib
    import json
    import tempfile
    from pathlib import Path
    from scipy.stats import spearmanr
    from oracle import rips as orips
    u = reference_import.load_utils()
    with tempfile.TemporaryDirectory() as td:
        mat, gdir = inputs.tiny_dataset(td)
        ns = {"np": np, "hashlib": hashlib, "ripser": orips.ripser}
        _reference_functions("tda_eeg_classification_v2.py",
                             ["process_file_features", "compute_persistence_diagram", "extract_persistence_features",
                              "validate_distance_matrix"], ns)
        out = {}
        for tag, kw in (("all", {}), ("rand10", {"max_windows_per_band": 10, "window_sampling": "random"}),
                        ("first7", {"max_windows_per_band": {"alpha": 7}, "window_sampling": "first"})):
            feats, meta = ns["process_file_features"](Path(gdir), u.FREQ_BANDS, **kw)
            out[tag] = {"features": {k: float(v) for k, v in feats.items()},
                        "n_windows_used": meta["n_windows_used"], "n_windows": meta["n_windows"]}
        ns2 = {k: getattr(u, k) for k in dir(u) if not k.startswith("_")}
        ns2.update({"np": np, "spearmanr": spearmanr, "DATA_DIR": Path(td) / "data", "GRAPHS_DIR": Path(td) / "graphs",
                    "WINDOW_SEC": 1.0, "OVERLAP": 0.75, "MAX_WINDOWS": 15})
        _reference_functions("tda_eeg_audio_comparison.py", ["process_recording"], ns2)
        out["process_recording"] = ns2["process_recording"]("S01_trial1.mat", "slow")
        ns3 = dict(ns2)
        _reference_functions("matched_vs_mismatched.py", ["get_audio_diagrams", "get_eeg_diagrams",
                                                          "compute_cross_wasserstein"], ns3)
        a = ns3["get_audio_diagrams"]("S01_trial1.mat", "slow")
        e = ns3["get_eeg_diagrams"]("S01_trial1.mat", "slow")
        out["cross_wasserstein_h1"] = {b: float(ns3["compute_cross_wasserstein"](e[b], a[b])) for b in u.FREQ_BANDS}
        out["n_audio_diagrams"] = {b: len(a[b]) for b in u.FREQ_BANDS}
    json.dump(out, open(os.path.join(HERE, "drivers.json"), "w"), indent=1, sort_keys=True)
    print("drivers.json:", sorted(out), out["process_recording"]["bands"]["alpha"]["wasserstein_h1"])


if __name__ == "__main__":
    drivers_golden()


def _notebook_functions(notebook, names, namespace):
    """exec the named function definitions of a reference NOTEBOOK where it lies: the code cells are
    parsed with ast (IPython magics dropped) and only the top-level `def`s asked for are executed --
    the cells' own analysis code is not run and nothing is copied."""
    import ast
    import json
    path = os.path.join(reference_import.REFERENCE_ROOT, "notebooks", notebook)
    body = []
    for cell in json.load(open(path))["cells"]:
        if cell["cell_type"] != "code":
            continue
        src = "".join(cell["source"])
        src = "\n".join(ln for ln in src.splitlines() if not ln.lstrip().startswith(("%", "!")))
        try:
            tree = ast.parse(src)
        except SyntaxError:
            continue
        body += [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert {n.name for n in body} == set(names), (notebook, names, [n.name for n in body])
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), namespace)
    return namespace


def notebooks_golden():
    """design/apply_bandpass_filter, create_sliding_windows (notebook 1) and compute_correlation_matrix,
    correlation_to_distance (notebook 2) of the reference, on seeded inputs -> notebooks.npz."""
    from scipy import signal
    u = reference_import.load_utils()
    ns = {"np": np, "signal": signal}
    _notebook_functions("1_preprocesamiento.ipynb",
                        ["design_bandpass_filter", "apply_bandpass_filter", "create_sliding_windows"], ns)
    _notebook_functions("2_graph_construction.ipynb", ["compute_correlation_matrix", "correlation_to_distance"], ns)
    rng = np.random.default_rng(20261018)
    out = {}
    x = rng.standard_normal((4, 8)) / np.sqrt(8) @ rng.standard_normal((8, 1000)) + 0.5 * rng.standard_normal((4, 1000))
    out["x"] = x
    for name, (lo, hi) in u.FREQ_BANDS.items():
        out[f"sos_{name}"] = ns["design_bandpass_filter"](lo, hi, 250, 4)
        out[f"filt_{name}"] = ns["apply_bandpass_filter"](x, lo, hi, 250, 4)
    wins, times = ns["create_sliding_windows"](out["filt_alpha"], 1.0, 0.75, 250)
    out["windows_alpha"], out["window_times"] = wins, times
    w = rng.standard_normal((47, 8)) / np.sqrt(8) @ rng.standard_normal((8, 250)) + 0.5 * rng.standard_normal((47, 250))
    w[5] = 2.5            # zero variance -> NaN -> 0
    w[9] = w[8]           # duplicate channel -> r = 1 -> d = 0
    out["window47"] = w
    c = ns["compute_correlation_matrix"](w)
    out["corr47"] = c
    for m in ("euclidean", "abs", "standard", "sqrt"):
        out[f"dist47_{m}"] = ns["correlation_to_distance"](c, method=m)
    np.savez_compressed(os.path.join(HERE, "notebooks.npz"), **out)
    print("notebooks.npz:", sorted(out)[:6], "...", wins.shape, c.shape)


if __name__ == "__main__":
    notebooks_golden()


def dataset_golden():
    """create_dataset / compute_min_windows_per_band / validate_distance_matrix of the reference
    (tda_eeg_classification_v2.py:110-140, 445-474, 499-606), executed where they lie on
    tests.inputs.small_graph_dataset with ripser replaced by the CPU oracle -> dataset.json."""
    import hashlib
    import io
    import json
    import tempfile
    from contextlib import redirect_stdout
    from pathlib import Path
    import pandas as pd
    from oracle import rips as orips
    u = reference_import.load_utils()
    ns = {"np": np, "pd": pd, "hashlib": hashlib, "ripser": orips.ripser, "N_JOBS": 1,
          "tqdm": lambda it, **kw: it}
    _reference_functions("tda_eeg_classification_v2.py",
                         ["create_dataset", "compute_min_windows_per_band", "process_file_features",
                          "compute_persistence_diagram", "extract_persistence_features", "validate_distance_matrix"], ns)
    out = {}
    with tempfile.TemporaryDirectory() as td:
        slow, fast = inputs.small_graph_dataset(td)
        out["min_windows"] = ns["compute_min_windows_per_band"]([slow, fast], u.FREQ_BANDS)
        for tag, kw in (("min_random", {}),
                        ("all_windows", {"equalize_windows": False, "max_windows_per_band": None}),
                        ("fixed12_first_batch1to4", {"max_windows_per_band": 12, "window_sampling": "first",
                                                     "batch_start": 1, "batch_end": 4})):
            with redirect_stdout(io.StringIO()):
                X, y, subjects, names, filenames, meta = ns["create_dataset"](slow, fast, u.FREQ_BANDS, **kw)
            out[tag] = {"X": [[None if v != v else float(v) for v in row] for row in X], "y": [int(v) for v in y],
                        "subjects": list(map(str, subjects)), "feature_names": names, "filenames": filenames,
                        "n_windows_used": [m["n_windows_used"] for m in meta],
                        "validation_issues": [m["validation_issues"] for m in meta]}
    # the input checker on its own
    rng = np.random.default_rng(5)
    D = inputs.eeg_like(rng, 1)[0].astype(np.float64)
    cases = {"ok": D.copy()}
    c = D.copy(); c[2, 5] += 1e-3; cases["asymmetric"] = c
    c = D.copy(); c[1, 4] = c[4, 1] = -0.25; cases["negative"] = c
    c = D.copy(); c[6, 6] = 1e-3; cases["diagonal"] = c
    c = D.copy(); c[0, 9] = c[9, 0] = np.nan; cases["nan"] = c
    c = D.copy(); c[3, 8] = c[8, 3] = np.inf; cases["inf"] = c
    c = D.copy(); c[0, 1] = np.nan; c[5, 5] = np.nan; c[7, 2] = -np.inf; cases["everything"] = c
    out["validate"] = {}
    for k, m in cases.items():
        ok, issues = ns["validate_distance_matrix"](m, k)
        out["validate"][k] = {"valid": bool(ok), "issues": issues}
    out["validate"]["not_square"] = dict(zip(("valid", "issues"), ns["validate_distance_matrix"](np.zeros((3, 4)))))
    out["validate"]["not_2d"] = dict(zip(("valid", "issues"), ns["validate_distance_matrix"](np.zeros((3,)))))
    json.dump(out, open(os.path.join(HERE, "dataset.json"), "w"), indent=1, sort_keys=True, ensure_ascii=False)
    print("dataset.json:", {k: (len(v["X"]), len(v["feature_names"])) for k, v in out.items() if isinstance(v, dict) and "X" in v},
          out["min_windows"])


if __name__ == "__main__":
    dataset_golden()
