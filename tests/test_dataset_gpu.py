"""create_dataset / compute_min_windows_per_band / validate_distance_matrix against
the outputs of the REFERENCE's own functions (tests/golden/dataset.json: executed out of
/root/reference/scripts/tda_eeg_classification_v2.py:110-140, 445-474, 499-606 in the build container
on tests.inputs.small_graph_dataset, ripser replaced by the CPU oracle)."""
import json
import os

import numpy as np
import pytest

from tests import inputs

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "dataset.json")))
TOL = 1e-5      # north_star: features within 1e-5 relative


@pytest.fixture(scope="module")
def graphs(tmp_path_factory):
    return inputs.small_graph_dataset(tmp_path_factory.mktemp("small"))


def test_min_windows_per_band_cpu(graphs):
    """header reads only: runs without a GPU"""
    from tda_eeg_audio_b200 import drivers, dsp
    slow, fast = graphs
    assert drivers.compute_min_windows_per_band([slow, fast], dsp.FREQ_BANDS) == G["min_windows"]
    assert drivers.compute_min_windows_per_band([slow / "missing"], dsp.FREQ_BANDS) == {b: 0 for b in dsp.FREQ_BANDS}


@pytest.mark.gpu
@pytest.mark.parametrize("tag,kw", [
    ("min_random", {}),
    ("all_windows", {"equalize_windows": False, "max_windows_per_band": None}),
    ("fixed12_first_batch1to4", {"max_windows_per_band": 12, "window_sampling": "first", "batch_start": 1,
                                 "batch_end": 4})])
def test_create_dataset(cuda, graphs, tag, kw):
    from tda_eeg_audio_b200 import drivers, dsp
    slow, fast = graphs
    X, y, subjects, names, filenames, meta = drivers.create_dataset(slow, fast, dsp.FREQ_BANDS, verbose=False, **kw)
    ref = G[tag]
    assert filenames == ref["filenames"] and names == ref["feature_names"]
    assert y.tolist() == ref["y"] and subjects.tolist() == ref["subjects"]
    assert [{k: int(v) for k, v in m["n_windows_used"].items()} for m in meta] == ref["n_windows_used"]
    assert [m["validation_issues"] for m in meta] == ref["validation_issues"]
    Xr = np.array([[np.nan if v is None else v for v in row] for row in ref["X"]])
    assert X.shape == Xr.shape and np.array_equal(np.isnan(X), np.isnan(Xr))
    np.testing.assert_allclose(X, Xr, rtol=TOL, atol=1e-9)


@pytest.mark.gpu
def test_validate_distance_matrix(cuda):
    from tda_eeg_audio_b200 import drivers
    rng = np.random.default_rng(5)
    D = inputs.eeg_like(rng, 1)[0].astype(np.float64)
    cases = {"ok": D.copy()}
    c = D.copy(); c[2, 5] += 1e-3; cases["asymmetric"] = c
    c = D.copy(); c[1, 4] = c[4, 1] = -0.25; cases["negative"] = c
    c = D.copy(); c[6, 6] = 1e-3; cases["diagonal"] = c
    c = D.copy(); c[0, 9] = c[9, 0] = np.nan; cases["nan"] = c
    c = D.copy(); c[3, 8] = c[8, 3] = np.inf; cases["inf"] = c
    c = D.copy(); c[0, 1] = np.nan; c[5, 5] = np.nan; c[7, 2] = -np.inf; cases["everything"] = c
    cases["not_square"] = np.zeros((3, 4))
    cases["not_2d"] = np.zeros((3,))
    for k, m in cases.items():
        ok, issues = drivers.validate_distance_matrix(m, k)
        assert ok == G["validate"][k]["valid"] and issues == G["validate"][k]["issues"], k
