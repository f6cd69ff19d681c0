"""CPU-only: the C-ABI library loads and exports every symbol include/tda_b200.h declares."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "tda_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tda_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    from tda_eeg_audio_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 8
    for n in names:
        assert hasattr(lib, n), f"{n} declared in tda_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in _lib.py"
    assert sorted(_lib.SIGNATURES) == names


def test_argument_errors_without_gpu():
    from tda_eeg_audio_b200 import _lib
    lib = _lib.load()
    assert lib.tda_version() >= 100
    assert lib.tda_rips_h01_workspace_bytes(10, 65) == 0
    assert lib.tda_rips_h01_workspace_bytes(10, 47) > 0
    # null pointers / bad sizes are rejected before any CUDA call
    assert lib.tda_rips_h01_batched(None, 1, 47, 47, 0, 2.0, None, None, None, None, None, 1, None, None, 0, None) == -1
    assert lib.tda_pers_features(None, 1, None, 1, 1, None, 11, None) == -1
    # a dense leading dimension below N is an error; 0 announces the condensed upper triangle
    import ctypes
    buf = (ctypes.c_char * 64)()
    a = ctypes.addressof(buf)
    assert lib.tda_rips_h01_batched(a, 1, 47, 46, 0, 2.0, a, None, a, None, a, 1, a, a, 0, None) == -1
    assert lib.tda_rips_h01_batched(a, 1, 47, 0, 0, 2.0, a, None, a, None, a, 1, a, a, 0, None) == -3  # workspace too small
    assert lib.tda_rips_h01_condensed_host(None, 1, 47, 2.0, None, None, None, None, None, 1, None, 0) == -1
    assert lib.tda_eeg_features_condensed_host(None, 1, 5, 60, 47, 2.0, 128, None, None, None, None, None, None, 0) == -1


def test_no_cpu_fallback():
    import pytest
    import torch
    from tda_eeg_audio_b200 import _lib, rips_h01_batched
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.TdaError):
        rips_h01_batched(torch.zeros((1, 4, 4)))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "tda_eeg_audio_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("oracle/pcoh_model.py", "").replace("oracle/pcoh_large_model.cpp", ""), f"{f} mentions the oracle"


def test_argument_errors_of_the_cloud_and_audio_entry_points():
    """size limits and null pointers are rejected before any CUDA call (no GPU needed)"""
    from tda_eeg_audio_b200 import _lib
    lib = _lib.load()
    assert lib.tda_rips_h01_large_workspace_bytes(4, 2049) == 0
    assert lib.tda_rips_h01_large_workspace_bytes(4, 300) > 0
    assert lib.tda_rips_h01_large(None, None, 1, 300, 300, 0, 2.0, None, None, 300, None, None, 8, None, None, None, 0, None) == -1
    assert lib.tda_resample_poly_f64(None, 1, 10, 10, 5, 882, None, 3529, 10, 15, None, 15, None) == -1
    assert lib.tda_hilbert_envelope_workspace_bytes(2, 15000) >= 2 * 15000 * 16
    assert lib.tda_hilbert_envelope_f64(None, 1, 10, 10, None, 10, None, 0, None) == -1
    assert lib.tda_filtfilt_f64(None, 1, 100, 100, 0, 1, 4, None, None, 27, None, None, 0, None) == -1


def test_condensed_order_is_ripser_py_own_expression():
    """rips.condense must produce exactly the vector ripser.py hands its C++ core for a dense matrix:
    `I, J = np.meshgrid(np.arange(n), np.arange(n)); DParam = np.array(dm[I > J], dtype=np.float32)`
    (ripser.py, dense branch) -- the upper triangle in row-major order."""
    import numpy as np
    import torch
    from tda_eeg_audio_b200.rips import condense
    rng = np.random.default_rng(0)
    for n in (2, 3, 5, 47, 64):
        dm = rng.random((n, n))          # deliberately NOT symmetric: the order must pick dm[i, j], i < j
        I, J = np.meshgrid(np.arange(n), np.arange(n))
        dparam = np.array(dm[I > J], dtype=np.float32)
        assert np.array_equal(condense(dm.astype(np.float32)), dparam)
        assert np.array_equal(condense(torch.from_numpy(dm.astype(np.float32))[None]).numpy()[0], dparam)
        iu = np.triu_indices(n, 1)
        assert np.array_equal(dparam, dm[iu].astype(np.float32))
