"""ctypes binding of libtda_b200.so (the C-ABI in include/tda_b200.h).

There is no CPU fallback: if the CUDA library is missing the import fails loudly, and every
compute entry point refuses to run without a CUDA device.
"""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtda_b200.so")

_c = ctypes
_vp, _i, _ll, _f, _sz = _c.c_void_p, _c.c_int, _c.c_longlong, _c.c_float, _c.c_size_t

# name -> (restype, argtypes); must list every symbol include/tda_b200.h declares
SIGNATURES = {
    "tda_version": (_i, []),
    "tda_launch_count": (_c.c_ulonglong, []),
    "tda_profile_enable": (_i, [_i]),
    "tda_profile_query": (_i, [_c.c_char_p, _c.POINTER(_c.c_double), _c.POINTER(_c.c_int)]),
    "tda_rips_h01_workspace_bytes": (_sz, [_i, _i]),
    "tda_rips_h01_batched": (_i, [_vp, _i, _i, _i, _ll, _f, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _sz, _vp]),
    "tda_pers_features": (_i, [_vp, _i, _vp, _i, _i, _vp, _i, _vp]),
    "tda_aggregate_windows": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "tda_filtfilt_workspace_bytes": (_sz, [_ll, _i, _ll, _i]),
    "tda_filtfilt_f64": (_i, [_vp, _ll, _ll, _ll, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _sz, _vp]),
    "tda_corrdist_windows": (_i, [_vp, _i, _i, _ll, _ll, _i, _i, _i, _vp, _vp, _ll, _vp]),
    "tda_corr_to_dist_f64": (_i, [_vp, _i, _i, _vp, _vp]),
    "tda_symmetrize_f64_to_f32": (_i, [_vp, _ll, _i, _vp, _vp]),
    "tda_validate_distance_f64": (_i, [_vp, _ll, _i, _vp, _vp, _vp]),
    "tda_compute_tau": (_i, [_vp, _ll, _i, _ll, _i, _vp, _vp]),
    "tda_takens_cloud": (_i, [_vp, _ll, _i, _ll, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "tda_pairwise_dist_f32": (_i, [_vp, _vp, _ll, _i, _i, _i, _vp, _vp]),
    "tda_wasserstein_batched": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _i, _i, _i, _vp, _vp, _ll, _vp, _vp]),
    "tda_wasserstein_batched_f64": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _i, _i, _i, _vp, _vp, _ll, _vp, _vp]),
    "tda_eeg_features_host": (_i, [_vp, _i, _i, _i, _i, _f, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i]),
    "tda_rips_h01_large_workspace_bytes": (_sz, [_i, _i]),
    "tda_rips_h01_large": (_i, [_vp, _vp, _i, _i, _i, _ll, _f, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "tda_resample_poly_f64": (_i, [_vp, _ll, _ll, _ll, _i, _i, _vp, _i, _ll, _ll, _vp, _ll, _vp]),
    "tda_hilbert_envelope_workspace_bytes": (_sz, [_ll, _ll]),
    "tda_hilbert_envelope_f64": (_i, [_vp, _ll, _ll, _ll, _vp, _ll, _vp, _sz, _vp]),
    "tda_rips_h01_host": (_i, [_vp, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i]),
    "tda_rips_h01_condensed_host": (_i, [_vp, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i]),
    "tda_eeg_features_condensed_host": (_i, [_vp, _i, _i, _i, _i, _f, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i]),
    "tda_eeg_features_f64_host": (_i, [_vp, _i, _i, _i, _i, _f, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i]),
}

_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m tda_eeg_audio_b200.build` "
                "(nvcc, sm_100a). There is no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError => header/library mismatch, fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class TdaError(RuntimeError):
    pass


def check(rc: int, what: str):
    if rc == 0:
        return
    if rc < 0:
        names = {-1: "TDA_E_ARG", -2: "TDA_E_SIZE", -3: "TDA_E_WORKSPACE"}
        raise TdaError(f"{what}: {names.get(rc, rc)}")
    raise TdaError(f"{what}: CUDA error {rc}")


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise TdaError("tda_eeg_audio_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def launch_count() -> int:
    return int(load().tda_launch_count())


def profile_enable(on: bool):
    load().tda_profile_enable(1 if on else 0)


def profile_query(kernel: str):
    """(total_ms, launches) of `kernel` since profile_enable(True); synchronises its events."""
    ms, n = _c.c_double(0), _c.c_int(0)
    check(load().tda_profile_query(kernel.encode(), _c.byref(ms), _c.byref(n)), "tda_profile_query")
    return ms.value, n.value
