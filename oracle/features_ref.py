"""TEST INFRASTRUCTURE ONLY — numpy restatement of the reference's feature extraction and window
aggregation.  Pinned by tests/golden/features.npz, which make_golden.py generates by calling the
reference's OWN extract_features (/root/reference/scripts/utils.py:144-177) in this container."""
from __future__ import annotations

import numpy as np

FEATURE_NAMES = ["n_features", "n_essential", "mean_birth", "std_birth", "mean_death", "std_death",
                 "mean_persistence", "std_persistence", "max_persistence", "total_persistence",
                 "persistence_entropy"]
BANDS = ["delta", "theta", "alpha", "beta", "gamma"]


def extract_features_vec(diagram):
    """utils.extract_features as an 11-vector (same key order as the reference's dict)."""
    d = np.asarray(diagram, dtype=np.float64).reshape(-1, 2)
    fin = np.isfinite(d).all(axis=1)
    fd = d[fin]
    out = np.zeros(11)
    out[1] = np.sum(~fin)
    if len(fd) == 0:
        return out
    b, de = fd[:, 0], fd[:, 1]
    p = de - b
    ent = 0.0
    if len(p) > 1 and np.sum(p) > 0:
        pn = p / np.sum(p)
        pn = pn[pn > 0]
        ent = -np.sum(pn * np.log(pn + 1e-10)) / np.log(len(p) + 1e-10)
    many = len(fd) > 1
    out[:] = [len(fd), np.sum(~fin), b.mean(), b.std() if many else 0, de.mean(), de.std() if many else 0,
              p.mean(), p.std() if many else 0, p.max(), p.sum(), ent]
    return out


def aggregate_windows(feats):
    """feats (R, Bd, Wn, 2, 11) -> (R, Bd*44): mean/std over windows in the column order of
    /root/reference/features/feature_names.txt
    (/root/reference/scripts/tda_eeg_classification_v2.py:429-436)."""
    R, Bd, Wn, _, _ = feats.shape
    out = np.zeros((R, Bd * 44))
    for band in range(Bd):
        for f in range(11):
            for dim in range(2):
                x = feats[:, band, :, dim, f]
                out[:, band * 44 + f * 4 + dim * 2] = np.mean(x, axis=1)
                out[:, band * 44 + f * 4 + dim * 2 + 1] = np.std(x, axis=1)
    return out


def feature_column_names(bands=BANDS):
    names = []
    for band in bands:
        for f in FEATURE_NAMES:
            names += [f"{band}_h0_{f}_mean", f"{band}_h0_{f}_std", f"{band}_h1_{f}_mean", f"{band}_h1_{f}_std"]
    return names
