"""CPU-only: pins the oracle.  Three independent statements of Rips H0/H1 must agree on pairs:
definition-level boundary reduction (rips_naive), Ripser-style cohomology (rips_cpu.cpp) and the
model of the CUDA kernel's cocycle sweep (pcoh_model)."""
import numpy as np
import pytest

from oracle import pcoh_model, rips, rips_naive
from tests import inputs


def _same(a, b):
    for k in range(2):
        assert a["dgms"][k].shape == b["dgms"][k].shape
        assert np.array_equal(a["dgms"][k], b["dgms"][k])
        assert np.array_equal(a["pairs"][k], b["pairs"][k])


@pytest.mark.parametrize("n", [2, 3, 4, 7, 12, 20])
def test_three_way_small(n):
    rng = np.random.default_rng(100 + n)
    for D in inputs.sym_uniform(rng, 6, n):
        for thr in (np.inf, 0.5):
            for q in (None, 4, 16):
                M = D if q is None else np.round(D * q) / q
                a = rips_naive.rips_h01_naive(M, thr)
                _same(a, rips.ripser(M, thresh=thr, distance_matrix=True))
                _same(a, pcoh_model.rips_h01_pcoh(M, thr))


def test_three_way_47():
    rng = np.random.default_rng(7)
    for D in list(inputs.sym_uniform(rng, 2, 47)) + list(inputs.eeg_like(rng, 2)):
        a = rips_naive.rips_h01_naive(D, 2.0)
        _same(a, rips.ripser(D, thresh=2.0, distance_matrix=True))
        _same(a, pcoh_model.rips_h01_pcoh(D, 2.0))
        Q = np.round(D * 64) / 64
        a = rips_naive.rips_h01_naive(Q, 2.0)
        _same(a, rips.ripser(Q, thresh=2.0, distance_matrix=True))
        _same(a, pcoh_model.rips_h01_pcoh(Q, 2.0))


def test_reference_smoke_input_anchor():
    """The only reproducible input the reference shows a diagram for
    (/root/reference/scripts/tda_eeg_classification_v2.py:253-258, plotted in
    paper/figures/fig_sample_persistence.png): 46 finite H0 bars + 1 essential, ~90 H1 points with
    deaths below ~0.45.  An eyeball anchor, not a golden vector (parity unpinned)."""
    D = np.random.default_rng(42).random((47, 47))
    D = (D + D.T) / 2
    np.fill_diagonal(D, 0)
    r = rips.ripser(D, maxdim=1, thresh=2.0, distance_matrix=True)
    h0, h1 = r["dgms"]
    assert h0.shape == (47, 2) and np.isinf(h0[-1, 1]) and np.isfinite(h0[:-1]).all()
    assert 0.0 < h0[0, 1] < 0.05 and 0.15 < h0[-2, 1] < 0.25
    assert 80 <= len(h1) <= 110 and h1[:, 1].max() < 0.45 and h1[:, 0].min() > 0.08


def test_invariants():
    rng = np.random.default_rng(3)
    D = inputs.eeg_like(rng, 8)
    r = rips.rips_h01_batched(D, 2.0)
    assert (r["counts"][:, 0] == 47).all()          # 46 finite + 1 essential, no zero-length bars
    for b in range(len(D)):
        n1 = r["counts"][b, 1]
        bd = r["bd1"][b, :n1]
        assert (bd[:, 1] > bd[:, 0]).all()
        assert (np.diff(bd[:, 0]) <= 0).all()       # ripser's emission order: descending birth
        # vertex relabelling leaves the multiset of bars unchanged
        perm = rng.permutation(47)
        r2 = rips.rips_h01_batched(D[b][perm][:, perm][None], 2.0)
        m = r2["counts"][0, 1]
        assert m == n1
        assert np.array_equal(np.sort(bd.view([("b", "f4"), ("d", "f4")]).ravel()),
                              np.sort(r2["bd1"][0, :m].view([("b", "f4"), ("d", "f4")]).ravel()))


def test_errors():
    with pytest.raises(Exception, match="not square"):
        rips.ripser(np.zeros((3, 4)), distance_matrix=True)
