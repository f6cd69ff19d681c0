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


def test_large_engine_model_matches_oracle():
    """oracle/pcoh_large_model.cpp (the algorithm of rips_large.cu: edge classification + cocycle
    sweep over visible edges, unified exact handling of tie runs) against the Ripser restatement."""
    from oracle import pcoh_large
    rng = np.random.default_rng(11)
    for trial in range(600):
        n = int(rng.integers(3, 26))
        kind = trial % 5
        if kind == 0:
            D = inputs.sym_uniform(rng, 1, n)[0]
        elif kind == 1:
            D = np.round(inputs.sym_uniform(rng, 1, n)[0] * 8) / 8
        elif kind == 2:
            P = rng.random((n, 2)); D = np.sqrt(((P[:, None] - P[None]) ** 2).sum(-1))
        elif kind == 3:
            P = rng.integers(0, 4, (n, 2)).astype(float); D = np.sqrt(((P[:, None] - P[None]) ** 2).sum(-1))
        else:
            D = np.full((n, n), 0.5); np.fill_diagonal(D, 0)
        thr = (np.inf, 0.6, 2.0)[trial % 3]
        _same(rips.ripser(D.astype(np.float32), thresh=thr, distance_matrix=True), pcoh_large.rips_h01(D, thr))
    for D in inputs.circle_cloud(rng, 2, 150):
        a = pcoh_large.rips_h01(D, 2.0)
        _same(rips.ripser(D, thresh=2.0, distance_matrix=True), a)
        assert a["stats"]["visited_edges"] < a["stats"]["edges"]


def _gf2_rank(rows):
    """rank over GF(2) of a list of python-int bit rows"""
    rank = 0
    rows = list(rows)
    while rows:
        r = rows.pop()
        if r == 0:
            continue
        rank += 1
        low = r & -r
        rows = [x ^ r if x & low else x for x in rows]
    return rank


def test_oracle_against_independent_libraries_and_linear_algebra():
    """Anchors that do not share code with the oracle: (1) the finite H0 deaths are the weights of
    scipy's minimum spanning tree; (2) at any scale t the number of H1 bars alive is the first Betti
    number of the Rips complex, computed from GF(2) ranks of its boundary matrices."""
    from itertools import combinations
    from scipy.sparse.csgraph import minimum_spanning_tree
    rng = np.random.default_rng(5)
    for D in list(inputs.circle_cloud(rng, 3, 14)) + list(inputs.sym_uniform(rng, 3, 12)):
        n = len(D)
        r = rips.ripser(D, thresh=np.inf, distance_matrix=True)
        h0 = r["dgms"][0]
        mst = np.sort(minimum_spanning_tree(np.triu(D.astype(np.float64), 1)).data.astype(np.float32))
        assert np.array_equal(h0[:-1, 1].astype(np.float32), mst) and np.isinf(h0[-1, 1])
        h1 = r["dgms"][1]
        for t in np.quantile(D[np.triu_indices(n, 1)], [0.2, 0.4, 0.6, 0.8]).astype(np.float32):
            edges = [(i, j) for i, j in combinations(range(n), 2) if D[i, j] <= t]
            eid = {e: k for k, e in enumerate(edges)}
            d1 = [(1 << i) | (1 << j) for i, j in edges]                       # rows: edges over vertices
            d2 = [(1 << eid[(a, b)]) | (1 << eid[(a, c)]) | (1 << eid[(b, c)])
                  for a, b, c in combinations(range(n), 3)
                  if (a, b) in eid and (a, c) in eid and (b, c) in eid]        # rows: triangles over edges
            betti1 = len(edges) - _gf2_rank(d1) - _gf2_rank(d2)
            alive = int(((h1[:, 0] <= t) & (h1[:, 1] > t)).sum())
            assert alive == betti1, (n, float(t), alive, betti1)


def _h1_multiset(dg):
    return sorted((np.float32(b), np.float32(d)) for b, d in dg)


def test_known_answers_from_theory():
    """Inputs whose diagrams follow from a proof, not from any implementation: small polytopes and
    evenly spaced points on a circle (Adamaszek-Adams).  Both CPU statements of the algorithm."""
    from oracle import rips_naive
    for name, D, thr, h0, h1 in inputs.known_answer_cases():
        for impl in (lambda M, t: rips.ripser(M, thresh=t, distance_matrix=True),
                     lambda M, t: rips_naive.rips_h01_naive(M, t)):
            r = impl(D, thr)
            d0 = r["dgms"][0]
            assert np.array_equal(np.sort(d0[np.isfinite(d0[:, 1]), 1]).astype(np.float32), np.array(h0, np.float32)), name
            assert np.isinf(d0[:, 1]).sum() == 1 and (d0[:, 0] == 0).all(), name
            assert _h1_multiset(r["dgms"][1]) == sorted(h1), name


@pytest.mark.parametrize("geometry", ["graph", "chord"])
@pytest.mark.parametrize("n", [5, 6, 7, 9, 12, 16, 31, 47, 64, 100])
def test_circle_points_known_answer(n, geometry):
    """H1 of n evenly spaced points on a circle is exactly one bar [1 hop, ceil(n/3) hops)."""
    D, b, d = inputs.cycle_metric(n, geometry)
    r = rips.ripser(D, thresh=np.inf, distance_matrix=True)
    d0, d1 = r["dgms"]
    assert d0.shape == (n, 2) and (d0[:-1, 1] == b).all() and np.isinf(d0[-1, 1])
    assert d1.shape == (1, 2) and np.float32(d1[0, 0]) == b and np.float32(d1[0, 1]) == d
    if n <= 16:
        a = rips_naive.rips_h01_naive(D, np.inf)
        assert np.array_equal(a["dgms"][1], d1) and np.array_equal(a["pairs"][1], r["pairs"][1])
