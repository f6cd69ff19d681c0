"""CPU-only: hand-computed answers pin the Wasserstein oracle (persim semantics, SURVEY.md A.2: q = 1,
L2 ground metric between off-diagonal points, Euclidean distance |d - b| / sqrt 2 to the diagonal,
sum of the matched costs; an empty diagram counts as the single point (0, 0))."""
import itertools

import numpy as np

from oracle import wasserstein_ref as W

R2 = np.sqrt(2.0)


def test_hand_computed_cases():
    cases = [
        # one point each: moving it (cost 1) beats sending both to the diagonal (1/sqrt2 + 2/sqrt2)
        ([[0, 1]], [[0, 2]], 1.0),
        # far apart: the diagonal is cheaper (1/sqrt2 + 10/sqrt2 < 9)
        ([[0, 1]], [[0, 10]], 11 / R2),
        # the extra point goes to the diagonal
        ([[0, 2], [1, 3]], [[0, 2]], 2 / R2),
        # identical diagrams
        ([[0, 1], [0.5, 2.5], [1, 1.25]], [[0, 1], [0.5, 2.5], [1, 1.25]], 0.0),
        # against the empty diagram ((0,0) sits on the diagonal and costs nothing there)
        ([[0, 1]], np.zeros((0, 2)), 1 / R2),
        ([[0, 3], [1, 2]], [[0, 0]], 3 / R2 + 1 / R2),
        # a 2 x 2 assignment where the crossed matching wins: |(0,4)-(0,5)| + |(2,3)-(2,3.5)| = 1.5
        ([[0, 4], [2, 3]], [[2, 3.5], [0, 5]], 1.5),
        # rows with an infinite death are ignored (safe_wasserstein / persim drop them)
        ([[0, 1], [0, np.inf]], [[0, 2]], 1.0),
    ]
    for a, b, want in cases:
        got = W.safe_wasserstein(np.array(a, float).reshape(-1, 2), np.array(b, float).reshape(-1, 2))
        assert abs(got - want) < 1e-12, (a, b, got, want)
        got = W.safe_wasserstein(np.array(b, float).reshape(-1, 2), np.array(a, float).reshape(-1, 2))
        assert abs(got - want) < 1e-12, ("symmetry", a, b, got, want)


def test_brute_force_over_all_partial_matchings():
    """the definition itself on tiny diagrams: minimum over all partial matchings of matched L2 costs
    plus the diagonal costs of everything left unmatched"""
    rng = np.random.default_rng(5)
    for _ in range(40):
        m, n = rng.integers(1, 5), rng.integers(1, 5)
        A = np.sort(rng.random((m, 2)), axis=1)
        B = np.sort(rng.random((n, 2)), axis=1)
        best = np.inf
        for k in range(0, min(m, n) + 1):
            for ia in itertools.combinations(range(m), k):
                for ib in itertools.permutations(range(n), k):
                    c = sum(np.hypot(*(A[i] - B[j])) for i, j in zip(ia, ib))
                    c += sum((A[i, 1] - A[i, 0]) / R2 for i in range(m) if i not in ia)
                    c += sum((B[j, 1] - B[j, 0]) / R2 for j in range(n) if j not in ib)
                    best = min(best, c)
        assert abs(W.wasserstein(A, B) - best) < 1e-12
