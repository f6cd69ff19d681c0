"""GPU parity of the batched exact Wasserstein solver against persim's semantics on scipy's LSAP
(oracle/wasserstein_ref.py).  Tolerance 1e-9 relative (north_star: 1e-5)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rand_dgm(rng, n, scale=1.0):
    b = rng.random(n) * scale
    return np.c_[b, b + rng.random(n) * scale * 0.5].astype(np.float32)


def test_random_pairs(cuda):
    import torch
    from oracle import wasserstein_ref
    from tda_eeg_audio_b200.wasserstein import wasserstein_batched
    rng = np.random.default_rng(0)
    capA, capB, K = 47, 124, 64
    A = np.zeros((K, capA, 2), np.float32); B = np.zeros((K, capB, 2), np.float32)
    nA = rng.integers(0, capA + 1, K).astype(np.int32); nB = rng.integers(0, capB + 1, K).astype(np.int32)
    nA[:4] = [0, 0, 1, capA]; nB[:4] = [0, 5, 0, capB]
    for k in range(K):
        A[k, :nA[k]] = _rand_dgm(rng, nA[k]); B[k, :nB[k]] = _rand_dgm(rng, nB[k], 0.7)
    A[5, 0, 1] = np.inf                              # essential bar: ignored
    got = wasserstein_batched(torch.from_numpy(A).cuda(), torch.from_numpy(nA).cuda(),
                              torch.from_numpy(B).cuda(), torch.from_numpy(nB).cuda()).cpu().numpy()
    for k in range(K):
        want = wasserstein_ref.safe_wasserstein(A[k, :nA[k]].astype(np.float64), B[k, :nB[k]].astype(np.float64))
        assert abs(got[k] - want) <= 1e-9 * max(1.0, abs(want)), (k, got[k], want)


def test_h0_like_and_index_pairing(cuda):
    """H0 diagrams (all births 0) EEG 46 bars vs audio ~120 bars, matched and mismatched pairing."""
    import torch
    from oracle import wasserstein_ref
    from tda_eeg_audio_b200.wasserstein import wasserstein_batched
    rng = np.random.default_rng(1)
    A = np.zeros((6, 47, 2), np.float32); B = np.zeros((4, 124, 2), np.float32)
    for k in range(6):
        A[k, :46, 1] = np.sort(rng.random(46)) * 1.4; A[k, 46] = (0, np.inf)
    for k in range(4):
        B[k, :119, 1] = np.sort(rng.random(119)) * 0.2; B[k, 119] = (0, np.inf)
    nA = np.full(6, 47, np.int32); nB = np.full(4, 120, np.int32)
    ia = np.array([0, 1, 2, 3, 4, 5, 0, 0], np.int32); ib = np.array([0, 1, 2, 3, 0, 1, 3, 2], np.int32)
    got = wasserstein_batched(torch.from_numpy(A).cuda(), torch.from_numpy(nA).cuda(), torch.from_numpy(B).cuda(),
                              torch.from_numpy(nB).cuda(), torch.from_numpy(ia).cuda(),
                              torch.from_numpy(ib).cuda()).cpu().numpy()
    for k in range(len(ia)):
        want = wasserstein_ref.safe_wasserstein(A[ia[k], :47].astype(np.float64), B[ib[k], :120].astype(np.float64))
        assert abs(got[k] - want) <= 1e-9 * max(1.0, abs(want))


def test_properties_and_dropins(cuda):
    from tda_eeg_audio_b200.wasserstein import safe_wasserstein, wasserstein
    rng = np.random.default_rng(2)
    d1, d2 = _rand_dgm(rng, 30), _rand_dgm(rng, 12)
    assert wasserstein(d1, d1) == 0.0
    assert abs(wasserstein(d1, d2) - wasserstein(d2, d1)) < 1e-12
    lone = np.array([[0.2, 0.8]])
    assert abs(wasserstein(lone, np.zeros((0, 2))) - 0.6 / np.sqrt(2)) < 1e-7     # to the diagonal
    assert safe_wasserstein(np.array([[0.0, np.inf]]), np.array([[0.0, np.inf]])) == 0.0
    assert safe_wasserstein(np.zeros(3), d2) == safe_wasserstein(np.zeros((0, 2)), d2)   # bad shape -> [[0,0]]


def test_one_dimensional_pairs_take_the_dynamic_programme_and_agree(cuda):
    """H0-shaped pairs (all births 0, deaths sorted) go through the 1-D dynamic programme, anything
    else through the general solver: both against scipy's LSAP on persim's cost matrix, including
    coincident deaths, duplicates, a shuffled copy (general solver on the same multiset) and births
    that are equal but not zero."""
    import torch
    from oracle import wasserstein_ref
    from tda_eeg_audio_b200.wasserstein import wasserstein_batched
    rng = np.random.default_rng(5)
    K, capA, capB = 48, 47, 250
    A = np.zeros((K, capA, 2), np.float32); B = np.zeros((K, capB, 2), np.float32)
    nA = rng.integers(1, capA + 1, K).astype(np.int32); nB = rng.integers(1, capB + 1, K).astype(np.int32)
    for k in range(K):
        a = np.sort(rng.random(nA[k])).astype(np.float32) * 1.4
        b = np.sort(rng.random(nB[k])).astype(np.float32) * (0.2 if k % 2 else 1.4)
        if k % 5 == 0 and nB[k] >= nA[k]:
            b[: nA[k]] = a                      # coincident deaths: zero-cost matches, Gram-trick noise
            b.sort()
        if k % 7 == 0:
            a[: len(a) // 2] = a[0]             # duplicates
        A[k, :nA[k], 1] = a; B[k, :nB[k], 1] = b
        if k % 11 == 0:
            A[k, :nA[k], 0] = 0.25; B[k, :nB[k], 0] = 0.25; A[k, :nA[k], 1] += 0.25; B[k, :nB[k], 1] += 0.25
    tA, tnA, tB, tnB = (torch.from_numpy(x).cuda() for x in (A, nA, B, nB))
    got = wasserstein_batched(tA, tnA, tB, tnB).cpu().numpy()
    perm = [rng.permutation(nA[k]) for k in range(K)]
    A2 = A.copy()
    for k in range(K):
        A2[k, :nA[k]] = A[k, perm[k]]           # same multiset, unsorted -> the general solver
    got2 = wasserstein_batched(torch.from_numpy(A2).cuda(), tnA, tB, tnB).cpu().numpy()
    for k in range(K):
        want = wasserstein_ref.wasserstein(A[k, :nA[k]].astype(np.float64), B[k, :nB[k]].astype(np.float64))
        assert abs(got[k] - want) <= 1e-9 * max(1.0, abs(want)), (k, got[k], want)
        assert abs(got2[k] - want) <= 1e-9 * max(1.0, abs(want)), (k, got2[k], want)


def test_large_pairs_and_independence_of_the_batch_caps(cuda):
    """Two diagrams of a few hundred points each (H1 of big clouds; H0 of two 248-point clouds): shared memory
    is linear in the sizes of a pair, so these run like any other pair and must agree with scipy; and the
    result of a pair must not depend on the capacities (padding) of the batch it travels in."""
    import torch
    from oracle import wasserstein_ref
    from tda_eeg_audio_b200.wasserstein import wasserstein, wasserstein_batched
    rng = np.random.default_rng(7)
    # (a) general pairs: 300 vs 420 points (1 MB of costs) and an H0-like 247 vs 247 (the dynamic programme)
    K = 3
    capA, capB = 300, 420
    A = np.zeros((K, capA, 2), np.float32); B = np.zeros((K, capB, 2), np.float32)
    nA = np.array([300, 0, 247], np.int32); nB = np.array([420, 57, 247], np.int32)
    A[0] = _rand_dgm(rng, capA); B[0] = _rand_dgm(rng, capB, 0.8)
    B[1, :57] = _rand_dgm(rng, 57)
    A[2, :247, 1] = np.sort(rng.random(247)).astype(np.float32); B[2, :247, 1] = np.sort(rng.random(247) * 0.9).astype(np.float32)
    got = wasserstein_batched(torch.from_numpy(A).cuda(), torch.from_numpy(nA).cuda(),
                              torch.from_numpy(B).cuda(), torch.from_numpy(nB).cuda()).cpu().numpy()
    for k in range(K):
        want = wasserstein_ref.safe_wasserstein(A[k, :nA[k]].astype(np.float64), B[k, :nB[k]].astype(np.float64))
        assert abs(got[k] - want) <= 1e-9 * max(1.0, abs(want)), (k, got[k], want)
    # (b) bit-identical whatever the caps: the same small pair inside a batch with big caps and alone
    a = _rand_dgm(rng, 40); b = _rand_dgm(rng, 90, 0.6)
    A2 = np.zeros((1, 600, 2), np.float32); B2 = np.zeros((1, 600, 2), np.float32)
    A2[0, :40] = a; B2[0, :90] = b
    big = wasserstein_batched(torch.from_numpy(A2).cuda(), torch.tensor([40], dtype=torch.int32).cuda(),
                              torch.from_numpy(B2).cuda(), torch.tensor([90], dtype=torch.int32).cuda()).item()
    small = wasserstein_batched(torch.from_numpy(A2[:, :40].copy()).cuda(), torch.tensor([40], dtype=torch.int32).cuda(),
                                torch.from_numpy(B2[:, :90].copy()).cuda(), torch.tensor([90], dtype=torch.int32).cuda()).item()
    assert big == small
    # (c) the float64 drop-in call on two large arbitrary diagrams
    d1 = _rand_dgm(rng, 350).astype(np.float64) + 1e-9; d2 = _rand_dgm(rng, 280).astype(np.float64)
    want = wasserstein_ref.safe_wasserstein(d1, d2)
    assert abs(wasserstein(d1, d2) - want) <= 1e-9 * max(1.0, abs(want))
