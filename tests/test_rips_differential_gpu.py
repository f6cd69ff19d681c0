"""Randomised differential test: both Rips engines (small, large)
against the CPU oracle on the same seeded matrices — random metrics, point clouds, lattices
(massive ties), quantised values, binding thresholds, NaN edges.  Pairs bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cases(rng, n_cases, nmax):
    out = []
    for t in range(n_cases):
        n = int(rng.integers(2, nmax + 1))
        kind = t % 6
        if kind == 0:
            A = rng.random((n, n)); D = (A + A.T) / 2
        elif kind == 1:
            A = np.round(rng.random((n, n)) * 8) / 8; D = (A + A.T) / 2
        elif kind == 2:
            P = rng.random((n, 3)); D = np.sqrt(((P[:, None] - P[None]) ** 2).sum(-1))
        elif kind == 3:
            P = rng.integers(0, 4, (n, 2)).astype(float); D = np.sqrt(((P[:, None] - P[None]) ** 2).sum(-1))
        elif kind == 4:
            th = rng.random(n) * 2 * np.pi
            P = np.c_[np.cos(th), np.sin(th)] * (1 + 0.05 * rng.standard_normal((n, 1)))
            D = np.sqrt(((P[:, None] - P[None]) ** 2).sum(-1))
        else:
            A = rng.random((n, n)); D = (A + A.T) / 2
            k = rng.integers(0, n, 2)
            if k[0] != k[1]:
                D[k[0], k[1]] = D[k[1], k[0]] = np.nan
        np.fill_diagonal(D, 0)
        out.append((D.astype(np.float32), (np.inf, 0.6, 2.0)[t % 3]))
    return out


def _check(engine, D, thr, monkeypatch=None):
    import torch
    from oracle import rips as orips
    from tda_eeg_audio_b200 import rips_h01_batched
    n = D.shape[0]
    cap1 = max(n * (n - 1) // 2, 1)
    r = rips_h01_batched(torch.from_numpy(D[None]).cuda(), thresh=thr, cap1=cap1, engine=engine)
    c = orips.rips_h01_batched(D[None], thr, cap1=cap1)
    n0, n1 = c["counts"][0]
    assert tuple(r["counts"][0].tolist()) == (n0, n1), (engine, n, thr)
    assert np.array_equal(r["bd0"][0, :n0].cpu().numpy().view(np.uint32), c["bd0"][0, :n0].view(np.uint32))
    assert np.array_equal(r["pr0"][0, :n0].cpu().numpy(), c["pr0"][0, :n0])
    assert np.array_equal(r["bd1"][0, :n1].cpu().numpy().view(np.uint32), c["bd1"][0, :n1].view(np.uint32))
    assert np.array_equal(r["pr1"][0, :n1].cpu().numpy(), c["pr1"][0, :n1])


@pytest.mark.parametrize("engine", ["small", "large"])
def test_engines_against_oracle(cuda, engine):
    for D, thr in _cases(np.random.default_rng(77), 150, 40):
        _check(engine, D, thr)


@pytest.mark.parametrize("engine", ["large"])
def test_cloud_engines_mid_sizes(cuda, engine):
    for D, thr in _cases(np.random.default_rng(78), 24, 150):
        if D.shape[0] >= 3:
            _check(engine, D, thr)
