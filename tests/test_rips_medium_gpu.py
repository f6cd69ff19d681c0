"""GPU parity of the CTA-per-cloud Rips engine (64 < N <= 254, also exercised on small N)."""
import numpy as np
import pytest

from tests import inputs

pytestmark = pytest.mark.gpu


def _compare(D, thresh, npts=None, cap1=None):
    import torch
    from oracle import rips as orips
    from tda_eeg_audio_b200 import rips_h01_batched
    nt = None if npts is None else torch.from_numpy(np.asarray(npts, np.int32)).cuda()
    r = rips_h01_batched(torch.from_numpy(D).cuda(), thresh=thresh, cap1=cap1, npts=nt, engine="medium")
    torch.cuda.synchronize()
    g = {k: v.cpu().numpy() for k, v in r.items() if k != "ws"}
    for b in range(len(D)):
        n = D.shape[1] if npts is None else int(npts[b])
        c = orips.rips_h01_batched(np.ascontiguousarray(D[b:b + 1, :n, :n]), thresh, cap1=cap1)
        assert tuple(g["counts"][b]) == tuple(c["counts"][0]), (b, g["counts"][b], c["counts"][0], g["status"][b])
        n0, n1 = c["counts"][0]
        n1 = min(n1, c["bd1"].shape[1])
        assert np.array_equal(g["bd0"][b, :n0].view(np.uint32), c["bd0"][0, :n0].view(np.uint32)), b
        assert np.array_equal(g["pr0"][b, :n0], c["pr0"][0, :n0]), b
        assert np.array_equal(g["bd1"][b, :n1].view(np.uint32), c["bd1"][0, :n1].view(np.uint32)), b
        assert np.array_equal(g["pr1"][b, :n1], c["pr1"][0, :n1]), b
    return g


def _takens_like(rng, B, n):
    """noisy quasi-periodic trajectories in [0,1]^3 -> float32 Gram-trick distances"""
    from sklearn.metrics import pairwise_distances
    out = np.zeros((B, n, n), np.float32)
    for b in range(B):
        t = np.arange(n + 40) * (0.15 + 0.2 * rng.random())
        s = np.sin(t) + 0.5 * np.sin(2.3 * t + 1) + 0.3 * rng.standard_normal(n + 40)
        pc = np.c_[s[:n], s[7:n + 7], s[14:n + 14]]
        pc = (pc - pc.min(0)) / (pc.max(0) - pc.min(0))
        out[b] = pairwise_distances(pc).astype(np.float32)
    return out


@pytest.mark.parametrize("n", [3, 17, 47, 64, 65, 97, 124, 128])
def test_medium_mw4(cuda, n):
    rng = np.random.default_rng(n)
    D = _takens_like(rng, 12, n) if n > 16 else inputs.sym_uniform(rng, 12, n)
    _compare(D, 2.0)
    _compare(D, 0.3)


@pytest.mark.parametrize("n", [129, 180, 248, 254])
def test_medium_mw8(cuda, n):
    D = _takens_like(np.random.default_rng(n), 6, n)
    _compare(D, 2.0)


def test_medium_uniform_and_ties(cuda):
    rng = np.random.default_rng(3)
    D = inputs.sym_uniform(rng, 8, 47)          # > 64 simultaneous classes -> W=8 tier
    _compare(D, 2.0)
    for q in (2, 8, 64):
        _compare((np.round(D * q) / q).astype(np.float32), 2.0)
    D = inputs.sym_uniform(rng, 4, 100)
    _compare((np.round(D * 16) / 16).astype(np.float32), 2.0)


def test_medium_degenerate_and_ragged(cuda):
    n = 110
    D = np.zeros((4, n, n), np.float32)
    D[1] = 1.0
    D[2] = _takens_like(np.random.default_rng(0), 1, n)[0]
    D[3] = _takens_like(np.random.default_rng(1), 1, n)[0]
    D[3, :, 5] = D[3, :, 4]; D[3, 5, :] = D[3, 4, :]; D[3, 4, 5] = D[3, 5, 4] = 0       # duplicate point
    for b in range(4):
        np.fill_diagonal(D[b], 0)
    _compare(D, 2.0)
    _compare(D, 2.0, npts=[1, 2, 60, 110])


def test_medium_cap1(cuda):
    D = _takens_like(np.random.default_rng(5), 4, 120)
    g = _compare(D, 2.0, cap1=8)
    assert (g["status"] & 1).all()
