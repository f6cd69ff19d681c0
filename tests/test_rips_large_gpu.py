"""GPU parity of the grid-cooperative Rips engine (rips_large.cu, N <= 2048) against the CPU
oracle: persistence pairs (simplex indices) and float32 births/deaths bit-exact."""
import numpy as np
import pytest

from tests import inputs

pytestmark = pytest.mark.gpu


def _compare(D, thresh, npts=None, cap1=None, engine="large"):
    import torch
    from oracle import rips as orips
    from tda_eeg_audio_b200 import rips_h01_batched
    nt = None if npts is None else torch.from_numpy(np.asarray(npts, np.int32)).cuda()
    r = rips_h01_batched(torch.from_numpy(D).cuda(), thresh=thresh, cap1=cap1, npts=nt, engine=engine)
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


def takens_like(rng, B, n):
    """noisy quasi-periodic trajectories in [0,1]^3 -> float32 Gram-trick distances (what
    ripser(point_cloud) is fed, /root/reference/scripts/utils.py:123-131)"""
    from sklearn.metrics import pairwise_distances
    out = np.zeros((B, n, n), np.float32)
    for b in range(B):
        t = np.arange(n + 40) * (0.15 + 0.2 * rng.random())
        s = np.sin(t) + 0.5 * np.sin(2.3 * t + 1) + 0.3 * rng.standard_normal(n + 40)
        pc = np.c_[s[:n], s[7:n + 7], s[14:n + 14]]
        pc = (pc - pc.min(0)) / (pc.max(0) - pc.min(0))
        out[b] = pairwise_distances(pc).astype(np.float32)
    return out


@pytest.mark.parametrize("n", [2, 3, 17, 47, 65, 124, 250, 256])
def test_large_small_sizes(cuda, n):
    rng = np.random.default_rng(n)
    D = takens_like(rng, 6, n) if n > 16 else inputs.sym_uniform(rng, 6, n)
    _compare(D, 2.0)
    _compare(D, 0.3)


@pytest.mark.parametrize("n", [257, 400, 700, 1024])
def test_large_mid_sizes(cuda, n):
    D = takens_like(np.random.default_rng(n), 3, n)
    _compare(D, 2.0, cap1=4096)
    _compare(D, 0.25, cap1=4096)


@pytest.mark.parametrize("n", [1025, 1500])
def test_large_two_apexes_per_thread(cuda, n):
    D = takens_like(np.random.default_rng(n), 2, n)
    _compare(D, 2.0, cap1=8192)


def test_large_uniform_and_ties(cuda):
    rng = np.random.default_rng(3)
    D = inputs.sym_uniform(rng, 6, 47)
    _compare(D, 2.0)
    for q in (2, 8, 64):
        _compare((np.round(D * q) / q).astype(np.float32), 2.0)
    D = inputs.sym_uniform(rng, 3, 100)          # many simultaneous classes
    _compare(D, 2.0)
    _compare((np.round(D * 16) / 16).astype(np.float32), 2.0)
    D = inputs.sym_uniform(rng, 2, 200)          # ~770 simultaneous classes -> W=32 tier, slot recycling
    _compare(D, 2.0, cap1=50000)
    P = rng.integers(0, 6, (3, 150, 2)).astype(np.float64)   # lattice points: massive ties, duplicates
    D = np.sqrt(((P[:, :, None] - P[:, None]) ** 2).sum(-1)).astype(np.float32)
    _compare(D, 10.0)
    _compare(D, 2.5)


@pytest.mark.parametrize("n", [100, 200, 300])   # the three size classes of the sort / classification kernels
def test_large_degenerate_ragged_nan(cuda, n):
    D = np.zeros((5, n, n), np.float32)
    D[1] = 1.0
    D[2] = takens_like(np.random.default_rng(0), 1, n)[0]
    D[3] = takens_like(np.random.default_rng(1), 1, n)[0]
    D[3, :, 5] = D[3, :, 4]; D[3, 5, :] = D[3, 4, :]; D[3, 4, 5] = D[3, 5, 4] = 0       # duplicate point
    D[4] = takens_like(np.random.default_rng(2), 1, n)[0]
    for b in range(5):
        np.fill_diagonal(D[b], 0)
    _compare(D, 2.0, cap1=2048)
    _compare(D, 2.0, npts=[1, 2, 60, n, 3], cap1=2048)
    D[4, 3, n - 7] = np.nan
    D[4, 10, 11] = np.nan
    g = _compare(D, 2.0, cap1=2048)
    assert g["status"][4] & 2 and not (g["status"][:4] & 2).any()


def test_large_cap1_truncation(cuda):
    D = takens_like(np.random.default_rng(5), 3, 280)
    g = _compare(D, 2.0, cap1=8)
    assert (g["status"] & 1).all()


def test_auto_routes_large(cuda):
    D = takens_like(np.random.default_rng(9), 2, 320)
    _compare(D, 2.0, cap1=2048, engine="auto")


def test_full_size_properties_without_the_oracle(cuda):
    """BASELINE configs[4] sizes (2,000 points): properties that do not need the (slow) CPU oracle.
    (1) the finite H0 deaths are the weights of scipy's minimum spanning tree; (2) relabelling the
    points leaves the (birth, death) multisets unchanged; (3) every bar is born before it dies and
    births are edge lengths of the cloud."""
    import torch
    from scipy.sparse.csgraph import minimum_spanning_tree
    from tda_eeg_audio_b200 import rips_h01_batched
    n = 2000
    D = takens_like(np.random.default_rng(42), 1, n)[0]
    perm = np.random.default_rng(1).permutation(n)
    Dp = D[np.ix_(perm, perm)]
    r = rips_h01_batched(torch.from_numpy(np.stack([D, Dp])).cuda(), thresh=2.0, cap1=16384, want_pairs=False)
    cnt = r["counts"].cpu().numpy()
    assert (r["status"].cpu().numpy() == 0).all() and tuple(cnt[0]) == tuple(cnt[1])
    bd0 = r["bd0"].cpu().numpy(); bd1 = r["bd1"].cpu().numpy()
    h0 = bd0[0, :cnt[0, 0]]
    finite = np.isfinite(h0[:, 1])
    mst = np.sort(minimum_spanning_tree(np.triu(D.astype(np.float64), 1)).data.astype(np.float32))
    assert np.array_equal(np.sort(h0[finite, 1]), mst[mst > 0]) and (~finite).sum() == 1
    for dgm_a, dgm_b in ((bd0[0, :cnt[0, 0]], bd0[1, :cnt[1, 0]]), (bd1[0, :cnt[0, 1]], bd1[1, :cnt[1, 1]])):
        a = dgm_a[np.lexsort((dgm_a[:, 1], dgm_a[:, 0]))]
        b = dgm_b[np.lexsort((dgm_b[:, 1], dgm_b[:, 0]))]
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    h1 = bd1[0, :cnt[0, 1]]
    assert cnt[0, 1] > 100 and (h1[:, 0] < h1[:, 1]).all()
    lengths = np.unique(D[np.triu_indices(n, 1)])
    assert np.isin(h1[:, 0], lengths).all() and np.isin(h1[np.isfinite(h1[:, 1]), 1], lengths).all()


@pytest.mark.parametrize("n", [150, 300])
def test_large_batch_in_several_chunks(cuda, n):
    """A workspace that holds only a few clouds: the batch is worked off chunk by chunk (every per-chunk
    array, counter and list is re-used) and must give the very bits of the one-chunk run."""
    import torch
    from tda_eeg_audio_b200 import _lib, rips_h01_batched
    lib = _lib.load()
    B, cap1 = 11, 1024
    D = torch.from_numpy(takens_like(np.random.default_rng(40 + n), B, n)).cuda()
    D[::3] = torch.round(D[::3] * 128) / 128          # tie runs in some of the clouds
    ref = rips_h01_batched(D, thresh=2.0, cap1=cap1, want_pairs=True, engine="large")
    torch.cuda.synchronize()
    wsb = int(lib.tda_rips_h01_large_workspace_bytes(3, n))
    assert wsb < int(lib.tda_rips_h01_large_workspace_bytes(B, n))
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    bd0 = torch.zeros((B, n, 2), dtype=torch.float32, device="cuda"); pr0 = torch.zeros((B, n, 2), dtype=torch.int64, device="cuda")
    bd1 = torch.zeros((B, cap1, 2), dtype=torch.float32, device="cuda"); pr1 = torch.zeros((B, cap1, 2), dtype=torch.int64, device="cuda")
    counts = torch.zeros((B, 2), dtype=torch.int32, device="cuda"); status = torch.zeros(B, dtype=torch.int32, device="cuda")
    rc = lib.tda_rips_h01_large(D.data_ptr(), None, B, n, D.stride(1), D.stride(0), 2.0, bd0.data_ptr(), pr0.data_ptr(), n,
                                bd1.data_ptr(), pr1.data_ptr(), cap1, counts.data_ptr(), status.data_ptr(),
                                ws.data_ptr(), wsb, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert rc == 0
    assert torch.equal(counts, ref["counts"]) and not bool(status.any())
    for b in range(B):
        n0, n1 = (int(x) for x in counts[b])
        assert torch.equal(bd0[b, :n0], ref["bd0"][b, :n0]) and torch.equal(pr0[b, :n0], ref["pr0"][b, :n0])
        assert torch.equal(bd1[b, :n1], ref["bd1"][b, :n1]) and torch.equal(pr1[b, :n1], ref["pr1"][b, :n1])
