"""GPU parity: the CUDA Rips engine (through the C-ABI) against the CPU oracle on identical
float32 distance matrices.  Integer simplex pairs and float32 births/deaths must be BIT-EXACT."""
import numpy as np
import pytest

from tests import inputs

pytestmark = pytest.mark.gpu


def _run_gpu(D, thresh, cap1=None):
    import torch
    from tda_eeg_audio_b200 import rips_h01_batched
    r = rips_h01_batched(torch.from_numpy(D).cuda(), thresh=thresh, cap1=cap1)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in r.items() if k != "ws"}


def _compare(D, thresh, cap1=None):
    from oracle import rips
    g = _run_gpu(D, thresh, cap1)
    c = rips.rips_h01_batched(D, thresh, cap1=cap1)
    assert np.array_equal(g["counts"], c["counts"]), np.nonzero((g["counts"] != c["counts"]).any(1))[0][:10]
    B = len(D)
    cap = c["bd1"].shape[1]
    for b in range(B):
        n0, n1 = c["counts"][b]
        n1 = min(n1, cap)
        assert np.array_equal(g["bd0"][b, :n0].view(np.uint32), c["bd0"][b, :n0].view(np.uint32)), b
        assert np.array_equal(g["pr0"][b, :n0], c["pr0"][b, :n0]), b
        assert np.array_equal(g["bd1"][b, :n1].view(np.uint32), c["bd1"][b, :n1].view(np.uint32)), b
        assert np.array_equal(g["pr1"][b, :n1], c["pr1"][b, :n1]), b
    return g, c


def test_eeg_like_47(cuda):
    D = inputs.eeg_like(np.random.default_rng(0), 512)
    g, _ = _compare(D, 2.0)
    assert (g["status"] == 0).all()


def test_uniform_stress_47(cuda):
    # ~100 H1 bars, up to ~80 simultaneous classes: exercises the W=4 and W=64 tiers
    D = inputs.sym_uniform(np.random.default_rng(1), 256, 47)
    _compare(D, 2.0)


@pytest.mark.parametrize("q", [2, 8, 64])
def test_ties_47(cuda, q):
    D = inputs.sym_uniform(np.random.default_rng(2), 64, 47)
    D = (np.round(D * q) / q).astype(np.float32)
    _compare(D, 2.0)


@pytest.mark.parametrize("n", [2, 3, 4, 5, 16, 31, 32, 33, 47, 63, 64])
def test_sizes(cuda, n):
    rng = np.random.default_rng(10 + n)
    D = inputs.sym_uniform(rng, 48, n)
    _compare(D, np.inf)
    _compare(D, 0.5)
    Dq = (np.round(D * 8) / 8).astype(np.float32)
    _compare(Dq, np.inf)


def test_threshold_binding_and_essential_h1(cuda):
    D = inputs.circle_cloud(np.random.default_rng(5), 64, 40)
    g, c = _compare(D, 0.9)          # the big loop is still alive at 0.9 -> (b, inf) rows
    assert np.isinf(c["bd1"]).any()
    _compare(D, 3.0)


def test_degenerate(cuda):
    n = 47
    ones = np.ones((1, n, n), np.float32)
    zeros = np.zeros((1, n, n), np.float32)
    sq2 = np.full((1, n, n), np.sqrt(2), np.float32)
    D = np.concatenate([ones, zeros, sq2])
    for b in range(len(D)):
        np.fill_diagonal(D[b], 0)
    _compare(D, 2.0)
    _compare(D, 0.5)


def test_only_upper_triangle_is_read(cuda):
    D = inputs.eeg_like(np.random.default_rng(6), 16)
    g1 = _run_gpu(D, 2.0)
    D2 = D.copy()
    il = np.tril_indices(47, -1)
    D2[:, il[0], il[1]] = 123.0
    D2[:, np.arange(47), np.arange(47)] = -5.0
    g2 = _run_gpu(D2, 2.0)
    for k in ("bd0", "pr0", "counts"):
        assert np.array_equal(g1[k], g2[k])
    n1 = g1["counts"][:, 1]
    for b in range(16):
        assert np.array_equal(g1["bd1"][b, :n1[b]], g2["bd1"][b, :n1[b]])


def test_nan_edges_are_dropped(cuda):
    D = inputs.eeg_like(np.random.default_rng(8), 8)
    D[:, 3, 10] = np.nan
    D[:, 10, 3] = np.nan
    g, c = _compare(D, 2.0)
    assert (g["status"] & 2).all()


def test_cap1_truncation(cuda):
    D = inputs.sym_uniform(np.random.default_rng(9), 8, 47)
    g, c = _compare(D, 2.0, cap1=16)
    assert (g["status"] & 1).all() and (g["counts"][:, 1] > 16).all()


def test_host_entry_point(cuda):
    import ctypes
    from tda_eeg_audio_b200 import _lib
    from oracle import rips
    D = inputs.eeg_like(np.random.default_rng(11), 300)
    B, n, cap1 = len(D), 47, 64
    bd0 = np.zeros((B, n, 2), np.float32); pr0 = np.zeros((B, n, 2), np.int64)
    bd1 = np.zeros((B, cap1, 2), np.float32); pr1 = np.zeros((B, cap1, 2), np.int64)
    counts = np.zeros((B, 2), np.int32); status = np.zeros(B, np.int32)
    rc = _lib.load().tda_rips_h01_host(D.ctypes.data, B, n, 2.0, bd0.ctypes.data, pr0.ctypes.data,
                                       bd1.ctypes.data, pr1.ctypes.data, counts.ctypes.data, cap1,
                                       status.ctypes.data, 0)
    assert rc == 0
    c = rips.rips_h01_batched(D, 2.0, cap1=cap1)
    assert np.array_equal(counts, c["counts"])
    for b in range(B):
        n1 = counts[b, 1]
        assert np.array_equal(bd1[b, :n1], c["bd1"][b, :n1]) and np.array_equal(pr1[b, :n1], c["pr1"][b, :n1])
        assert np.array_equal(bd0[b], c["bd0"][b]) and np.array_equal(pr0[b], c["pr0"][b])


@pytest.mark.parametrize("n", [2, 3, 17, 33, 47, 64])
def test_condensed_input_equals_dense(cuda, n):
    """ld == 0: the condensed upper triangle (ripser's own FFI vector DParam) gives bit-identical
    outputs to the dense matrix, on every tier (uniform random matrices overflow the first ones),
    with ties and with a binding threshold"""
    import torch
    from tda_eeg_audio_b200 import rips_h01_batched
    from tda_eeg_audio_b200.rips import condense
    rng = np.random.default_rng(100 + n)
    U = inputs.sym_uniform(rng, 96, n)
    cases = [(U, 2.0), (np.round(U * 16).astype(np.float32) / 16, 2.0), (U, 0.45)]
    if n == 47:
        cases.append((inputs.eeg_like(rng, 256), 2.0))
    for D, th in cases:
        Dd = torch.from_numpy(D).cuda()
        Dc = condense(Dd)
        assert Dc.shape == (len(D), n * (n - 1) // 2)
        assert np.array_equal(condense(D), Dc.cpu().numpy())
        a = rips_h01_batched(Dd, thresh=th)
        b = rips_h01_batched(Dc, thresh=th, n_points=n)
        torch.cuda.synchronize()
        assert torch.equal(a["counts"], b["counts"]) and torch.equal(a["status"], b["status"])
        cnt = a["counts"].cpu().numpy()
        for k in ("bd0", "pr0", "bd1", "pr1"):
            x, y = a[k].cpu().numpy(), b[k].cpu().numpy()
            col = 0 if k.endswith("0") else 1
            for i in range(len(D)):
                m = min(cnt[i, col], x.shape[1])
                assert np.array_equal(x[i, :m].view(np.uint8), y[i, :m].view(np.uint8)), (k, i)


def test_condensed_host_entry_points(cuda):
    """tda_rips_h01_condensed_host / tda_eeg_features_condensed_host against their dense twins"""
    from tda_eeg_audio_b200 import _lib
    from tda_eeg_audio_b200.rips import condense
    lib = _lib.load()
    R, Bd, Wn, n, cap1 = 2, 5, 6, 47, 64
    B = R * Bd * Wn
    D = inputs.eeg_like(np.random.default_rng(12), B)
    Dc = condense(D)
    outs = []
    for fn, src in ((lib.tda_rips_h01_host, D), (lib.tda_rips_h01_condensed_host, Dc)):
        bd0 = np.zeros((B, n, 2), np.float32); pr0 = np.zeros((B, n, 2), np.int64)
        bd1 = np.zeros((B, cap1, 2), np.float32); pr1 = np.zeros((B, cap1, 2), np.int64)
        counts = np.zeros((B, 2), np.int32); status = np.zeros(B, np.int32)
        assert fn(src.ctypes.data, B, n, 2.0, bd0.ctypes.data, pr0.ctypes.data, bd1.ctypes.data,
                  pr1.ctypes.data, counts.ctypes.data, cap1, status.ctypes.data, 0) == 0
        outs.append((bd0, pr0, bd1, pr1, counts, status))
    (bd0a, pr0a, bd1a, pr1a, ca, sa), (bd0b, pr0b, bd1b, pr1b, cb, sb) = outs
    assert np.array_equal(ca, cb) and np.array_equal(sa, sb) and ca[:, 1].min() > 0
    for i in range(B):    # rows beyond the counts are not written
        n0, n1 = ca[i, 0], min(ca[i, 1], cap1)
        assert np.array_equal(bd0a[i, :n0].view(np.uint8), bd0b[i, :n0].view(np.uint8))
        assert np.array_equal(pr0a[i, :n0], pr0b[i, :n0])
        assert np.array_equal(bd1a[i, :n1].view(np.uint8), bd1b[i, :n1].view(np.uint8))
        assert np.array_equal(pr1a[i, :n1], pr1b[i, :n1])
    tabs = []
    for fn, src in ((lib.tda_eeg_features_host, D), (lib.tda_eeg_features_condensed_host, Dc)):
        table = np.zeros((R, Bd * 44)); feats = np.zeros((B, 2, 11))
        assert fn(src.ctypes.data, R, Bd, Wn, n, 2.0, cap1, None, None, None, None, feats.ctypes.data,
                  table.ctypes.data, 0) == 0
        tabs.append((table, feats))
    assert np.array_equal(tabs[0][0], tabs[1][0]) and np.array_equal(tabs[0][1], tabs[1][1])
    assert np.abs(tabs[0][0]).sum() > 0


def test_ripser_shim(cuda):
    from tda_eeg_audio_b200 import ripser
    from oracle import rips
    D = inputs.eeg_like(np.random.default_rng(12), 1)[0].astype(np.float64)
    a = ripser(D, maxdim=1, thresh=2.0, distance_matrix=True)
    b = rips.ripser(D, maxdim=1, thresh=2.0, distance_matrix=True)
    for k in range(2):
        assert a["dgms"][k].dtype == np.float64 and np.array_equal(a["dgms"][k], b["dgms"][k])
        assert np.array_equal(a["pairs"][k], b["pairs"][k])
    with pytest.raises(Exception, match="not square"):
        ripser(np.zeros((3, 4)), distance_matrix=True)


def test_h0_is_the_minimum_spanning_tree(cuda):
    """independent anchor at the EEG size: finite H0 deaths == scipy's MST weights, window by window"""
    import torch
    from scipy.sparse.csgraph import minimum_spanning_tree
    from tda_eeg_audio_b200 import rips_h01_batched
    D = inputs.eeg_like(np.random.default_rng(31), 400)
    r = rips_h01_batched(torch.from_numpy(D).cuda(), thresh=2.0, cap1=128, want_pairs=False)
    bd0 = r["bd0"].cpu().numpy(); cnt = r["counts"].cpu().numpy()
    for b in range(len(D)):
        mst = np.sort(minimum_spanning_tree(np.triu(D[b].astype(np.float64), 1)).data.astype(np.float32))
        assert cnt[b, 0] == 47 and np.array_equal(bd0[b, :46, 1], mst) and np.isinf(bd0[b, 46, 1])


def test_known_answers_from_theory(cuda):
    """Diagrams that follow from a proof (tests/inputs.py: small polytopes; evenly spaced points on
    a circle, Adamaszek-Adams): the engines against the theory directly, no oracle in between.
    Every engine that accepts the size; dense and condensed input."""
    import torch
    from tda_eeg_audio_b200 import rips_h01_batched
    from tda_eeg_audio_b200.rips import condense

    def check(name, D, thr, h0, h1, **kw):
        r = rips_h01_batched(D, thresh=float(thr), **kw)
        torch.cuda.synchronize()
        n0, n1 = (int(v) for v in r["counts"][0].tolist())
        assert int(r["status"][0]) == 0, name
        d0 = r["bd0"][0, :n0].cpu().numpy()
        d1 = r["bd1"][0, :n1].cpu().numpy()
        fin = np.isfinite(d0[:, 1])
        assert np.array_equal(np.sort(d0[fin, 1]), np.array(h0, np.float32)) and (~fin).sum() == 1 and (d0[:, 0] == 0).all(), name
        assert sorted(map(tuple, d1)) == sorted(h1), (name, d1)

    cases = inputs.known_answer_cases()
    for n in (5, 6, 7, 9, 12, 16, 31, 33, 47, 64, 100, 250):
        for geo in ("graph", "chord"):
            D, b, d = inputs.cycle_metric(n, geo)
            cases.append((f"circle{n}{geo}", D, np.inf, [b] * (n - 1), [(b, d)]))
    for name, D, thr, h0, h1 in cases:
        n = D.shape[0]
        Dg = torch.from_numpy(D).cuda()[None]
        if n <= 64:
            check(name, Dg, thr, h0, h1)
            check(name + "/condensed", condense(Dg), thr, h0, h1, n_points=n)
        check(name + "/large", Dg, thr, h0, h1, engine="large")


def test_host_entry_points_on_two_devices_in_one_process(cuda):
    """The *_host entry points cache staging buffers and streams per device ordinal: one process calling
    with device 0, then 1, then 0 again must get the same bits from each (skipped on a one-GPU box)."""
    import torch
    from tda_eeg_audio_b200 import _lib
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    D = inputs.eeg_like(np.random.default_rng(12), 2000)
    B, n, cap1 = len(D), 47, 64
    outs = []
    for dev in (0, 1, 0):
        bd0 = np.zeros((B, n, 2), np.float32); pr0 = np.zeros((B, n, 2), np.int64)
        bd1 = np.zeros((B, cap1, 2), np.float32); pr1 = np.zeros((B, cap1, 2), np.int64)
        counts = np.zeros((B, 2), np.int32); status = np.zeros(B, np.int32)
        rc = _lib.load().tda_rips_h01_host(D.ctypes.data, B, n, 2.0, bd0.ctypes.data, pr0.ctypes.data,
                                           bd1.ctypes.data, pr1.ctypes.data, counts.ctypes.data, cap1,
                                           status.ctypes.data, dev)
        assert rc == 0, (dev, rc)
        outs.append((bd0, pr0, bd1, pr1, counts, status))
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert np.array_equal(a, b)
    feats = []
    for dev in (1, 0):
        table = np.zeros((2, 5 * 44)); f = np.zeros((B, 2, 11))
        D5 = np.ascontiguousarray(D[:600].reshape(2, 5, 60, n, n))
        rc = _lib.load().tda_eeg_features_host(D5.ctypes.data, 2, 5, 60, n, 2.0, 128, None, None, None, None,
                                               f.ctypes.data, table.ctypes.data, dev)
        assert rc == 0, (dev, rc)
        feats.append(table)
    assert np.array_equal(feats[0], feats[1])
