"""GPU parity of the audio pre-Rips stages against the reference's own functions (golden fixtures
from /root/reference/scripts/utils.py) and their numpy/sklearn restatement."""
import os

import numpy as np
import pytest

from oracle import signal_ref

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "signal.npz"))


def test_tau_matches_reference_fixture(cuda):
    import torch
    from tda_eeg_audio_b200 import takens
    for b, name in enumerate(signal_ref.FREQ_BANDS):
        wins = signal_ref.create_windows(G[f"bp_{name}"], 250, 62)
        tau = takens.compute_tau_batched(torch.from_numpy(wins).cuda(), max_lag=125).cpu().numpy()
        assert list(tau) == list(G["taus"][b]), name
    w = G["windows_alpha"][0]
    assert takens.compute_tau(w) == signal_ref.compute_tau(w)                 # default max_lag = len // 4
    assert takens.compute_tau(np.ones(250)) == signal_ref.compute_tau(np.ones(250)) == 1   # flat signal
    ramp = np.arange(250.0) ** 2
    assert takens.compute_tau(ramp, 20) == signal_ref.compute_tau(ramp, 20)   # no crossing -> max_lag // 10


def test_takens_embedding_dropin(cuda):
    from tda_eeg_audio_b200 import takens
    w0 = G["windows_alpha"][0]
    assert np.array_equal(takens.takens_embedding(w0, 3, 7, 2), G["takens_tau7_sub2"])
    assert np.array_equal(takens.takens_embedding(w0, 3, 12, 1), G["takens_tau12_sub1"])
    assert takens.takens_embedding(w0, 3, 125, 1).shape == (0, 3)


def test_normalised_cloud_and_distances(cuda):
    import torch
    from sklearn.metrics import pairwise_distances
    from tda_eeg_audio_b200 import takens
    wins = G["windows_alpha"][:6].copy()
    wins[5] = 3.0                                            # constant window: range 0 -> 1
    taus = np.array([7, 7, 12, 3, 40, 5], np.int32)
    pts, npts = takens.takens_cloud_batched(torch.from_numpy(wins).cuda(), torch.from_numpy(taus).cuda(), 3, 2)
    D = takens.pairwise_distance_f32(pts, npts).cpu().numpy()
    pts, npts = pts.cpu().numpy(), npts.cpu().numpy()
    for b in range(6):
        ref = signal_ref.normalise_cloud(signal_ref.takens_embedding(wins[b], 3, int(taus[b]), 2))
        assert npts[b] == len(ref)
        np.testing.assert_allclose(pts[b, :npts[b]], ref, rtol=0, atol=1e-15)
        dref = pairwise_distances(ref).astype(np.float32)
        n = npts[b]
        # 1e-5 relative is the contract; the Gram trick is replayed so float32 values agree to 1 ulp
        np.testing.assert_allclose(D[b, :n, :n], dref, rtol=2e-7, atol=1e-9)
        assert (np.diag(D[b, :n, :n]) == 0).all()


def test_compute_audio_persistence_chain(cuda):
    """windows -> tau -> Takens -> normalise -> distances -> Rips, against the same chain on the CPU
    (reference functions restated in oracle/signal_ref.py + the Ripser-style oracle)."""
    import torch
    from oracle import rips as orips
    from tda_eeg_audio_b200 import rips_h01_batched, takens
    wins = G["windows_alpha"][:8]
    tw = torch.from_numpy(wins).cuda()
    tau = int(takens.compute_tau_batched(tw[:1], 125)[0].item())
    pts, npts = takens.takens_cloud_batched(tw, tau, 3, 2)
    D = takens.pairwise_distance_f32(pts, npts)
    r = rips_h01_batched(D, thresh=2.0, npts=npts, cap1=256)
    counts = r["counts"].cpu().numpy()
    for b in range(len(wins)):
        pc = signal_ref.normalise_cloud(signal_ref.takens_embedding(wins[b], 3, signal_ref.compute_tau(wins[0], 125), 2))
        ref = orips.ripser(pc, maxdim=1, thresh=2.0)
        n0, n1 = counts[b]
        assert (n0, n1) == (len(ref["dgms"][0]), len(ref["dgms"][1]))
        # values within 1e-5 relative (float32 rounding of the distance can differ in the last bit)
        np.testing.assert_allclose(r["bd0"][b, :n0].cpu().numpy(), ref["dgms"][0], rtol=1e-5)
        np.testing.assert_allclose(r["bd1"][b, :n1].cpu().numpy(), ref["dgms"][1], rtol=1e-5)
