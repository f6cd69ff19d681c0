"""End-to-end EEG-audio coupling on two tiny synthetic recordings: the batched GPU drivers against
the reference's per-window Python loop restated on the CPU oracle
(process_recording, /root/reference/scripts/tda_eeg_audio_comparison.py:57-122;
compute_cross_wasserstein, /root/reference/scripts/matched_vs_mismatched.py:87-95)."""
import numpy as np
import pytest

from oracle import signal_ref

pytestmark = pytest.mark.gpu


def _cpu_chain(env, eeg_D, max_windows=15):
    from oracle import rips as orips, wasserstein_ref
    out = {}
    for bi, (name, (lo, hi)) in enumerate(signal_ref.FREQ_BANDS.items()):
        band = signal_ref.bandpass_filter(env, 250, lo, hi)
        wins = signal_ref.create_windows(band, 250, 62)
        n_win = min(len(wins), eeg_D.shape[1])
        idx = np.linspace(0, n_win - 1, max_windows, dtype=int) if n_win > max_windows else np.arange(n_win)
        tau = signal_ref.compute_tau(wins[idx[0]], max_lag=125)
        w0, w1 = [], []
        for w in idx:
            pc = signal_ref.takens_embedding(wins[w], 3, tau, 2)
            a = orips.ripser(signal_ref.normalise_cloud(pc), maxdim=1, thresh=2.0)["dgms"]
            dm = eeg_D[bi, w].astype(np.float64)
            dm = np.maximum((dm + dm.T) / 2, 0)
            np.fill_diagonal(dm, 0)
            e = orips.ripser(dm, maxdim=1, thresh=2.0, distance_matrix=True)["dgms"]
            w0.append(wasserstein_ref.safe_wasserstein(e[0], a[0]))
            w1.append(wasserstein_ref.safe_wasserstein(e[1], a[1]))
        out[name] = (tau, np.array(w0), np.array(w1), idx)
    return out


def test_process_recording_equivalent(cuda):
    import torch
    from tda_eeg_audio_b200 import dsp, pipeline, rips_h01_batched
    rng = np.random.default_rng(5)
    R, T = 2, 1500
    eeg = np.stack([signal_ref.eeg_like_recording(rng, T=T) for _ in range(R)])
    env = np.abs(rng.standard_normal((R, T))) * (1 + 0.6 * np.sin(np.arange(T) * 2 * np.pi * 3.1 / 250))
    eeg_D = dsp.eeg_distances_from_raw(torch.from_numpy(eeg).cuda())            # (R, 5, W, 47, 47)
    aud = pipeline.audio_diagrams_from_envelope(torch.from_numpy(env).cuda(), max_windows=15)
    idx = aud["idx"]
    Rr, nb, ns = aud["shape"]
    sel = eeg_D[:, :, torch.from_numpy(idx).cuda()].contiguous()               # (R, 5, n_sel, 47, 47)
    er = rips_h01_batched(sel.view(-1, 47, 47), thresh=2.0, cap1=128, want_pairs=False)
    w0, w1 = pipeline.cross_wasserstein(er, aud["rips"])
    w0 = w0.view(Rr, nb, ns).cpu().numpy(); w1 = w1.view(Rr, nb, ns).cpu().numpy()
    tau = aud["tau"].cpu().numpy()
    eD = eeg_D.cpu().numpy()
    for r in range(R):
        ref = _cpu_chain(env[r], eD[r])
        for bi, name in enumerate(signal_ref.FREQ_BANDS):
            t, r0, r1, ridx = ref[name]
            assert t == tau[r, bi] and np.array_equal(ridx, idx)
            np.testing.assert_allclose(w0[r, bi], r0, rtol=1e-5)      # north_star tolerance
            np.testing.assert_allclose(w1[r, bi], r1, rtol=1e-5, atol=1e-9)
            assert abs(np.nanmean(w1[r, bi]) - np.nanmean(r1)) <= 1e-5 * abs(np.nanmean(r1))


def test_mismatch_map():
    from tda_eeg_audio_b200.pipeline import mismatch_reference_recording
    m = mismatch_reference_recording(200)
    assert m[0] == 45 and m[45] == 0 and m[90] == 45 and m[135] == 0 and m[46] == 1
    assert (mismatch_reference_recording(40) == -1).all()
