"""Parity at scale on the STATED benchmark input (tools/synth.py): >= 100,000
band-passed 47-channel windows (all five bands, step 250 and step 62) plus >= 20,000 tie-heavy ones
through tda_rips_h01_batched with simplex pairs, compared BIT-EXACTLY with the CPU oracle, and the
number of windows each capacity tier of the engine finished.  One 1,000-point and one 2,000-point
Takens cloud against the oracle, pairs included (BASELINE configs[4] sizes)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CAP1 = 128


def _assert_equal_diagrams(g, c, what):
    """vectorised bit-exact comparison of padded diagram arrays (rows beyond the counts are ignored)"""
    assert np.array_equal(g["counts"], c["counts"]), (what, np.nonzero((g["counts"] != c["counts"]).any(1))[0][:10])
    for dim, (bd, pr) in enumerate((("bd0", "pr0"), ("bd1", "pr1"))):
        cap = c[bd].shape[1]
        live = np.arange(cap)[None, :] < np.minimum(c["counts"][:, dim], cap)[:, None]
        gb, cb = g[bd].view(np.uint32), c[bd].view(np.uint32)
        bad = ((gb != cb).any(2) | (g[pr] != c[pr]).any(2)) & live
        assert not bad.any(), (what, bd, np.nonzero(bad.any(1))[0][:10])


def _gpu_vs_oracle(D, what, chunk=32768):
    """D: CUDA float32 (B, 47, 47).  Returns the tier counts summed over the chunks."""
    import torch
    from oracle import rips as orips
    from tda_eeg_audio_b200 import rips_h01_batched
    from tda_eeg_audio_b200.rips import tier_counts
    tiers = {1: 0, 2: 0, 4: 0, 64: 0}
    for b0 in range(0, D.shape[0], chunk):
        Dc = D[b0:b0 + chunk]
        out = rips_h01_batched(Dc, thresh=2.0, cap1=CAP1, want_pairs=True)
        torch.cuda.synchronize()
        for k, v in tier_counts(out, 47).items():
            tiers[k] += v
        g = {k: v.cpu().numpy() for k, v in out.items() if k != "ws"}
        assert not (g["status"] & 4).any(), what
        c = orips.rips_h01_batched(Dc.cpu().numpy(), 2.0, cap1=CAP1)
        _assert_equal_diagrams(g, c, f"{what}[{b0}:]")
    return tiers


def test_stated_generator_100k_windows_bit_exact(cuda):
    import torch
    from tools.synth import eeg_distance_matrices
    total = 0
    tiers = {1: 0, 2: 0, 4: 0, 64: 0}
    for step, n_rec, rec0 in ((250, 200, 0), (62, 40, 5000)):
        D, _ = eeg_distance_matrices(rec0, n_rec, cuda, step=step)
        assert D.shape[1] == 5 and D.shape[2] == (60 if step == 250 else 238)
        t = _gpu_vs_oracle(D.view(-1, 47, 47), f"step{step}")
        for k in tiers:
            tiers[k] += t[k]
        total += D.shape[0] * D.shape[1] * D.shape[2]
        del D
        torch.cuda.empty_cache()
    assert total >= 100_000 and sum(tiers.values()) == total
    print("tier counts on the stated generator:", tiers, "of", total)
    # the one-word tier is sized for this input: it must finish nearly all of it, and what it hands
    # over must end in the next tiers with the same bits (checked above)
    assert tiers[1] >= 0.98 * total, tiers


def test_tie_heavy_20k_windows_bit_exact(cuda):
    """exact ties at scale: distances quantised to 1/64; a duplicated channel (r = 1 -> d = 0), a negated
    one (r = -1 -> d = 2) and a constant one (r := 0 -> a whole row of sqrt 2) in band-passed recordings"""
    import torch
    from tools.synth import raw_eeg_to_device
    from tda_eeg_audio_b200 import dsp
    x = raw_eeg_to_device(9000, 72, cuda)                                  # 72 x 5 x 60 = 21,600 windows
    D = dsp.eeg_distances_from_raw(x, overlap=0.0).view(-1, 47, 47)
    Dq = torch.round(D * 64) / 64
    tq = _gpu_vs_oracle(Dq, "quantised 1/64")
    # degenerate channels in the BAND-PASSED signal (a constant raw channel is rounding noise after the filter)
    import numpy as np
    sos = np.stack([dsp.design_bandpass_filter(lo, hi, 250) for lo, hi in dsp.FREQ_BANDS.values()])
    R, C, T = x.shape
    filt = dsp.sosfiltfilt_batched(x.view(R * C, T), sos).view(5, R, C, T)
    filt[:, :, 9] = filt[:, :, 8]
    filt[:, :, 20] = -filt[:, :, 3]
    filt[:, :, 5] = 2.5
    Dd = torch.stack([dsp.corrdist_windows(filt[b], 250, 250) for b in range(5)]).view(-1, 47, 47)
    # (numpy's own c / sd_i / sd_j of a duplicated channel is 1 or 1 - 2^-53: d = 0 or ~1.5e-8)
    assert float(Dd[:, 8, 9].max()) < 1e-7 and bool((Dd[:, 8, 9] == 0).any())
    assert bool((Dd[:, 5, 6] == Dd[:, 5, 30]).all()) and abs(float(Dd[0, 5, 6]) - 2 ** 0.5) < 1e-6
    assert float(Dd[:, 3, 20].min()) > 1.9999
    td = _gpu_vs_oracle(Dd, "duplicate / negated / constant channel")
    print("tier counts, quantised:", tq, " degenerate channels:", td)
    assert Dq.shape[0] + Dd.shape[0] >= 20_000


@pytest.mark.parametrize("n", [1000, 2000])
def test_stress_cloud_against_oracle_with_pairs(cuda, n):
    """BASELINE configs[4]: one Takens cloud of the stress sizes, pairs (simplex indices) and float32
    births / deaths bit-exact against the CPU oracle (seconds of oracle time per cloud)."""
    import torch
    from oracle import rips as orips
    from tda_eeg_audio_b200 import rips_h01_batched
    from tests.test_rips_large_gpu import takens_like
    D = takens_like(np.random.default_rng(100 + n), 1, n)
    cap1 = 16384
    r = rips_h01_batched(torch.from_numpy(D).cuda(), thresh=2.0, cap1=cap1, want_pairs=True, engine="large")
    torch.cuda.synchronize()
    g = {k: v.cpu().numpy() for k, v in r.items() if k != "ws"}
    assert int(g["status"][0]) == 0
    c = orips.rips_h01_batched(D, 2.0, cap1=cap1)
    _assert_equal_diagrams(g, c, f"cloud{n}")
    assert c["counts"][0, 1] > 50


def test_audio_sized_clouds_at_scale_bit_exact(cuda):
    """The audio path (BASELINE configs[2]: Takens clouds of 97-248 points) at scale through the grid-wide
    engine: 2,400 clouds with RAGGED point counts spread over the three shared-memory sort layouts (up to 128,
    170 and 256 points), a quarter of them with distances quantised to 1/256 (long tie runs, which the bit-row
    classification hands to the rank-row walk), pairs and float32 births / deaths bit-exact against the oracle."""
    import torch
    from oracle import rips as orips
    from tda_eeg_audio_b200 import rips_h01_batched
    from tests.test_rips_large_gpu import takens_like
    rng = np.random.default_rng(2026)
    total = 0
    for N, lo, B in ((128, 66, 900), (170, 129, 700), (256, 171, 800)):
        D = takens_like(rng, B, N)
        D[::4] = np.round(D[::4] * 256) / 256
        npts = rng.integers(lo, N + 1, B).astype(np.int32)
        npts[:3] = [N, lo, N - 1]
        cap1 = 1024
        out = rips_h01_batched(torch.from_numpy(D).cuda(), thresh=2.0, cap1=cap1, want_pairs=True,
                               npts=torch.from_numpy(npts).cuda(), engine="large")
        torch.cuda.synchronize()
        g = {k: v.cpu().numpy() for k, v in out.items() if k != "ws"}
        assert not g["status"].any(), (N, np.nonzero(g["status"])[0][:10])
        # the oracle takes one size per call: group the clouds by point count
        for n in np.unique(npts):
            sel = np.nonzero(npts == n)[0]
            c = orips.rips_h01_batched(np.ascontiguousarray(D[sel][:, :n, :n]), 2.0, cap1=cap1)
            gs = {"counts": g["counts"][sel], "bd0": g["bd0"][sel][:, :n], "pr0": g["pr0"][sel][:, :n],
                  "bd1": g["bd1"][sel], "pr1": g["pr1"][sel]}
            _assert_equal_diagrams(gs, c, f"N{N} n{n}")
        total += B
    assert total >= 2400
