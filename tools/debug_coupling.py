import sys, os, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import inputs
from oracle import signal_ref, rips as orips, wasserstein_ref
from tda_eeg_audio_b200 import drivers, pipeline, audio as _audio, dsp, storage
from tda_eeg_audio_b200.drivers import _eeg_rips, _cuda, _diagram_lists
td = tempfile.mkdtemp()
mat, gdir = inputs.tiny_dataset(td)
a = _audio.load_audio(mat)
env_ref = signal_ref.compute_envelope(signal_ref.resample_audio(a), 250)
env = _audio.audio_envelope_from_raw(_cuda(a)[None])[0]
print("env maxdiff", np.abs(env.cpu().numpy() - env_ref).max())
band = sys.argv[1] if len(sys.argv) > 1 else "alpha"
lo, hi = dsp.FREQ_BANDS[band]
ab = signal_ref.bandpass_filter(env_ref, 250, lo, hi)
wins = signal_ref.create_windows(ab, 250, 62)
dm = np.load(gdir / f"{band}_distances.npy")
n_win = min(len(wins), dm.shape[0])
idx = np.linspace(0, n_win - 1, 15, dtype=int) if n_win > 15 else np.arange(n_win)
tau = signal_ref.compute_tau(wins[idx[0]], max_lag=125)
ares = pipeline.audio_diagrams_from_envelope(env[None], bands={band: (lo, hi)}, max_windows=None, window_idx=idx)
eres = _eeg_rips(dm[idx], 2.0)
w0, w1 = pipeline.cross_wasserstein(eres, ares["rips"])
al = _diagram_lists(ares["rips"]); el = _diagram_lists(eres)
print("tau", tau, int(ares["tau"][0, 0]))
for k, w in enumerate(idx):
    pc = signal_ref.takens_embedding(wins[w], 3, tau, 2)
    mn = pc.min(0); rg = pc.max(0) - mn; rg[rg == 0] = 1
    ra = orips.ripser((pc - mn) / rg, thresh=2.0)
    d = dm[w]; d = (d + d.T) / 2; np.fill_diagonal(d, 0); d = np.maximum(d, 0)
    re = orips.ripser(d, thresh=2.0, distance_matrix=True)
    def clean(x):
        x = x[np.isfinite(x).all(1)]
        return x if len(x) else np.array([[0.0, 0.0]])
    r1 = wasserstein_ref.wasserstein(clean(re["dgms"][1]), clean(ra["dgms"][1]))
    # cross checks: GPU diagrams through the CPU wasserstein
    g1 = wasserstein_ref.wasserstein(clean(el[k][1]), clean(al[k][1]))
    same_a = ra["dgms"][1].shape == al[k][1].shape and np.abs(ra["dgms"][1] - al[k][1]).max() if ra["dgms"][1].shape == al[k][1].shape else "shape"
    same_e = np.array_equal(re["dgms"][1], el[k][1])
    print(k, w, "ref", r1, "gpu", float(w1[k]), "gpu-dgms/cpu-W", g1, "audio dgm diff", same_a, "eeg same", same_e, "nA", len(al[k][1]), "nE", len(el[k][1]))
