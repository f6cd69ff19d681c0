"""Per-stage throughput of the hot path on one GPU (device-resident inputs, CUDA events).
Prints one JSON line per stage; used for the tables in DESIGN.md / profiles/."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tda_eeg_audio_b200 import _lib, dsp, pipeline, takens, rips_h01_batched
from tda_eeg_audio_b200.wasserstein import wasserstein_batched

R = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = "cuda"
PEAK = 6530.0


def timed(fn, n=3, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


g = torch.Generator(device=dev); g.manual_seed(1)
C, T = 47, 15000
from tools.synth import raw_eeg_to_device
x = raw_eeg_to_device(0, R, dev)                     # BASELINE 5(b) generator
sos = np.stack([dsp.design_bandpass_filter(lo, hi, 250) for lo, hi in dsp.FREQ_BANDS.values()])
filt = torch.empty((5, R * C, T), dtype=torch.float64, device=dev)
ws = torch.empty((int(_lib.load().tda_filtfilt_workspace_bytes(R * C, 5, T, 27)),), dtype=torch.uint8, device=dev)
ms = timed(lambda: dsp.sosfiltfilt_batched(x.view(R * C, T), sos, out=filt, ws=ws))
samples = R * C * T * 5
print(json.dumps({"stage": "sosfiltfilt 5 bands", "ms": ms, "sample_bands_per_s": samples / ms * 1e3,
                  "alg_GBps": samples * 16 / ms / 1e6, "frac_hbm": samples * 16 / ms / 1e6 / PEAK}))
for step in (250, 62):
    W = dsp.n_windows(T, 250, step)
    D = torch.empty((R, 5, W, C, C), dtype=torch.float32, device=dev)
    def corr():
        for b in range(5):
            dsp.corrdist_windows(filt[b].view(R, C, T), 250, step, out=D[:, b], out_rec_stride=D.stride(0))
    ms = timed(corr)
    nwin = R * 5 * W
    print(json.dumps({"stage": f"corrdist step={step}", "ms": ms, "windows_per_s": nwin / ms * 1e3,
                      "fp64_TFLOPs": nwin * 2 * 47 * 47 * 250 / ms / 1e9}))
D = dsp.eeg_distances_from_raw(x, overlap=0.0)       # (R,5,60,47,47)
st = {}
ms = timed(lambda: pipeline.eeg_features_from_distances(D, state=st))
print(json.dumps({"stage": "rips47+features", "ms": ms, "diagrams_per_s": D.shape[0] * 300 / ms * 1e3,
                  "mean_h1": float(st["rips"]["counts"][:, 1].float().mean())}))
ms = timed(lambda: dsp.eeg_distances_from_raw(x, overlap=0.0, out=D))
print(json.dumps({"stage": "raw EEG -> distance matrices (5 bands, 60 windows)", "ms": ms,
                  "recordings_per_s": R / ms * 1e3}))
# ---- audio
env = torch.abs(torch.randn((R, T), generator=g, device=dev, dtype=torch.float64)) * \
    (1 + 0.6 * torch.sin(torch.arange(T, device=dev, dtype=torch.float64) * 2 * np.pi * 3.1 / 250))
for sub in (2, 1):
    aud = {}
    def audio():
        aud.update(pipeline.audio_diagrams_from_envelope(env, overlap=0.0, subsample=sub, max_windows=None, cap1=256))
    ms_all = timed(audio, n=2, warm=1)
    Dm, npts = aud["D"], aud["npts"]
    out = {}
    ms = timed(lambda: rips_h01_batched(Dm, thresh=2.0, cap1=256, want_pairs=False, npts=npts, engine="large", out=out), n=2, warm=1)
    print(json.dumps({"stage": f"audio Takens sub={sub}", "clouds": Dm.shape[0], "npts_mean": float(npts.float().mean()),
                      "npts_max": int(npts.max()), "chain_ms": ms_all, "rips_large_ms": ms,
                      "clouds_per_s": Dm.shape[0] / ms * 1e3,
                      "mean_h1": float(out["counts"][:, 1].float().mean()),
                      "status_bad": int((out["status"] & 4).sum())}))
    if sub == 2:
        eeg = st["rips"]
        nB = min(eeg["counts"].shape[0], out["counts"].shape[0])
        e_bd0, e_c = eeg["bd0"][:nB], eeg["counts"][:nB]
        ms0 = timed(lambda: wasserstein_batched(e_bd0, e_c[:, 0], out["bd0"][:nB], out["counts"][:nB, 0]), n=2, warm=1)
        ms1 = timed(lambda: wasserstein_batched(eeg["bd1"][:nB], e_c[:, 1], out["bd1"][:nB], out["counts"][:nB, 1]), n=2, warm=1)
        print(json.dumps({"stage": "wasserstein", "pairs": nB, "H0_ms": ms0, "H0_pairs_per_s": nB / ms0 * 1e3,
                          "H1_ms": ms1, "H1_pairs_per_s": nB / ms1 * 1e3}))
# audio front end
from tda_eeg_audio_b200 import audio as _audio
xa = torch.randn((min(R, 64), 2646000), generator=g, device=dev, dtype=torch.float64)
ms = timed(lambda: _audio.resample_poly_batched(xa, 250, 44100), n=3, warm=1)
print(json.dumps({"stage": "resample_poly 44.1k->250", "recordings": xa.shape[0], "ms": ms,
                  "fp64_TFLOPs": xa.shape[0] * 15000 * 3529 * 2 / ms / 1e9, "input_GBps": xa.numel() * 8 / ms / 1e6}))
ya = _audio.resample_poly_batched(xa, 250, 44100)
ms = timed(lambda: _audio.compute_envelope_batched(ya, 250), n=3, warm=1)
print(json.dumps({"stage": "hilbert envelope + LP50", "recordings": xa.shape[0], "ms": ms}))
del xa
# CPU baselines for the audio clouds (bounded sample)
try:
    from oracle import rips as orips
    n = 256
    Dh = Dm[:n].cpu().numpy(); nh = npts[:n].cpu().numpy()
    t0 = time.perf_counter()
    for k in range(n):
        orips.rips_h01_batched(np.ascontiguousarray(Dh[k:k + 1, :nh[k], :nh[k]]), 2.0, nthreads=1)
    dt = time.perf_counter() - t0
    print(json.dumps({"stage": "cpu oracle audio sub=1, 1 thread", "clouds_per_s": n / dt, "cores": os.cpu_count()}))
except Exception as e:
    print("cpu baseline failed", e)
