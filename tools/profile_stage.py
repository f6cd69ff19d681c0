"""Small fixed workloads for ncu, one stage of the hot path per invocation:

    python tools/profile_stage.py <iir|corrdist|rips_small|features|wasserstein|takens|resample|rips_large> [size]

Each runs its kernel a few times on device-resident inputs of the benchmark's shape (the stated EEG
generator of tools/synth.py where the stage consumes EEG)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from tda_eeg_audio_b200 import _lib, dsp, pipeline, takens
from tools import synth

stage = sys.argv[1]
size = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = "cuda"
C, T = 47, 15000

if stage in ("iir", "corrdist"):
    R = size or 128
    x = synth.raw_eeg_to_device(0, R, dev)
    sos = np.stack([dsp.design_bandpass_filter(lo, hi, 250) for lo, hi in dsp.FREQ_BANDS.values()])
    filt = torch.empty((5, R * C, T), dtype=torch.float64, device=dev)
    ws = torch.empty((int(_lib.load().tda_filtfilt_workspace_bytes(R * C, 5, T, 27)),), dtype=torch.uint8, device=dev)
    for _ in range(2 if stage == "iir" else 1):
        dsp.sosfiltfilt_batched(x.view(R * C, T), sos, out=filt, ws=ws)
    if stage == "corrdist":
        D = torch.empty((R, 60, C, C), dtype=torch.float32, device=dev)
        for _ in range(2):
            dsp.corrdist_windows(filt[2].view(R, C, T), 250, 250, out=D)
elif stage in ("rips_small", "features"):
    R = size or 200
    D, _ = synth.eeg_distance_matrices(0, R, dev)
    st = {}
    for _ in range(3):
        pipeline.eeg_features_from_distances(D, thresh=2.0, cap1=128, state=st)
elif stage in ("wasserstein", "takens", "rips_large"):
    R = size or 64
    env = torch.abs(torch.randn((R, T), device=dev, dtype=torch.float64)) * \
        (1 + 0.6 * torch.sin(torch.arange(T, device=dev, dtype=torch.float64) * 2 * np.pi * 3.1 / 250))
    aud = pipeline.audio_diagrams_from_envelope(env, overlap=0.0, subsample=2, max_windows=None, cap1=256)
    if stage == "wasserstein":
        from tda_eeg_audio_b200.wasserstein import wasserstein_batched
        D, _ = synth.eeg_distance_matrices(0, R, dev)
        eeg = pipeline.eeg_features_from_distances(D, thresh=2.0, cap1=128)["rips"]
        a = aud["rips"]
        for _ in range(2):
            wasserstein_batched(eeg["bd0"], eeg["counts"][:, 0], a["bd0"], a["counts"][:, 0])
            wasserstein_batched(eeg["bd1"], eeg["counts"][:, 1], a["bd1"], a["counts"][:, 1])
    else:
        for _ in range(2):
            pipeline.audio_diagrams_from_envelope(env, overlap=0.0, subsample=2, max_windows=None, cap1=256)
elif stage == "resample":
    from tda_eeg_audio_b200 import audio as _audio
    xa = torch.randn((size or 32, 2646000), device=dev, dtype=torch.float64)
    for _ in range(2):
        ya = _audio.resample_poly_batched(xa, 250, 44100)
    _audio.compute_envelope_batched(ya, 250)
else:
    raise SystemExit(f"unknown stage {stage}")
torch.cuda.synchronize()
print("ok", stage)
