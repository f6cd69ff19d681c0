"""A/B of the generations of the zero-phase IIR pass kernel (TDA_IIR=0 / default / 3 = staged) in one
process: time per call and bit-equality of the outputs, sos form (EEG bands) and ba form (audio
bands), at two chunk sizes.  One JSON line per measurement."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from tda_eeg_audio_b200 import _lib, dsp

dev = "cuda"
GENS = sys.argv[1].split(",") if len(sys.argv) > 1 else ["0", "1", "3"]
C, T = 47, 15000
g = torch.Generator(device=dev)
g.manual_seed(3)


def timed(fn, n=3, warm=1):
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


sos = np.stack([dsp.design_bandpass_filter(lo, hi, 250) for lo, hi in dsp.FREQ_BANDS.values()])
for R in (256, 512):
    x = torch.randn((R * C, T), generator=g, device=dev, dtype=torch.float64)
    ws = torch.empty((int(_lib.load().tda_filtfilt_workspace_bytes(R * C, 5, T, 27)),), dtype=torch.uint8, device=dev)
    outs = {}
    for gen in GENS:
        os.environ["TDA_IIR"] = gen
        out = torch.empty((5, R * C, T), dtype=torch.float64, device=dev)
        ms = timed(lambda: dsp.sosfiltfilt_batched(x, sos, out=out, ws=ws))
        outs[gen] = out
        samples = R * C * T * 5
        print(json.dumps({"form": "sos", "recordings": R, "TDA_IIR": gen, "ms": round(ms, 3),
                          "alg_GBps": round(samples * 16 / ms / 1e6, 1),
                          "fp64_ops_per_s": round(samples * 72 / ms * 1e3 / 1e12, 3)}), flush=True)
    print(json.dumps({"form": "sos", "recordings": R,
                      "bit_equal_to_first": {g_: bool(torch.equal(outs[GENS[0]], outs[g_])) for g_ in GENS[1:]}}), flush=True)
    del x, ws, outs, out
    torch.cuda.empty_cache()
# ba form: the audio envelope bands (9+9 taps), ragged group count (last CTA partly filled)
ba = [dsp.design_bandpass_ba(lo, hi, 250) for lo, hi in dsp.FREQ_BANDS.values()]
env = torch.abs(torch.randn((1416 + 7, T), generator=g, device=dev, dtype=torch.float64))
outs = {}
for gen in GENS:
    os.environ["TDA_IIR"] = gen
    holder = {}
    ms = timed(lambda: holder.__setitem__("y", dsp.filtfilt_batched(env, ba)))
    outs[gen] = holder["y"]
    print(json.dumps({"form": "ba", "sequences": env.shape[0], "TDA_IIR": gen, "ms": round(ms, 3)}), flush=True)
print(json.dumps({"form": "ba", "bit_equal_to_first": {g_: bool(torch.equal(outs[GENS[0]], outs[g_])) for g_ in GENS[1:]}}),
      flush=True)
