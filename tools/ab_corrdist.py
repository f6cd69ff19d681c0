"""A/B of the two generations of the window -> correlation -> distance kernel (default: FP64 FMA
register tiles, corrdist.cu; TDA_CORRDIST=mma: Gram on the FP64 tensor pipe, corrdist_mma.cu — staged,
see its header) in one process: time per call, largest differences between their outputs, and the
exact cases (duplicate channel -> r = 1, d = 0; zero-variance channel -> r = 0).  One JSON line each.
Parity against numpy for the staged kernel:  TDA_CORRDIST=mma python -m pytest tests/test_dsp_gpu.py -m gpu"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tda_eeg_audio_b200 import dsp

R = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = "cuda"
C, T = 47, 15000
g = torch.Generator(device=dev)
g.manual_seed(5)
A = torch.randn((R, C, 8), generator=g, device=dev, dtype=torch.float64) / 8 ** 0.5
x = A @ torch.randn((R, 8, T), generator=g, device=dev, dtype=torch.float64) + \
    0.5 * torch.randn((R, C, T), generator=g, device=dev, dtype=torch.float64)
x[:, 9] = x[:, 8]          # duplicate channel: r = 1 exactly, d = 0
x[:, 5] = 2.5              # zero variance: NaN -> r = 0, d = sqrt 2
x += 3.0                   # a DC offset the centring has to remove


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


res = {}
for step in (250, 62):
    W = dsp.n_windows(T, 250, step)
    for gen in ("fma", "mma"):
        os.environ["TDA_CORRDIST"] = gen
        D = torch.zeros((R, W, C, C), dtype=torch.float32, device=dev)
        ms = timed(lambda: dsp.corrdist_windows(x, 250, step, out=D))
        Dc, corr = dsp.corrdist_windows(x, 250, step, want_corr=True)
        res[gen] = (D.clone(), corr)
        nwin = R * W
        print(json.dumps({"step": step, "kernel": gen, "ms": round(ms, 3), "windows_per_s": round(nwin / ms * 1e3),
                          "fp64_TFLOPs": round(nwin * 2 * C * C * 250 / ms / 1e9, 2)}), flush=True)
    (D0, c0), (D1, c1) = res["fma"], res["mma"]
    off = ~torch.eye(C, dtype=torch.bool, device=dev)
    rel = ((D0 - D1).abs() / D0.clamp_min(1e-30))[..., off]
    print(json.dumps({"step": step, "max_abs_diff_corr": float((c0 - c1).abs().max()),
                      "max_rel_diff_dist": float(rel[torch.isfinite(rel)].max()),
                      "dist_bits_differ": int((D0.view(torch.int32) != D1.view(torch.int32)).sum()),
                      "duplicate_channel_r_is_1": bool((c1[..., 8, 9] == 1.0).all()),
                      "duplicate_channel_d_is_0": bool((D1[..., 8, 9] == 0).all()),
                      "zero_variance_r_is_0": bool((c1[..., 5, :] == 0).all())}), flush=True)
