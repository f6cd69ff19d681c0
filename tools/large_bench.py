"""Throughput of the Rips engines on Takens clouds: the
grid-cooperative engine (rips_large) at N = 1000/1500/2000 and at the audio sizes.  Device-resident distance matrices, CUDA events.  One JSON line per case."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tda_eeg_audio_b200 import _lib, rips_h01_batched
from tda_eeg_audio_b200.takens import pairwise_distance_f32


def takens_clouds(B, n, tau=7, seed=0, dev="cuda"):
    """alpha-band-like envelope windows -> Takens (dim 3) -> min-max normalised clouds -> f32 distances"""
    g = torch.Generator(device=dev); g.manual_seed(seed)
    L = n + 2 * tau
    t = torch.arange(L, device=dev, dtype=torch.float64) / 250.0
    s = torch.zeros((B, L), dtype=torch.float64, device=dev)
    for k in range(6):
        f = 8 + 5 * torch.rand((B, 1), generator=g, device=dev, dtype=torch.float64)
        ph = 6.283 * torch.rand((B, 1), generator=g, device=dev, dtype=torch.float64)
        a = torch.rand((B, 1), generator=g, device=dev, dtype=torch.float64)
        s += a * torch.sin(6.283185307179586 * f * t + ph)
    s += 0.05 * torch.randn((B, L), generator=g, device=dev, dtype=torch.float64)
    pc = torch.stack([s[:, k * tau:k * tau + n] for k in range(3)], dim=2)
    mn = pc.amin(dim=1, keepdim=True); rg = pc.amax(dim=1, keepdim=True) - mn
    rg[rg == 0] = 1
    return pairwise_distance_f32(((pc - mn) / rg).contiguous())


def timed(fn, n=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


CASES = [("large", 64, 1000, 2.0), ("large", 64, 2000, 2.0), ("large", 64, 2000, 0.25),
         ("large", 296, 1000, 2.0), ("large", 592, 1000, 2.0), ("large", 1184, 1000, 2.0),
         ("large", 148, 2000, 2.0), ("large", 296, 2000, 2.0), ("large", 592, 2000, 2.0),
         ("large", 8192, 124, 2.0), ("large", 8192, 248, 2.0), ("large", 32768, 124, 2.0), ("large", 32768, 248, 2.0)]


def main():
    cases = CASES
    if len(sys.argv) > 1:   # point counts to run, e.g. `large_bench.py 124 248`
        want = {int(a) for a in sys.argv[1:]}
        cases = [c for c in cases if c[2] in want]
    for engine, B, n, thr in cases:
        D = takens_clouds(B, n)
        out = {}
        cap1 = 4 * n
        ms = timed(lambda: rips_h01_batched(D, thr, cap1=cap1, want_pairs=True, out=out, engine=engine))
        parts = {}
        if engine == "large":
            _lib.profile_enable(True)
            rips_h01_batched(D, thr, cap1=cap1, want_pairs=True, out=out, engine=engine)
            torch.cuda.synchronize()
            for k in ("rank", "kruskal", "classify", "classify_tied", "sweep_t0", "sweep_t1", "sweep_t2"):
                parts[k] = round(_lib.profile_query("rips_large_" + k)[0], 3)
            _lib.profile_enable(False)
        c = out["counts"]
        print(json.dumps({"engine": engine, "B": B, "N": n, "thresh": thr, "ms": round(ms, 3),
                          "clouds_per_s": round(B / ms * 1e3, 1), "mean_h1": float(c[:, 1].float().mean()),
                          "max_h1": int(c[:, 1].max()), "status_nonzero": int((out["status"] != 0).sum()),
                          "kernel_ms": parts}), flush=True)
        del D, out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
