"""A/B of the one-word 47-point tier's variants (TDA_RIPS_OPT, csrc/rips_small.cu) in one process:
throughput on EEG-like windows and bit-equality of every output against variant 0, on EEG-like,
uniform-random (all tiers) and tie-heavy batches.  One JSON line per variant."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from tda_eeg_audio_b200 import rips_h01_batched
from tests import inputs
from tools.synth import eeg_like_distance_matrices

B = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
variants = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [8, 10, 11, 12]
steps = 5


def fresh_out(Bn, n, cap1):
    z = lambda shape, dt: torch.zeros(shape, dtype=dt, device="cuda")
    return {"bd0": z((Bn, n, 2), torch.float32), "bd1": z((Bn, cap1, 2), torch.float32),
            "pr0": z((Bn, n, 2), torch.int64), "pr1": z((Bn, cap1, 2), torch.int64),
            "counts": z((Bn, 2), torch.int32), "status": z((Bn,), torch.int32)}


def run(D, thresh, cap1):
    out = fresh_out(D.shape[0], D.shape[1], cap1)
    rips_h01_batched(D, thresh, cap1=cap1, want_pairs=True, out=out)
    torch.cuda.synchronize()
    return {k: out[k].clone() for k in ("bd0", "bd1", "pr0", "pr1", "counts", "status")}


def same(a, b):
    # bit-equality (float rows compared as integers so that inf/NaN patterns count too)
    return all(torch.equal(a[k].view(torch.int32) if a[k].dtype == torch.float32 else a[k],
                           b[k].view(torch.int32) if b[k].dtype == torch.float32 else b[k]) for k in a)


Deeg = eeg_like_distance_matrices(B)
rng = np.random.default_rng(77)
U = inputs.sym_uniform(rng, 512, 47)
side = {
    "uniform": (torch.from_numpy(U).cuda(), 2.0),
    "ties64": (torch.from_numpy(np.round(U * 64) / 64).float().cuda(), 2.0),
    "ties8": (torch.from_numpy(np.round(U * 8) / 8).float().cuda(), 2.0),
    "thresh": (torch.from_numpy(U).cuda(), 0.4),
    "eeg_thresh": (Deeg[:2048].clone(), 1.2),
}
ref = {}
for v in variants:
    os.environ["TDA_RIPS_OPT"] = str(v)
    res = {"opt": v, "B": B}
    out = {}
    for _ in range(3):
        rips_h01_batched(Deeg, 2.0, cap1=128, want_pairs=False, out=out)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(steps):
        rips_h01_batched(Deeg, 2.0, cap1=128, want_pairs=False, out=out)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / steps
    res["ms"] = round(ms, 3)
    res["diagrams_per_s"] = round(B / ms * 1e3)
    res["tier_overflow"] = out["ws"][:12].view(torch.int32).tolist()
    got = {"eeg": run(Deeg[:65536], 2.0, 128)}
    for name, (D, th) in side.items():
        got[name] = run(D, th, 1035)
    if not ref:
        ref = got
        res["equal_to_first"] = None
    else:
        res["equal_to_first"] = {k: same(ref[k], got[k]) for k in got}
    res["status_nonzero"] = {k: int((got[k]["status"] != 0).sum()) for k in got}
    print(json.dumps(res), flush=True)
