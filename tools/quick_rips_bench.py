import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.synth import eeg_like_distance_matrices
from tda_eeg_audio_b200 import rips_h01_batched

B = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
D = eeg_like_distance_matrices(B)
out = {}
for cap1 in (1035, 64):
    for want in (True, False):
        out = {}
        for _ in range(3):
            rips_h01_batched(D, 2.0, cap1=cap1, want_pairs=want, out=out)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(5):
            rips_h01_batched(D, 2.0, cap1=cap1, want_pairs=want, out=out)
        ev[1].record(); torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 5
        c = out["counts"]
        print(f"B={B} cap1={cap1} pairs={want}: {ms:.3f} ms  {B/ms*1e3:.3e} diag/s  meanH1={c[:,1].float().mean().item():.2f} maxH1={c[:,1].max().item()} status_nonzero={(out['status']!=0).sum().item()}")
ws = out["ws"][:64].view(torch.int32)
print("tier overflow counters:", ws[:3].tolist())
