"""Small fixed workload for ncu: the grid-cooperative Rips engine on Takens clouds."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.large_bench import takens_clouds
from tda_eeg_audio_b200 import rips_h01_batched

B, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 124)
D = takens_clouds(B, n)
out = {}
for _ in range(2):
    rips_h01_batched(D, 2.0, cap1=4 * n, want_pairs=False, out=out, engine="large")
torch.cuda.synchronize()
print("ok", B, n, float(out["counts"][:, 1].float().mean()))
