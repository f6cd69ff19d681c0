"""Small fixed workload for ncu: a few passes of the Rips + feature pipeline on 59,200 windows."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.synth import eeg_like_distance_matrices
from tda_eeg_audio_b200 import pipeline

B = 148 * 400
D = eeg_like_distance_matrices(B).view(148, 5, 80, 47, 47)
state = {}
for _ in range(4):
    pipeline.eeg_features_from_distances(D, thresh=2.0, cap1=128, state=state)
torch.cuda.synchronize()
print("ok", B)
