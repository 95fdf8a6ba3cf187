"""Tiny driver for ncu: a few qw_log_mel calls at batch 64 x 30 s."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qasr_ijcnlp_b200 import audio as qa
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
a = 0.1 * torch.randn(B, 480000, device="cuda:0")
for _ in range(3):
    m = qa.log_mel_spectrogram(a)
torch.cuda.synchronize()
print("ok", tuple(m.shape))
