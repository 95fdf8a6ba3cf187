"""Tiny driver for ncu: a few eager stem steps (8 kernel launches each) at the bench workload."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda:0")
r = bench.StemRunner(a.batch, dev, 2)
for i in range(a.steps):
    r.step(i % 2)
torch.cuda.synchronize()
print("ok", a.steps, "steps")
