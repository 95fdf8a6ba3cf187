"""Stem step through the GENERAL QuantumConv1d path (n_qubits > 4 and / or angle embedding): conv1 + conv2 forward + backward at
batch 16 x 80 x 3000, per configuration, timed with CUDA events around the C-ABI calls (same buffers as bench.py).

    python tools/bench_general.py [--out profiles/r1_general_path.jsonl]
"""
import argparse, ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qasr_ijcnlp_b200 import QuantumConv1d, _lib

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=None)
ap.add_argument("--batch", type=int, default=16)
a = ap.parse_args()
dev = torch.device("cuda:0")
B = a.batch
rows = []
for q, Lq, emb in ((4, 1, "amplitude"), (4, 1, "angle"), (6, 1, "amplitude"), (8, 1, "amplitude"), (8, 2, "amplitude"), (10, 1, "amplitude")):
    torch.manual_seed(0)
    c1 = QuantumConv1d(80, 384, 3, padding=1, n_qubits=q, n_layers=Lq, embedding=emb).to(dev)
    c2 = QuantumConv1d(384, 384, 3, stride=2, padding=1, n_qubits=q, n_layers=Lq, embedding=emb).to(dev)
    x = torch.randn(B, 80, 3000, device=dev)
    def step():
        h = c1(x)
        h.retain_grad()
        y = c2(h)
        y.backward(torch.ones_like(y))
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    row = {"n_qubits": q, "n_layers": Lq, "embedding": emb, "batch": B, "path": "fused TMA kernels" if (q <= 4 and emb == "amplitude") else "general (composed)",
           "stem_fwd_bwd_ms": round(ms, 4), "windows_per_s": round(B * 4500 / (ms * 1e-3), 1)}
    rows.append(row)
    print(json.dumps(row), flush=True)
if a.out:
    with open(a.out, "w") as fh:
        for r in rows: fh.write(json.dumps(r) + "\n")
