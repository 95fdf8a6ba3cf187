"""Per-kernel CUDA-event times (qw_profile_*) of one stem step through the GENERAL QuantumConv1d path, batch 16 x 80 x 3000.

    python tools/prof_general.py [--q 6 8] [--batch 16]
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qasr_ijcnlp_b200 import QuantumConv1d, _lib

ap = argparse.ArgumentParser()
ap.add_argument("--q", type=int, nargs="+", default=[6, 8])
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--embedding", default="amplitude")
a = ap.parse_args()
dev = torch.device("cuda:0")
B = a.batch
for q in a.q:
    torch.manual_seed(0)
    c1 = QuantumConv1d(80, 384, 3, padding=1, n_qubits=q, embedding=a.embedding).to(dev)
    c2 = QuantumConv1d(384, 384, 3, stride=2, padding=1, n_qubits=q, embedding=a.embedding).to(dev)
    x = torch.randn(B, 80, 3000, device=dev)
    def step():
        h = c1(x)
        h.retain_grad()
        y = c2(h)
        y.backward(torch.ones_like(y))
    for _ in range(3): step()
    torch.cuda.synchronize()
    _lib.profile_read(True); _lib.profile_enable(True)
    for _ in range(5): step()
    torch.cuda.synchronize()
    _lib.profile_enable(False)
    prof = {k: (round(v[0] / v[1] * 1e3, 1), v[1] // 5) for k, v in _lib.profile_read(True).items()}
    tot = sum(us * n for us, n in prof.values())
    print(json.dumps({"n_qubits": q, "embedding": a.embedding, "batch": B, "sum_us_per_step": round(tot, 1),
                      "kernels_us_x_launches_per_step": prof}), flush=True)
