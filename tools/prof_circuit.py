"""Tiny driver for ncu: circuit forward + adjoint backward at n_qubits = 10 (and 8), 2 layers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qasr_ijcnlp_b200 import quantum_circuit
for q, W in ((10, 1 << 16), (8, 1 << 18)):
    pre = torch.randn(W, q, device="cuda:0", requires_grad=True)
    w = torch.randn(2, q, 3, device="cuda:0", requires_grad=True)
    for _ in range(2):
        out = quantum_circuit(pre, w, n_layers=2)
        out.backward(torch.ones_like(out))
torch.cuda.synchronize()
print("ok")
