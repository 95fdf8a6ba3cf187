import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib
from qasr_ijcnlp_b200 import _lib
qc = importlib.import_module("qasr_ijcnlp_b200.quantum_conv1d")
lib = _lib.load()
cuda = torch.device("cuda:0")
names = ["y", "gx", "gw_pre", "gb_pre", "gqw", "gw_post", "gb_post"]
for B in (int(v) for v in sys.argv[1].split(",")):
    for (C, S) in ((80, 1), (384, 2)):
        torch.manual_seed(21)
        m = qc.QuantumConv1d(C, 384, 3, stride=S, padding=1, n_qubits=4).to(cuda)
        x = torch.randn(B, C, 3000, device=cuda, requires_grad=True)
        res = []
        for fast in (1, 0):
            lib.qw_set_fast_path(fast)
            y = m(x)
            gy = torch.ones_like(y) * torch.linspace(-1, 1, y.shape[-1], device=cuda)
            grads = torch.autograd.grad(y, [x] + list(m.parameters()), gy)
            res.append([y.detach()] + [g for g in grads])
        lib.qw_set_fast_path(1)
        out = []
        for n, a, b in zip(names, *res):
            out.append("%s %.2e" % (n, (a - b).abs().max().item() / max(1.0, b.abs().max().item())))
        print("B", B, "C", C, "S", S, " ".join(out), flush=True)
