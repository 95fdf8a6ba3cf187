"""Per-kernel event timings of the stem layers with and without the fused GELU (forward epilogue / gy-pass prologue)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from qasr_ijcnlp_b200 import QuantumConv1d, _lib

dev = torch.device("cuda:0")
B = 16
torch.manual_seed(0)
for (C, S) in ((80, 1), (384, 2)):
    m = QuantumConv1d(C, 384, 3, stride=S, padding=1, n_qubits=4).to(dev)
    xs = [torch.randn(B, C, 3000, device=dev, requires_grad=(C == 384)) for _ in range(4)]
    Lo = 3000 if S == 1 else 1500
    gys = [torch.randn(B, 384, Lo, device=dev) for _ in range(4)]
    params = list(m.parameters())
    for fused in (False, True):
        def step(i):
            x = xs[i % 4]
            y = m.forward_gelu(x) if fused else m(x)
            torch.autograd.grad(y, ([x] if x.requires_grad else []) + params, gys[i % 4])
        for i in range(3):
            step(i)
        torch.cuda.synchronize()
        _lib.profile_read(True); _lib.profile_enable(True)
        for i in range(20):
            step(i)
        torch.cuda.synchronize()
        _lib.profile_enable(False)
        prof = _lib.profile_read(True)
        print(f"C={C} S={S} fused_gelu={fused}: " + ", ".join(f"{k.replace('qconv_', '')} {v[0] / v[1] * 1e3:.1f} us" for k, v in sorted(prof.items())))
# reference: ATen GELU forward / backward on the two tensor sizes
for shape in ((B, 384, 3000), (B, 384, 1500)):
    t = torch.randn(*shape, device=dev, requires_grad=True)
    g = torch.randn(*shape, device=dev)
    for _ in range(3):
        y = F.gelu(t); y.backward(g)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda.synchronize()
    tf = tb = 0.0
    for _ in range(10):
        e[0].record(); y = F.gelu(t); e[1].record(); y.backward(g); e[2].record(); torch.cuda.synchronize()
        tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
    print(f"ATen gelu {shape}: fwd {tf * 100:.1f} us, bwd {tb * 100:.1f} us (event brackets)")
# the whole training stem through stem_train_forward (one forward kernel at this batch + the chained backward), per launch in order
from qasr_ijcnlp_b200 import stem_train_forward
c1 = QuantumConv1d(80, 384, 3, padding=1, n_qubits=4).to(dev)
c2 = QuantumConv1d(384, 384, 3, stride=2, padding=1, n_qubits=4).to(dev)
xs = [torch.rand(B, 80, 3000, device=dev) * 3 - 1.5 for _ in range(4)]
gys = [torch.randn(B, 384, 1500, device=dev) for _ in range(4)]
prm = list(c1.parameters()) + list(c2.parameters())
def sstep(i):
    y = stem_train_forward(c1, c2, xs[i % 4], gelu=True)
    torch.autograd.grad(y, prm, gys[i % 4])
for i in range(3):
    sstep(i)
torch.cuda.synchronize()
_lib.profile_read(True); _lib.profile_enable(True)
for i in range(20):
    sstep(i)
torch.cuda.synchronize()
_lib.profile_enable(False)
prof = _lib.profile_read(True)
print("stem_train_forward (gelu) + chained backward, mean us per launch (launches per step): " +
      ", ".join(f"{k.replace('qconv_', '')} {v[0] / v[1] * 1e3:.1f} ({v[1] // 20})" for k, v in sorted(prof.items())) +
      f"; sum {sum(v[0] for v in prof.values()) / 20 * 1e3:.1f} us per step")
