import os, sys, importlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import qconv_oracle as qo
qc = importlib.import_module("qasr_ijcnlp_b200.quantum_conv1d")
from qasr_ijcnlp_b200 import _lib
dev = torch.device("cuda:0")
geoms = [(2, 80, 200, 3, 1, 1, 384, 4), (2, 384, 208, 3, 2, 1, 384, 4)]
for geom in geoms:
    B, C, L, K, S, P, O, q = geom
    params64 = qo.make_params(C, O, K, q, seed=5)
    g = torch.Generator().manual_seed(6)
    x64 = torch.randn(B, C, L, generator=g, dtype=torch.float64)
    Lo = qo.out_length(L, K, S, P)
    gy64 = torch.randn(B, O, Lo, generator=g, dtype=torch.float64)
    ref = qo.qconv1d_grads(x64.float().double(), [p.float().double() for p in params64], gy64.float().double(), K, S, P)
    x = x64.float().to(dev).requires_grad_(True)
    ps = [p.float().to(dev).requires_grad_(True) for p in params64]
    try:
        y = qc.quantum_conv1d(x, *ps, kernel_size=K, stride=S, padding=P)
        torch.cuda.synchronize()
        print(geom, "fwd ok, err", (y.detach().cpu().double() - ref["y"]).abs().max().item(), flush=True)
        grads = torch.autograd.grad(y, [x] + ps, gy64.float().to(dev))
        torch.cuda.synchronize()
        for n, gr in zip(["x", "w_pre", "b_pre", "qweights", "w_post", "b_post"], grads):
            r = ref[n]
            print("   grad", n, (gr.cpu().double() - r).abs().max().item() / max(1.0, r.abs().max().item()), flush=True)
    except Exception as e:
        print(geom, "FAILED:", str(e)[:300], flush=True)
        break
