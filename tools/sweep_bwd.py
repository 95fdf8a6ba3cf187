"""Time the backward kernels for arbitrary (B, C, L, S, O) through the C ABI with the in-library event timers."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qasr_ijcnlp_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
def p(t): return ctypes.c_void_p(t.data_ptr()) if t is not None else None
def run(B, C, L, S, O, need_gx, iters=20, nsets=3):
    K, P, Q = 3, 1, 4
    Lout = (L + 2 * P - K) // S + 1
    g = torch.Generator(device=dev).manual_seed(0)
    w_pre = torch.randn(Q, C * K, device=dev, generator=g) * 0.05
    b_pre = torch.randn(Q, device=dev, generator=g) * 0.05
    qw = torch.randn(Q, 3, device=dev, generator=g)
    w_post = torch.randn(O, Q, device=dev, generator=g) * 0.5
    b_post = torch.randn(O, device=dev, generator=g) * 0.5
    sets = []
    for s in range(nsets):
        x = torch.randn(B, C, L, device=dev, generator=g)
        sets.append(dict(x=x, y=torch.empty(B, O, Lout, device=dev), gy=torch.randn(B, O, Lout, device=dev, generator=g),
                         pre=torch.empty(2, B * Lout, Q, device=dev), gx=torch.empty_like(x) if need_gx else None))
    grads = [torch.empty_like(t) for t in (w_pre, b_pre, qw, w_post, b_post)]
    n = lib.qw_conv1d_workspace_bytes(B, C, L, K, S, P, O, Q, 1, 4)
    ws = torch.empty(n, device=dev, dtype=torch.uint8)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    dims = (B, C, L, K, S, P, O, Q, 1, 0)
    def fwd(t): _lib.check(lib.qw_conv1d_forward(p(t["x"]), p(w_pre), p(b_pre), p(qw), p(w_post), p(b_post), p(t["y"]), p(t["pre"]), *dims, st), "f")
    def bwd(t): _lib.check(lib.qw_conv1d_backward(p(t["gy"]), p(t["x"]), p(t["pre"]), p(w_pre), p(qw), p(w_post), p(t["gx"]), *[p(g_) for g_ in grads], p(ws), n, *dims, st), "b")
    for t in sets: fwd(t); bwd(t)
    torch.cuda.synchronize()
    _lib.profile_read(True); _lib.profile_enable(True)
    for i in range(iters):
        fwd(sets[i % nsets]); bwd(sets[i % nsets])
    _lib.profile_enable(False)
    prof = _lib.profile_read(True)
    W = B * Lout
    out = {k: v[0] / v[1] * 1e3 for k, v in prof.items()}
    bytes_ = {"qconv_fwd_kernel": W * 4 * (C * L / Lout + O), "qconv_bwd_post_kernel": W * 4 * O,
              "qconv_bwd_pre_kernel": W * 4 * C * L / Lout * (2 if need_gx else 1)}
    print(f"B={B} C={C} L={L} S={S} O={O} Lout={Lout} W={W}: " + "  ".join(
        f"{k.replace('qconv_','').replace('_kernel','')}={v:.1f}us" + (f"({bytes_[k]/v/1e3/6537:.2f})" if k in bytes_ else "") for k, v in out.items()))
for args in [(16, 80, 3000, 1, 384, False), (16, 384, 3000, 2, 384, True), (32, 384, 1500, 1, 384, True), (16, 384, 3072, 2, 384, True),
             (64, 384, 3000, 2, 384, True), (64, 80, 3000, 1, 384, False), (128, 384, 1500, 1, 384, True), (32, 384, 6000, 2, 384, True)]:
    run(*args)
