"""BASELINE.json config 4: n_qubits x n_layers sweep of the circuit alone (forward + adjoint backward), windows/s vs the
min(FMA, HBM) roofline.  Input pre = randn(W, q) (seed 3), weights randn(Lq, q, 3).  One JSON line per point.

    python tools/sweep_circuit.py [--embedding amplitude|angle] [--out profiles/r1_config4_sweep.jsonl]
"""
import argparse, ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qasr_ijcnlp_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--embedding", default="amplitude")
ap.add_argument("--out", default=None)
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
lib = _lib.load()
dev = torch.device("cuda:0")
emb = {"amplitude": 0, "angle": 1}[a.embedding]
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}
HBM = peaks["hbm_gbs"] * 1e9
FMA = 69.4e12  # fp32 FMA flop/s measured on B200 by tools/probe/ffma2_probe.cu (profiles/r1_fp32_issue_probe.txt; theoretical 74.4)
p = lambda t: ctypes.c_void_p(t.data_ptr())
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
rows = []
for q in (4, 6, 8, 10, 12):
    for Lq in (1, 2, 4):
        W = (1 << 20) if q <= 8 else (1 << 18) if q == 10 else (1 << 16)
        g = torch.Generator(device=dev).manual_seed(3)
        pre = torch.randn(W, q, device=dev, generator=g)
        w = torch.randn(Lq, q, 3, device=dev, generator=g)
        gout = torch.randn(W, q, device=dev, generator=g)
        out, gpre, gw = torch.empty_like(pre), torch.empty_like(pre), torch.empty_like(w)
        n = lib.qw_circuit_workspace_bytes(W, q, Lq, 4)
        ws = torch.empty(n, device=dev, dtype=torch.uint8)
        def fwd(): _lib.check(lib.qw_circuit_forward(p(pre), p(w), p(out), W, q, Lq, emb, st()), "f")
        def bwd(): _lib.check(lib.qw_circuit_backward(p(pre), p(w), p(gout), p(gpre), p(gw), p(ws), n, W, q, Lq, emb, st()), "b")
        fwd(); bwd(); torch.cuda.synchronize()
        _lib.profile_read(True); _lib.profile_enable(True)
        for _ in range(a.iters): fwd(); bwd()
        _lib.profile_enable(False)
        prof = _lib.profile_read(True)
        ms = {k: v[0] / v[1] for k, v in prof.items()}
        t_f, t_b = ms["circuit_fwd_kernel"], ms["circuit_bwd_kernel"] + ms.get("circuit_finalize_kernel", 0.0)
        N = 1 << q
        flop_f = 14.0 * q * N * Lq + (3 + q) * N
        # fwd + adjoint backward: recompute (1x) + un-apply psi (1x) + un-apply lambda (1x) + gate-gradient products
        # (8 of 14 flop per amplitude-gate, 0.6x) = 4.6x the forward flops; the adjoint method keeps no state per window
        flop_fb = flop_f * 4.6
        bytes_fb = 4.0 * q * (2 + 3)                       # fwd: pre in, out out; bwd: pre, gout in, gpre out
        ceil = min(FMA / flop_fb, HBM / bytes_fb)
        wps = W / ((t_f + t_b) * 1e-3)
        collapsed = None
        if a.embedding == "amplitude":
            # opt-in collapsed quadratic-form evaluation (SURVEY.md 8a iii), reported BESIDE the statevector numbers: same inputs,
            # M read off the statevector kernel on q(q+1)/2 probes; timed end to end through the public function with CUDA events
            from qasr_ijcnlp_b200 import quantum_circuit
            pre_r, w_r = pre.clone().requires_grad_(True), (w[0] if Lq == 1 else w).clone().requires_grad_(True)
            def both():
                o = quantum_circuit(pre_r, w_r, n_layers=Lq, simulator="collapsed")
                return o, torch.autograd.grad(o, [pre_r, w_r], gout)
            o_c, (gp_c, gw_c) = both()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters): both()
            e1.record(); torch.cuda.synchronize()
            t_c = e0.elapsed_time(e1) / a.iters
            collapsed = {"fwd_bwd_ms": round(t_c, 4), "windows_per_s_fwd_bwd": round(W / (t_c * 1e-3), 1),
                         "max_abs_diff_out_vs_statevector": float((o_c - out).abs().max()),
                         "max_rel_diff_gqw_vs_statevector": float((gw_c.reshape(-1) - gw.reshape(-1)).abs().max() / max(1.0, float(gw.abs().max())))}
        row = {"config": 4, "embedding": a.embedding, "n_qubits": q, "n_layers": Lq, "windows": W, "fwd_ms": round(t_f, 4),
               "bwd_ms": round(t_b, 4), "windows_per_s_fwd_bwd": round(wps, 1), "roofline_windows_per_s": round(ceil, 1),
               "bound": "fma" if FMA / flop_fb < HBM / bytes_fb else "hbm", "frac": round(wps / ceil, 4),
               "fwd_windows_per_s": round(W / (t_f * 1e-3), 1), "collapsed": collapsed}
        rows.append(row)
        print(json.dumps(row), flush=True)
if a.out:
    with open(a.out, "a") as fh:
        for r in rows: fh.write(json.dumps(r) + "\n")
