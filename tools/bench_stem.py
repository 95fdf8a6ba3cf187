"""Inference stem: fused (qw_stem_forward) vs operator-by-operator (conv1 -> gelu -> conv2 -> gelu -> permute + pos), utt/s.

    python tools/bench_stem.py [--batches 1,4,16,64,256] >> profiles/rN_stem_fused.jsonl
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="1,4,16,64,256")
    ap.add_argument("--reps", type=int, default=50)
    a = ap.parse_args()
    from qasr_ijcnlp_b200 import QuantumConv1d, _lib, fused_stem_forward
    from qasr_ijcnlp_b200.encoder import sinusoids

    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    c1 = QuantumConv1d(80, 384, kernel_size=3, padding=1, n_qubits=4).to(dev)
    c2 = QuantumConv1d(384, 384, kernel_size=3, stride=2, padding=1, n_qubits=4).to(dev)
    pos = sinusoids(1500, 384).to(dev)
    for B in [int(v) for v in a.batches.split(",")]:
        nsets = max(2, min(8, (256 << 20) // (B * 80 * 3000 * 4) + 1))  # rotate inputs; outputs alone exceed L2 from B = 64
        xs = [torch.rand(B, 80, 3000, device=dev) * 3 - 1.5 for _ in range(nsets)]
        it = [0]

        def fused():
            it[0] += 1
            return fused_stem_forward(c1, c2, xs[it[0] % nsets], pos)

        def unfused():
            it[0] += 1
            with torch.no_grad():
                return F.gelu(c2(F.gelu(c1(xs[it[0] % nsets])))).permute(0, 2, 1) + pos

        with torch.no_grad():
            err = (fused_stem_forward(c1, c2, xs[0], pos) - (F.gelu(c2(F.gelu(c1(xs[0])))).permute(0, 2, 1) + pos)).abs().max().item()
        t_f, t_u = timed(fused, a.reps), timed(unfused, a.reps)
        # graph-replayed fused path (no Python / launch overhead)
        side = torch.cuda.Stream()
        g = torch.cuda.CUDAGraph()
        fused()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=side):
            out = fused_stem_forward(c1, c2, xs[0], pos)
        t_g = timed(g.replay, a.reps)
        _lib.profile_read(reset=True)
        _lib.profile_enable(True)
        for _ in range(10):
            fused()
        torch.cuda.synchronize()
        _lib.profile_enable(False)
        prof = {k: round(v[0] / v[1], 5) for k, v in _lib.profile_read(reset=True).items()}
        algo_bytes = B * 4 * (80 * 3000 + 1500 * 384)  # mel in + (B, 1500, 384) out
        print(json.dumps({"workload": "inference stem: mel (B,80,3000) -> (B,1500,384), n_qubits=4", "batch": B,
                          "fused_ms": round(t_f, 5), "fused_graph_ms": round(t_g, 5), "unfused_ms": round(t_u, 5),
                          "speedup": round(t_u / t_f, 2), "fused_utt_per_s": round(B / t_g * 1e3, 1),
                          "unfused_utt_per_s": round(B / t_u * 1e3, 1), "max_abs_diff": err, "kernel_ms": prof,
                          "fused_GBps_algorithmic": round(algo_bytes / t_g / 1e6, 1)}), flush=True)


if __name__ == "__main__":
    main()
