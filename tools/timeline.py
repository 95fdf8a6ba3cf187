"""In-graph kernel timeline of the stem step (debug tool): spans and gaps of the library's kernels INSIDE the replayed CUDA
graph of bench.py, read from %globaltimer stamps (qw_timeline_set).  Events cannot see inside a graph replay and ncu serialises
the kernels, so this is the only view of how the programmatic-dependent-launch chain really overlaps.

    python tools/timeline.py [--batch 16] [--reps 20] > profiles/rN_timeline_b16.txt
"""
import argparse
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--nsets", type=int, default=4)
    a = ap.parse_args()
    from qasr_ijcnlp_b200 import _lib

    rank, world, local = bench.dist_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(local)
    runner = bench.StemRunner(a.batch, dev, a.nsets)
    if world > 1:  # torchrun: the all-reduce fused into every layer's finalize kernel (qw_conv1d_backward_dp)
        from qasr_ijcnlp_b200 import dp

        torch.distributed.init_process_group("nccl", device_id=dev)
        runner.fused_dp = {name: dp.FusedLayerGradAllReduce(cfg["C"], cfg["O"], cfg["K"], cfg["S"], cfg["P"], bench.Q, 1, device=dev)
                           for name, cfg in bench.LAYERS.items()}
    lib = _lib.load()
    for i in range(2):
        n0 = _lib.launch_count()
        runner.step(i % a.nsets)
        per_step = _lib.launch_count() - n0
    torch.cuda.synchronize()
    nslots = per_step
    buf = torch.zeros(2 * nslots, dtype=torch.int64, device=dev)

    def reset():
        buf[0::2] = torch.iinfo(torch.int64).max
        buf[1::2] = 0

    _lib.check(lib.qw_timeline_set(ctypes.c_void_p(buf.data_ptr()), nslots), "qw_timeline_set")
    side = torch.cuda.Stream()
    graphs = []
    for s in range(a.nsets):
        _lib.check(lib.qw_timeline_set(ctypes.c_void_p(buf.data_ptr()), nslots), "qw_timeline_set")  # slot counter back to 0
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            runner.step(s)
        graphs.append(g)
    torch.cuda.synchronize()
    for i in range(8):
        graphs[i % a.nsets].replay()
    torch.cuda.synchronize()
    names = [n.replace("bwd_post(gy)", "gy").replace("bwd_adj", "adj").replace("bwd_pre", "pre").replace("bwd_finalize", "fin")
             .replace("bwd_fused", "fused") for n in bench.step_kernel_names(runner)]
    rows = []
    for r in range(a.reps):
        reset()
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        graphs[r % a.nsets].replay()
        torch.cuda.synchronize()
        t = buf.cpu().tolist()
        rows.append([(t[2 * k], t[2 * k + 1]) for k in range(nslots)])
    lib.qw_timeline_set(None, 0)
    # median over reps of span / gap, relative to the first kernel's start
    import statistics as st

    if rank != 0:
        torch.distributed.destroy_process_group()
        return
    print(f"# stem step, batch {a.batch}, world {world}, {per_step} kernels per step, median of {a.reps} single graph replays (ns from %globaltimer)")
    print(f"# {'kernel':12s} {'start':>9s} {'end':>9s} {'span':>8s} {'gap_to_prev_end':>16s}")
    tot = []
    for k in range(nslots):
        s0 = st.median(r[k][0] - r[0][0] for r in rows)
        e0 = st.median(r[k][1] - r[0][0] for r in rows)
        gap = st.median((r[k][0] - r[k - 1][1]) for r in rows) if k else 0
        nm = names[k] if k < len(names) else f"k{k}"
        print(f"  {nm:12s} {s0 / 1e3:9.2f} {e0 / 1e3:9.2f} {(e0 - s0) / 1e3:8.2f} {gap / 1e3:16.2f}")
        tot.append(e0)
    print(f"# first start -> last end: {max(tot) / 1e3:.2f} us; timer granularity seen: "
          f"{min(abs(x - y) for r in rows for (x, _), (y, _) in zip(r, r[1:]) if x != y)} ns (smallest non-zero start delta)")


if __name__ == "__main__":
    main()
