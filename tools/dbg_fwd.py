"""Forward-kernel times at the bench workload with parts of the work switched off (option DBG_FWD, results garbage): which part of
a kernel's duration is arithmetic (1 = pre_conv FMAs, 2 = post_conv / conv2 pre_conv FMAs, 4 = circuits) and which is its output
stores (8; fused training forward only).  200 back-to-back launches per number, CUDA events.

    python tools/dbg_fwd.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from qasr_ijcnlp_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
r = bench.StemRunner(16, dev, 4)
def t(fn, n=200):
    for i in range(8): fn(i % 4)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i % 4)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
try:
    for dbg in (0, 1, 2, 4, 7, 8, 15):
        lib.qw_set_option(b"DBG_FWD", dbg)
        print("DBG_FWD %2d  fused stem forward %.1f us   conv1 forward %.1f us   conv2 forward %.1f us" %
              (dbg, t(r.fwd_stem), t(lambda s: r.fwd("conv1", s)), t(lambda s: r.fwd("conv2", s))), flush=True)
finally:
    lib.qw_set_option(b"DBG_FWD", 0)
