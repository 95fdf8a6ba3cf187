#!/usr/bin/env python
"""bench.py -- QuantumConv1d windows/s forward+backward on the Whisper-Tiny quantum stem (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--batch 16]

One "step" = one pass of the hot path over one batch of synthetic input: conv1 (80->384, k3 s1 p1, 3000 frames)
forward + backward and conv2 (384->384, k3 s2 p1) forward + backward (with grad_x), n_qubits = 4, at
`--batch` utterances per GPU (BASELINE.json configs[1]/[2]: batch 16, 80 mel x 3000 frames) = 4500 windows per
utterance per step.

  value      windows/s with every input already resident in HBM; the step is one CUDA graph of the C-ABI calls,
             replayed over `nsets` rotating buffer sets so no kernel finds its input in L2.
  roofline   the dominant kernel (largest share of the step), timed with CUDA events inside the library
             (qw_profile_*), algorithmic bytes / duration against the measured HBM peak.
  e2e        the same step through the public nn.Module API with HOST input: pinned mel batch -> H2D -> conv1 ->
             GELU -> conv2 -> loss -> backward -> D2H of loss and the quantum-layer gradients, all inside the
             timed region.
  cpu_baseline  the literal-loop fp64 restatement of the reference (oracle/, PennyLane is not installable) on a
             bounded sample, on this box's host cores.
`--impl reference` times that same literal-loop restatement as the reference arm.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

_REAL_STDOUT = None


def emit(line: dict) -> None:
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


METRIC = "qconv_windows_per_sec_fwd_bwd"
UNIT = "windows/s"
N_MELS, N_STATE, N_FRAMES, Q = 80, 384, 3000, 4
LAYERS = {
    "conv1": dict(C=N_MELS, O=N_STATE, K=3, S=1, P=1, L=N_FRAMES, Lout=3000, need_gx=False),
    "conv2": dict(C=N_STATE, O=N_STATE, K=3, S=2, P=1, L=N_FRAMES, Lout=1500, need_gx=True),
}
WINDOWS_PER_UTT = 4500


# --------------------------------------------------------------------------------------------- helpers
def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="utterances per GPU")
    ap.add_argument("--nsets", type=int, default=4, help="rotating buffer sets (L2 defeat)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-encoder", action="store_true")
    ap.add_argument("--e2e-eager", action="store_true", help="do not wrap the e2e module in torch.cuda.make_graphed_callables")
    ap.add_argument("--stem-forward", default="auto", choices=["auto", "fused", "split"],
                    help="fused: conv1 + conv2 forward in one kernel (qw_stem_train_forward); split: one qw_conv1d_forward per layer; "
                         "auto: whichever the library prefers for this batch (qw_stem_train_forward_preferred)")
    ap.add_argument("--stem-backward", default="chained", choices=["chained", "split"],
                    help="chained: conv2's backward writes no grad_x, conv1's gy kernel rebuilds it from conv2's gpre rows "
                         "(qw_conv1d_backward_chained); split: the gradient goes through HBM between the two layers' backwards")
    ap.add_argument("--collective", default="hybrid", choices=["hybrid", "fused", "p2p", "nccl"],
                    help="gradient all-reduce at N > 1: fused into the backward's last kernel / own one-shot NVLink kernel / NCCL")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", d
    return 6650.0, "fallback (B200_PROFILING.md)", {}


def load_traffic():
    """Per-launch DRAM traffic of the dominant kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as fh:
            return json.load(fh)
    return {}


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        try:
            for line in open(self.path):
                f = [t.strip() for t in line.split(",")]
                if len(f) < 8:
                    continue
                try:
                    sm.append(float(f[0]))
                    mx.append(float(f[1]))
                    pw.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            busy = [c for c, p in zip(sm, pw) if p >= 0.5 * max(pw)] or sm
            out.update(sm_mhz=statistics.median(busy), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(pw))
        return out


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


# --------------------------------------------------------------------------------------------- B200 arm
class StemBuffers:
    """HBM-resident inputs/outputs of one step, for one rotating set."""

    def __init__(self, B, dev, seed):
        g = torch.Generator(device=dev).manual_seed(seed)
        self.t = {}
        for name, cfg in LAYERS.items():
            C, O, L, Lout = cfg["C"], cfg["O"], cfg["L"], cfg["Lout"]
            self.t[name] = dict(
                x=torch.randn(B, C, L, device=dev, generator=g),
                y=torch.empty(B, O, Lout, device=dev),
                gy=torch.randn(B, O, Lout, device=dev, generator=g),
                pre=torch.empty(2, B * Lout, Q, device=dev),
                gx=torch.empty(B, C, L, device=dev) if cfg["need_gx"] else None,
            )


class StemParams:
    def __init__(self, dev, seed=0):
        from qasr_ijcnlp_b200 import QuantumConv1d

        torch.manual_seed(seed)
        self.mods = {
            "conv1": QuantumConv1d(N_MELS, N_STATE, kernel_size=3, padding=1, n_qubits=Q).to(dev),
            "conv2": QuantumConv1d(N_STATE, N_STATE, kernel_size=3, stride=2, padding=1, n_qubits=Q).to(dev),
        }
        self.p, self.g = {}, {}
        for name, m in self.mods.items():
            ps = [m.pre_conv.weight, m.pre_conv.bias, m.quantum_weights, m.post_conv.weight, m.post_conv.bias]
            self.p[name] = [t.detach().contiguous() for t in ps]
        # one flat gradient bucket (9 440 floats, padded per tensor to 16 B): the backward kernels write straight into it and the
        # data-parallel all-reduce runs on it in place
        sizes = [(name, i, (t.numel() + 3) // 4 * 4) for name in self.mods for i, t in enumerate(self.p[name])]
        self.flat_grads = torch.zeros(sum(n for _, _, n in sizes), device=dev)
        off = 0
        self.span = {}  # layer -> (start, end) inside the flat bucket
        for name, i, n in sizes:
            t = self.p[name][i]
            self.g.setdefault(name, []).append(self.flat_grads[off:off + t.numel()].view_as(t))
            lo, _ = self.span.get(name, (off, off))
            self.span[name] = (lo, off + n)
            off += n


class StemRunner:
    def __init__(self, B, dev, nsets):
        from qasr_ijcnlp_b200 import _lib

        self._lib = _lib
        self.lib = _lib.load()
        self.B, self.dev = B, dev
        self.params = StemParams(dev)
        self.sets = [StemBuffers(B, dev, 100 + s) for s in range(nsets)]
        self.ws = {}
        for name, cfg in LAYERS.items():
            n = self.lib.qw_conv1d_workspace_bytes(B, cfg["C"], cfg["L"], cfg["K"], cfg["S"], cfg["P"], cfg["O"], Q, 1, 4)
            self.ws[name] = (torch.empty(n, device=dev, dtype=torch.uint8), n)

    def _dims(self, cfg):
        return (self.B, cfg["C"], cfg["L"], cfg["K"], cfg["S"], cfg["P"], cfg["O"], Q, 1, 0)

    def fwd(self, name, s):
        cfg, t, p = LAYERS[name], self.sets[s].t[name], self.params.p[name]
        st = self.lib.qw_conv1d_forward(_p(t["x"]), *[_p(w) for w in p], _p(t["y"]), _p(t["pre"]), *self._dims(cfg),
                                        ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        self._lib.check(st, "qw_conv1d_forward")

    fused_fwd = True                 # conv1 + conv2 forward as ONE kernel (qw_stem_train_forward): conv1's output lands in conv2's x buffer

    def fwd_stem(self, s):
        """Both forwards in one kernel: y1 = conv1(x) is written straight into the buffer conv2's backward reads as its x (in the
        split form conv1 writes its own y buffer and conv2 reads an equally large x buffer: same bytes written, one read fewer)."""
        t1, t2 = self.sets[s].t["conv1"], self.sets[s].t["conv2"]
        p1, p2 = self.params.p["conv1"], self.params.p["conv2"]
        c1, c2 = LAYERS["conv1"], LAYERS["conv2"]
        st = self.lib.qw_stem_train_forward(_p(t1["x"]), *[_p(w) for w in p1], *[_p(w) for w in p2], _p(t2["x"]), _p(t1["pre"]),
                                            _p(t2["y"]), _p(t2["pre"]), self.B, c1["C"], c1["L"], c1["O"], c2["O"], 1, 0,
                                            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        self._lib.check(st, "qw_stem_train_forward")

    hybrid = None                    # (P2P reducer of conv2's bucket, side stream) with fused_dp = {"conv1": ...}
    fused_dp = None                  # {layer: dp.FusedLayerGradAllReduce}: the all-reduce rides in the backward's finalize kernel

    chained = True                   # conv1's backward rebuilds its incoming gradient from conv2's gpre rows (qw_conv1d_backward_chained):
                                     # conv2's backward writes no grad_x and conv1's gy kernel reads none

    def bwd(self, name, s):
        cfg, t, p, g = LAYERS[name], self.sets[s].t[name], self.params.p[name], self.params.g[name]
        ws, n = self.ws[name]
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        if self.chained and name == "conv2":
            # no gradient for conv2's input is written: conv1's chained backward does not need it
            st = self.lib.qw_conv1d_backward(_p(t["gy"]), _p(t["x"]), _p(t["pre"]), _p(p[0]), _p(p[2]), _p(p[3]), None,
                                             _p(g[0]), _p(g[1]), _p(g[2]), _p(g[3]), _p(g[4]), _p(ws), n, *self._dims(cfg), stream)
            self._lib.check(st, "qw_conv1d_backward")
            return
        if self.chained and name == "conv1":
            ws2, _ = self.ws["conv2"]
            p2 = self.params.p["conv2"]
            dp = self.fused_dp[name].args() if (self.fused_dp is not None and name in self.fused_dp) else (None, None, 0, 1, 1.0)
            st = self.lib.qw_conv1d_backward_chained(_p(ws2), _p(p2[0]), LAYERS["conv2"]["O"], _p(t["x"]), _p(t["pre"]), _p(p[0]), _p(p[2]),
                                                     _p(p[3]), _p(p[4]), _p(t["gx"]), _p(g[0]), _p(g[1]), _p(g[2]), _p(g[3]), _p(g[4]),
                                                     _p(ws), n, *self._dims(cfg), 0, *dp, stream)
            self._lib.check(st, "qw_conv1d_backward_chained")
            return
        if self.fused_dp is not None and name in self.fused_dp:
            st = self.lib.qw_conv1d_backward_dp(_p(t["gy"]), _p(t["x"]), _p(t["pre"]), _p(p[0]), _p(p[2]), _p(p[3]), _p(t["gx"]),
                                                _p(g[0]), _p(g[1]), _p(g[2]), _p(g[3]), _p(g[4]), _p(ws), n, *self._dims(cfg),
                                                *self.fused_dp[name].args(),
                                                ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
            self._lib.check(st, "qw_conv1d_backward_dp")
            return
        st = self.lib.qw_conv1d_backward(_p(t["gy"]), _p(t["x"]), _p(t["pre"]), _p(p[0]), _p(p[2]), _p(p[3]), _p(t["gx"]),
                                         _p(g[0]), _p(g[1]), _p(g[2]), _p(g[3]), _p(g[4]), _p(ws), n, *self._dims(cfg),
                                         ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        self._lib.check(st, "qw_conv1d_backward")

    def step(self, s):
        # training order of the stem: conv1 fwd, conv2 fwd, [rest of the model], conv2 bwd, conv1 bwd, gradient all-reduce
        if self.fused_fwd:
            self.fwd_stem(s)
        else:
            self.fwd("conv1", s)
            self.fwd("conv2", s)
        self.bwd("conv2", s)
        if self.hybrid is not None:
            # conv2's gradients: own one-shot NVLink all-reduce kernel on a side stream, hidden under conv1's backward;
            # conv1's gradients (the last to be produced): all-reduce fused into conv1's finalize kernel -> nothing after it
            ar2, side = self.hybrid
            main = torch.cuda.current_stream()
            ev = torch.cuda.Event()
            ev.record(main)
            side.wait_event(ev)
            lo, hi = self.params.span["conv2"]
            with torch.cuda.stream(side):
                ar2(self.params.flat_grads[lo:hi])
                ev2 = torch.cuda.Event()
                ev2.record(side)
            self.bwd("conv1", s)
            main.wait_event(ev2)
            return
        if self.allreduce_split is not None:
            # conv2's gradients are final: reduce them on a side stream while conv1's backward runs; only conv1's (smaller)
            # bucket is reduced on the critical path
            ar1, ar2, side = self.allreduce_split
            main = torch.cuda.current_stream()
            ev = torch.cuda.Event()
            ev.record(main)
            side.wait_event(ev)
            lo, hi = self.params.span["conv2"]
            with torch.cuda.stream(side):
                ar2(self.params.flat_grads[lo:hi])
                ev2 = torch.cuda.Event()
                ev2.record(side)
            self.bwd("conv1", s)
            lo, hi = self.params.span["conv1"]
            ar1(self.params.flat_grads[lo:hi])
            main.wait_event(ev2)
            return
        self.bwd("conv1", s)
        if self.allreduce is not None:
            self.allreduce(self.params.flat_grads)

    allreduce = None                 # set by run_b200 when world > 1 (single collective at the end of the step)
    allreduce_split = None           # or (reducer conv1, reducer conv2, side stream): per-layer buckets, conv2's overlapped


CHAINED = False  # set by run_b200: the step uses qw_conv1d_backward_chained


def algorithmic_bytes(kernel, layer, B):
    """Algorithmic HBM bytes of ONE launch (SURVEY.md 8d per-window figures x windows per launch)."""
    if layer == "stem":  # fused forward of both layers: the SURVEY figure is per layer, so the sum of the two forwards
        return algorithmic_bytes(kernel, "conv1", B) + algorithmic_bytes(kernel, "conv2", B) if kernel == "qconv_fwd_kernel" else 0.0
    cfg = LAYERS[layer]
    W = B * cfg["Lout"]
    x_per_win = 4.0 * cfg["C"] * cfg["L"] / cfg["Lout"]
    y_per_win = 4.0 * cfg["O"]
    if kernel == "qconv_fwd_kernel":
        return W * (x_per_win + y_per_win)
    if kernel == "qconv_bwd_post_kernel":
        return W * y_per_win  # reads gy once
    if kernel == "qconv_bwd_pre_kernel":
        # re-read x (+ write grad_x -- not in the chained backward, where conv2's gradient for its input is never written)
        return W * x_per_win * (2.0 if (cfg["need_gx"] and not (CHAINED and layer == "conv2")) else 1.0)
    if kernel == "qconv_bwd_fused_kernel":
        return W * (y_per_win + x_per_win * (2.0 if cfg["need_gx"] else 1.0))  # the whole backward of the layer
    return 0.0


def stem_fwd_moved_bytes(B):
    """Bytes the fused forward kernel really moves: x in, y1 out, y2 out, both pre_save buffers (conv2's read of y1 is gone)."""
    c1, c2 = LAYERS["conv1"], LAYERS["conv2"]
    return 4.0 * B * (c1["C"] * c1["L"] + c1["O"] * c1["Lout"] + c2["O"] * c2["Lout"] + 8 * (c1["Lout"] + c2["Lout"]))


def make_config(B, world, nsets):
    """The workload both arms are measured on (the reference arm times a bounded sample of it: `cpu_baseline.sample`)."""
    return {
        "workload": f"Whisper-Tiny quantum stem: QuantumConv1d conv1(80->384,k3,s1,p1)+conv2(384->384,k3,s2,p1) "
                    f"fwd+bwd, n_qubits=4, batch {B}/GPU x 80 mel x 3000 frames (BASELINE configs[1]/[2] shape)",
        "windows_per_step_per_gpu": B * WINDOWS_PER_UTT, "batch_per_gpu": B, "n_qubits": Q, "n_layers": 1,
        "l2": f"GPU arm: {nsets} rotating buffer sets ({nsets} x {runner_bytes(B) / 1e6:.0f} MB > 126 MB L2), CUDA graphs of {nsets} consecutive "
              f"steps (one per buffer set); "
              f"CPU reference arm: not applicable",
        "parallelism": f"dp{world} (batch shard over {world} GPU(s); forward: no collective; training: one mean of the gradients)",
    }


def check_dp_parity(runner, world, dev):
    """One step with the configured gradient collective (NVLink peer-memory kernels) and one with the plain backward followed by
    torch.distributed.all_reduce(AVG) (NCCL) on the same inputs and parameters: max error relative to the largest reference
    entry, bitwise equality of the result across ranks, and the collectives' status words.  (The reference trains single
    process, train_quantum_whisper.py:195-214: this is the correctness proof of what replaces DistributedDataParallel.)"""
    import torch.distributed as dist

    runner.step(0)
    torch.cuda.synchronize()
    got = runner.params.flat_grads.clone()
    saved = (runner.fused_dp, runner.hybrid, runner.allreduce, runner.allreduce_split)
    runner.fused_dp = runner.hybrid = runner.allreduce = runner.allreduce_split = None
    try:
        runner.step(0)
        torch.cuda.synchronize()
        ref = runner.params.flat_grads.clone()
    finally:
        runner.fused_dp, runner.hybrid, runner.allreduce, runner.allreduce_split = saved
    local_max = ref.abs().max().item()
    dist.all_reduce(ref, op=dist.ReduceOp.AVG)
    err = (got - ref).abs().max().item() / max(1.0, ref.abs().max().item())
    gathered = [torch.empty_like(got) for _ in range(world)]
    dist.all_gather(gathered, got)
    bitwise = all(torch.equal(gathered[0], t) for t in gathered)
    status = 0
    for obj in list((runner.fused_dp or {}).values()) + ([runner.hybrid[0]] if runner.hybrid else []) + \
            (list(runner.allreduce_split[:2]) if runner.allreduce_split else []):
        status |= int(obj.status())
    t = torch.tensor([err, 0.0 if bitwise else 1.0, float(status)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {"max_rel_err": float(t[0].item()), "bitwise_equal_across_ranks": bool(t[1].item() == 0.0), "status_word": int(t[2].item()),
            "reference": "the same step with every gradient collective removed (qw_conv1d_backward[_chained], world = 1 arguments) + "
                         "torch.distributed.all_reduce(AVG) (NCCL) on the same inputs",
            "tolerance": 1e-6, "grad_floats": int(got.numel()), "local_grad_absmax": local_max}


def time_events(fn, steps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for i in range(steps):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b)


def barrier(world):
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()


def max_over_ranks(ms, world, dev):
    if world == 1:
        return ms
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


def run_b200(args):
    rank, world, local = dist_env()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pin_to_gpu_numa_node(local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    from qasr_ijcnlp_b200 import _lib

    B, K, Wm, nsets = args.batch, args.steps, max(args.warmup, 3), args.nsets
    runner = StemRunner(B, dev, nsets)
    runner.fused_fwd = (args.stem_forward == "fused" or
                        (args.stem_forward == "auto" and bool(runner.lib.qw_stem_train_forward_preferred(B, LAYERS["conv1"]["L"]))))
    runner.chained = args.stem_backward == "chained"
    global CHAINED
    CHAINED = runner.chained
    windows_per_step = B * WINDOWS_PER_UTT
    collective = "none (single GPU)"
    if world > 1:
        from qasr_ijcnlp_b200 import dp
        n = runner.params.flat_grads.numel()
        try:
            if args.collective not in ("fused", "hybrid"):
                raise RuntimeError("not requested")
            layers = LAYERS if args.collective == "fused" else {"conv1": LAYERS["conv1"]}
            runner.fused_dp = {name: dp.FusedLayerGradAllReduce(cfg["C"], cfg["O"], cfg["K"], cfg["S"], cfg["P"], Q, 1, device=dev)
                               for name, cfg in layers.items()}
            collective = ("fused into each layer's backward: the finalize kernel exchanges its reduced columns over NVLink peer "
                          "memory and averages them (qw_conv1d_backward_dp); no separate collective kernel")
            if args.collective == "hybrid":
                lo, hi = runner.params.span["conv2"]
                runner.hybrid = (dp.P2PGradAllReduce(hi - lo, dev), torch.cuda.Stream())
                collective = (f"conv1 (last gradients of the step): all-reduce fused into its backward's finalize kernel over NVLink "
                              f"peer memory (qw_conv1d_backward_dp); conv2: own one-shot NVLink all-reduce kernel of its {4 * (hi - lo)} B "
                              f"bucket on a side stream under conv1's backward; all inside the step's CUDA graph")
        except Exception as e_f:
            runner.fused_dp = None
            runner.hybrid = None
            fused_note = "" if args.collective != "fused" else f" [fused unavailable: {type(e_f).__name__}: {str(e_f)[:60]}]"
        try:
            if runner.fused_dp is not None:
                raise StopIteration
            if args.collective == "nccl":
                raise RuntimeError("forced")
            sp = runner.params.span
            n1, n2 = sp["conv1"][1] - sp["conv1"][0], sp["conv2"][1] - sp["conv2"][0]
            runner.allreduce_split = (dp.P2PGradAllReduce(n1, dev), dp.P2PGradAllReduce(n2, dev), torch.cuda.Stream())
            collective = (f"own one-shot NVLink peer-memory all-reduce kernel, fused 1/N scale, in the step's CUDA graph: conv2's "
                          f"{4 * n2} B bucket on a side stream under conv1's backward, conv1's {4 * n1} B bucket at the end")
        except StopIteration:
            pass
        except Exception as e:  # symmetric memory unavailable -> NCCL
            flat = runner.params.flat_grads
            runner.allreduce = lambda t: torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.AVG)
            collective = f"NCCL all_reduce(AVG) of the {4 * n} B bucket, in the step's CUDA graph ({type(e).__name__}: {str(e)[:80]})"

    # ---- data-parallel parity (driver-visible): the configured collective against plain backward + NCCL all_reduce(AVG)
    dp_parity = None
    if world > 1:
        dp_parity = check_dp_parity(runner, world, dev)
        if dp_parity["max_rel_err"] > 1e-6 or not dp_parity["bitwise_equal_across_ranks"] or dp_parity["status_word"] != 0:
            if rank == 0:
                emit({"metric": METRIC, "error": "data-parallel gradient parity failed", "dp_parity": dp_parity, "n_gpus": world})
            torch.distributed.destroy_process_group()
            sys.exit(3)

    # ---- warm up eagerly (also sets function attributes), then capture one CUDA graph per buffer set
    for i in range(2):
        n_before = _lib.launch_count()
        runner.step(i % nsets)
        launches_per_step = _lib.launch_count() - n_before  # counted, not assumed: every kernel of the step is the library's own
    torch.cuda.synchronize()
    side = torch.cuda.Stream()

    def capture():
        gs = []
        for s_ in range(nsets):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                runner.step(s_)
            gs.append(g)
        # `nsets` consecutive steps (one per buffer set) as ONE graph: the boundary between two steps becomes an ordinary
        # kernel-to-kernel edge of the programmatic-dependent-launch chain instead of a graph-launch gap (~3.5 us on B200)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for s_ in range(nsets):
                runner.step(s_)
        gs.append(g)
        return gs

    post = None
    try:
        graphs = capture()
    except Exception as e:  # a collective that cannot be captured: keep it eager, right behind the graph
        if runner.allreduce is None and runner.allreduce_split is None:
            raise
        if runner.allreduce_split is not None:
            ar1, ar2, _ = runner.allreduce_split
            sp = runner.params.span
            runner.allreduce_split = None
            runner.allreduce = lambda t: (ar2(t[sp["conv2"][0]:sp["conv2"][1]]), ar1(t[sp["conv1"][0]:sp["conv1"][1]]))
        post, runner.allreduce = runner.allreduce, None
        collective += f" [not capturable: {type(e).__name__}; issued eagerly after the graph]"
        torch.cuda.synchronize()
        graphs = capture()
    torch.cuda.synchronize()

    def replay(i):
        graphs[i % nsets].replay()
        if post is not None:
            post(runner.params.flat_grads)

    def replay_steps(n):
        """exactly n steps: blocks of `nsets` steps through the multi-step graph, the remainder step by step"""
        if post is not None:
            for i in range(n):
                replay(i)
            return
        q, r = divmod(n, nsets)
        for _ in range(q):
            graphs[nsets].replay()
        for i in range(r):
            graphs[i].replay()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    replay_steps(max(Wm, nsets))
    barrier(world)
    ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev_a.record()
    replay_steps(K)
    ev_b.record()
    torch.cuda.synchronize()
    ms = ev_a.elapsed_time(ev_b)
    barrier(world)
    ms = max_over_ranks(ms, world, dev)
    value = world * windows_per_step * K / (ms * 1e-3)
    own_per_step = launches_per_step
    gpu_launches = K * own_per_step

    # ---- per-kernel durations (CUDA events inside the library, eager launches, same rotating buffers)
    kern = {}
    calls = {}
    todo = [(layer, what, fn) for layer in ("conv1", "conv2") for what, fn in (("fwd", runner.fwd), ("bwd", runner.bwd))]
    if runner.fused_fwd:  # the step's forward is ONE kernel for both layers
        todo = [("stem", "fwd", lambda _l, s_: runner.fwd_stem(s_))] + [t for t in todo if t[1] == "bwd"]
    for layer, what, fn in todo:
        if True:
            for i in range(3):
                fn(layer, i % nsets)
            torch.cuda.synchronize()
            _lib.profile_read(reset=True)
            _lib.profile_enable(True)
            call_ms = time_events(lambda i: fn(layer, i % nsets), K)
            _lib.profile_enable(False)
            prof = _lib.profile_read(reset=True)
            calls[f"{layer}.{what}"] = call_ms / K
            for kname, (tot, n) in prof.items():
                kern[(layer, kname)] = tot / n
    peak, peak_src, peaks_raw = load_peaks()
    step_kernel_ms = sum(kern.values())
    dom = max(kern, key=lambda k: kern[k])
    dom_bytes = algorithmic_bytes(dom[1], dom[0], B)
    # The dominant kernel's AVERAGE LAUNCH DURATION, CUDA events on the launching stream: `reps` launches of that C-ABI call back
    # to back over the rotating buffer sets between two events (the way the kernel runs inside the step: queued behind its
    # predecessor, programmatic dependent launch on), divided by `reps`.  The single-launch bracket (an event pair around ONE
    # launch, qw_profile_*) is kept beside it: it adds ~3-5 us of launch / event latency that no step ever pays.
    dom_ms, dom_how = kern[dom], "CUDA events around one launch (qw_profile_*)"
    if dom[1] == "qconv_fwd_kernel":
        reps = max(40, K)
        dom_fn = (lambda s_: runner.fwd_stem(s_)) if dom[0] == "stem" else (lambda s_: runner.fwd(dom[0], s_))
        for i in range(8):
            dom_fn(i % nsets)
        dom_ms = time_events(lambda i: dom_fn(i % nsets), reps) / reps
        call_name = "qw_stem_train_forward" if dom[0] == "stem" else f"qw_conv1d_forward[{dom[0]}]"
        dom_how = (f"CUDA events around {reps} back-to-back launches of {call_name} over {nsets} rotating buffer sets "
                   f"/ {reps} (average launch duration in stream order)")
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    symbols = {}
    for layer in ("conv1", "conv2"):  # the symbols an ncu launch list shows, per layer (template arguments differ)
        runner.fwd(layer, 0)
        runner.bwd(layer, 0)
        for kname, sym in _lib.kernel_symbols().items():
            symbols[(layer, kname)] = sym
    if runner.fused_fwd:
        runner.fwd_stem(0)
        symbols[("stem", "qconv_fwd_kernel")] = _lib.kernel_symbols().get("qconv_fwd_kernel", "stem_train_fwd_kernel")
    torch.cuda.synchronize()
    traffic = load_traffic().get(f"{dom[0]}.{dom[1]}")
    roofline = {
        "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
        "traffic": traffic, "kernel": f"qw::{symbols.get(dom, dom[1])}", "layer": dom[0], "kernel_ms": round(dom_ms, 5), "how": dom_how,
        "kernel_ms_single_launch_bracket": round(kern[dom], 5),
        "frac_single_launch_bracket": round(dom_bytes / (kern[dom] * 1e-3) / 1e9 / peak, 4),
        "kernel_share_of_step": round(kern[dom] / step_kernel_ms, 4), "algorithmic_bytes_per_launch": dom_bytes,
        "peak_source": peak_src,
    }
    if dom[0] == "stem":
        # `achieved` follows the contract (SURVEY 8d per-layer bytes x windows): the fused kernel does the work of both forwards.
        # It MOVES less -- conv2's read of the (B, hidden, L) activation is gone -- and that number is stated beside it.
        moved = stem_fwd_moved_bytes(B)
        roofline.update({"moved_bytes_per_launch": moved, "moved_GBps": round(moved / (dom_ms * 1e-3) / 1e9, 1),
                         "frac_of_moved_bytes": round(moved / (dom_ms * 1e-3) / 1e9 / peak, 4),
                         "note": "one kernel for conv1 + conv2 forward (qw_stem_train_forward): algorithmic bytes = the two layers' "
                                 "SURVEY figures (the work it replaces); moved bytes = what this formulation needs"})
    kernels = {}
    for (layer, kname), t in sorted(kern.items()):
        ab = algorithmic_bytes(kname, layer, B)
        kernels[f"{layer}.{kname}"] = {"symbol": "qw::" + symbols.get((layer, kname), kname), "ms": round(t, 5),
                                       "share": round(t / step_kernel_ms, 4),
                                       "GBps": round(ab / (t * 1e-3) / 1e9, 1) if ab else None,
                                       "frac": round(ab / (t * 1e-3) / 1e9 / peak, 4) if ab else None}
    # whole-step roofline: algorithmic bytes of the step (3712 + 12288/2 ... per window) / step time
    step_bytes = B * (3000 * 3712 + 1500 * 12288)
    step_frac = step_bytes / (ms / K * 1e-3) / 1e9 / peak

    # ---- where the step goes INSIDE the replayed graph (CUDA events cannot see there): %globaltimer stamps per kernel
    in_graph = None
    try:
        in_graph = in_graph_timeline(runner, nsets, dev, B, peak, world=world)
        if world == 1:
            # the same step with the arithmetic of the three streaming kernels removed (A/B switches DBG_FWD / DBG_GY: TMA
            # pipelines, barriers, stores and launch structure unchanged, results garbage): what the memory system and the
            # kernel boundaries alone cost at this batch -- the floor `step_us` can be compared with
            lib = _lib.load()
            try:
                lib.qw_set_option(b"DBG_FWD", 7)
                lib.qw_set_option(b"DBG_GY", 1)
                floor = in_graph_timeline(runner, nsets, dev, B, peak, world=world)
            finally:
                lib.qw_set_option(b"DBG_FWD", 0)
                lib.qw_set_option(b"DBG_GY", 0)
            in_graph["zero_compute_floor"] = {
                "step_us": floor["step_us"], "kernels_us": {k: v["us"] for k, v in floor["kernels"].items()},
                "step_roofline_frac": round(step_bytes / (floor["step_us"] * 1e-6) / 1e9 / peak, 4),
                "how": "same graph, forward / gy kernels with their FMAs / MMAs / circuit removed (qw_set_option DBG_FWD=7, DBG_GY=1)"}
    except Exception as e:
        in_graph = {"error": repr(e)[:200]}

    # ---- e2e through the nn.Module API with host buffers
    e2e = run_e2e(runner, B, K, Wm, world, dev, graphed=not args.e2e_eager)

    # ---- encoder forward utt/s (BASELINE.json configs[1])
    enc = None
    if not args.no_encoder:
        try:
            enc = run_encoder_fwd(B, max(3, min(K, 20)), dev, world)
        except Exception as e:  # keep the headline line alive
            enc = {"error": repr(e)[:200]}

    stem_inf = None
    if not args.no_encoder and rank == 0:
        try:
            stem_inf = run_stem_infer(B, max(3, min(K, 50)), dev)
        except Exception as e:
            stem_inf = {"error": repr(e)[:200]}

    config1 = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            config1 = run_config1(dev, K)
        except Exception as e:
            config1 = {"error": repr(e)[:200]}

    stem_train = None
    if not args.no_encoder and rank == 0:
        try:
            stem_train = run_stem_train(B, max(3, min(K, 50)), dev)
        except Exception as e:
            stem_train = {"error": repr(e)[:200]}

    enc_train = None
    if not args.no_encoder:
        try:
            enc_train = run_encoder_train(B, max(3, min(K, 10)), dev, world, rank)
        except Exception as e:
            enc_train = {"error": repr(e)[:200]}

    # ---- clocks: make sure the sampler saw the workload for >= 1.5 s
    n_extra = min(200000, int(1.5 / max(1e-6, ms / K * 1e-3)))  # same count on every rank (ms is the max over ranks)
    for i in range(0, n_extra, 64):
        replay_steps(64)
        torch.cuda.synchronize()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else {}
    barrier(world)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(target_s=12.0)

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": round(ms / K, 5), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": make_config(B, world, nsets),
            "collective": collective, "kernels_per_step": own_per_step,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": gpu_launches, "clocks": clocks,
            "dp_parity": dp_parity,
            "step_roofline_frac": round(step_frac, 4), "kernels": kernels, "in_graph": in_graph,
            "calls_ms": {k: round(v, 5) for k, v in calls.items()},
            "config1_fwd": config1, "stem_infer": stem_inf, "stem_train": stem_train, "encoder_fwd": enc, "encoder_train": enc_train, "launch_count_check": _lib.launch_count() - launches0,
        }
        emit(line)
    if world > 1:
        torch.distributed.destroy_process_group()


def step_kernel_names(runner):
    """Kernel labels of one step in launch order, from the launch count of each C-ABI call (a layer's backward is either
    gy / adjoint / pre_conv^T / finalize or, for a small-batch data layer, fused / finalize)."""
    from qasr_ijcnlp_b200 import _lib

    names = []
    order = (("conv1", "fwd"), ("conv2", "fwd"), ("conv2", "bwd"), ("conv1", "bwd"))
    if runner.fused_fwd:
        order = (("stem", "fwd"), ("conv2", "bwd"), ("conv1", "bwd"))
    for layer, what in order:
        n0 = _lib.launch_count()
        if layer == "stem":
            runner.fwd_stem(0)
        else:
            (runner.fwd if what == "fwd" else runner.bwd)(layer, 0)
        n = _lib.launch_count() - n0
        if what == "fwd":
            kinds = ["fwd"] * n
        elif n == 4:
            kinds = ["bwd_post(gy)", "bwd_adj", "bwd_pre", "bwd_finalize"]
        elif n == 2:
            kinds = ["bwd_fused", "bwd_finalize"]
        else:
            kinds = [f"bwd_k{i}" for i in range(n)]
        names += [f"{layer}.{k}" for k in kinds]
    torch.cuda.synchronize()
    return names


def in_graph_timeline(runner, nsets, dev, B, peak, reps=20, world=1):
    """Per-kernel critical-path times inside the replayed step graph (qw_timeline_set: every kernel records min(CTA start) /
    max(CTA end) of %globaltimer).  delta = end of the kernel - end of its predecessor, i.e. what the kernel adds to the step
    with the programmatic-dependent-launch overlap it really has; median over `reps` single replays.  Explains `value`; the
    `roofline` object stays on the conservative per-launch event brackets."""
    from qasr_ijcnlp_b200 import _lib

    lib = _lib.load()
    n0 = _lib.launch_count()
    runner.step(0)
    per_step = _lib.launch_count() - n0
    torch.cuda.synchronize()
    buf = torch.zeros(2 * per_step, dtype=torch.int64, device=dev)
    side = torch.cuda.Stream()
    graphs = []
    try:
        for s_ in range(nsets):
            _lib.check(lib.qw_timeline_set(ctypes.c_void_p(buf.data_ptr()), per_step), "qw_timeline_set")  # slot counter -> 0
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                runner.step(s_)
            graphs.append(g)
        torch.cuda.synchronize()
        for i in range(2 * nsets):
            graphs[i % nsets].replay()
        rows = []
        for r in range(reps):
            buf[0::2] = torch.iinfo(torch.int64).max
            buf[1::2] = 0
            barrier(world)  # all ranks replay together (the step holds the gradient exchange)
            graphs[r % nsets].replay()
            torch.cuda.synchronize()
            t = buf.cpu().tolist()
            rows.append([(t[2 * k], t[2 * k + 1]) for k in range(per_step)])
    finally:
        lib.qw_timeline_set(None, 0)
    barrier(world)
    names = step_kernel_names(runner)
    # kernels without a timeline slot (the one-shot all-reduce kernel of the hybrid collective) leave their slot untouched
    used = [k for k in range(per_step) if rows[0][k][1] > 0]
    if len(used) == len(names):
        rows = [[r[k] for k in used] for r in rows]
        per_step = len(used)
    if per_step != len(names):
        names = [f"k{k}" for k in range(per_step)]
    algo = {"stem.fwd": algorithmic_bytes("qconv_fwd_kernel", "stem", B)}
    for layer in ("conv1", "conv2"):
        algo[f"{layer}.fwd"] = algorithmic_bytes("qconv_fwd_kernel", layer, B)
        algo[f"{layer}.bwd_post(gy)"] = algorithmic_bytes("qconv_bwd_post_kernel", layer, B)
        algo[f"{layer}.bwd_pre"] = algorithmic_bytes("qconv_bwd_pre_kernel", layer, B)
        algo[f"{layer}.bwd_fused"] = algorithmic_bytes("qconv_bwd_fused_kernel", layer, B)
    out = {}
    for k in range(per_step):
        d = statistics.median((r[k][1] - (r[k - 1][1] if k else r[0][0])) for r in rows) / 1e3
        ent = {"us": round(d, 2)}
        ab = algo.get(names[k])
        if ab and d > 0:
            ent["frac"] = round(ab / (d * 1e-6) / 1e9 / peak, 4)
        out[names[k]] = ent
    total = statistics.median(max(e for _, e in r) - r[0][0] for r in rows) / 1e3
    return {"step_us": round(total, 2), "kernels": out,
            "how": "median of %d single graph replays; us a kernel adds to the step = its last CTA end - its predecessor's last CTA "
                   "end (%%globaltimer); frac = algorithmic bytes / that time / HBM peak" % reps}


def host_topology(dev):
    """NUMA node of the GPU's PCIe function and of this process' allowed CPUs (the pinned staging buffers are first-touched by
    this process, so they live where it runs)."""
    out = {}
    try:
        bus = torch.cuda.get_device_properties(dev).pci_bus_id
        dom = torch.cuda.get_device_properties(dev).pci_domain_id
        devn = torch.cuda.get_device_properties(dev).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{devn:02x}.0/numa_node"
        out["gpu_numa_node"] = int(open(path).read().strip())
    except Exception:
        out["gpu_numa_node"] = None
    try:
        out["cpus_allowed"] = len(os.sched_getaffinity(0))
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        out["numa_nodes"] = len(nodes)
    except Exception:
        pass
    return out


def pin_to_gpu_numa_node(local):
    """Run this rank (and first-touch its pinned buffers) on the NUMA node its GPU hangs off, when the box has several."""
    try:
        p = torch.cuda.get_device_properties(local)
        path = f"/sys/bus/pci/devices/{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def runner_bytes(B):
    n = 0
    for cfg in LAYERS.values():
        n += B * cfg["C"] * cfg["L"] * 4 * (2 if cfg["need_gx"] else 1) + 2 * B * cfg["O"] * cfg["Lout"] * 4
    return n


def run_e2e(runner, B, K, Wm, world, dev, graphed=True):
    """Stem training step through the public nn.Module API from pinned host memory."""
    mods = runner.params.mods
    conv1, conv2 = mods["conv1"], mods["conv2"]
    params = list(conv1.parameters()) + list(conv2.parameters())
    nhost = 3
    g = torch.Generator().manual_seed(7)
    host_in = [torch.randn(B, N_MELS, N_FRAMES, generator=g).pin_memory() for _ in range(nhost)]
    n_grad = sum(p.numel() for p in params)
    host_out = [torch.empty(1 + n_grad).pin_memory() for _ in range(nhost)]
    dev_in = [torch.empty(B, N_MELS, N_FRAMES, device=dev) for _ in range(2)]
    copy_stream = torch.cuda.Stream()
    ev_copied = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    main = torch.cuda.current_stream()

    def issue_copy(i):
        slot = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_free[slot])
            dev_in[slot].copy_(host_in[i % nhost], non_blocking=True)
            ev_copied[slot].record(copy_stream)

    # the user-facing module, optionally wrapped by PyTorch's own CUDA-graph utility (torch.cuda.make_graphed_callables):
    # forward and backward of conv1 -> GELU -> conv2 replay as two graphs, which removes ~25 Python-side launches per step
    stem = torch.nn.Sequential(conv1, torch.nn.GELU(), conv2)
    api = "QuantumConv1d nn.Module x2 + GELU (nn.Sequential), eager autograd"
    call = stem
    stem_helper = False
    use_helper = os.environ.get("QW_E2E_PLAIN", "0") != "1"
    if use_helper:
        # the package's stem helper on the same two modules: conv1 -> GELU -> conv2 with the GELU inside conv1's kernels and the
        # gradient between the layers never written (stem_train_forward; falls back to the plain modules outside its regime)
        from qasr_ijcnlp_b200 import stem_train_forward

        class _Stem(torch.nn.Module):
            def __init__(self, c1, c2):
                super().__init__()
                self.conv1, self.conv2 = c1, c2

            def forward(self, x):
                return stem_train_forward(self.conv1, self.conv2, x, gelu=(True, False))

        stem = _Stem(conv1, conv2)
        call = stem
        stem_helper = True
        api = "stem_train_forward(conv1, conv2, x, gelu=(True, False)) on two QuantumConv1d nn.Modules, eager autograd"
    # data parallel: the two layers average their parameter gradients inside their own backward (qw_conv1d_backward_dp), so the
    # step needs no collective call at all; NCCL all_reduce of the flat gradient vector if symmetric memory is unavailable
    dp_note = ""
    fused_grads = False
    if world > 1:
        try:
            conv1.fuse_grad_allreduce()
            conv2.fuse_grad_allreduce()
            fused_grads = True
            dp_note = "; gradient mean over ranks fused into each layer's backward (NVLink peer memory; the stem helper passes the "\
                      "layers' contexts to qw_conv1d_backward_dp / _chained), loss kept per rank"
        except Exception as e:
            conv1._grad_allreduce = conv2._grad_allreduce = None
            dp_note = f"; NCCL all_reduce(AVG) of loss + gradients ({type(e).__name__})"

    def step_math(x):
        y2 = call(x)
        loss = y2.square().mean()
        grads = torch.autograd.grad(loss, params)
        return torch.cat([loss.detach().reshape(1)] + [g_.reshape(-1) for g_ in grads])

    # whole-step capture (the documented PyTorch pattern for static-shape training steps): forward, loss and backward of one
    # input slot replay as ONE CUDA graph, so the Python side of a step is an event wait, a replay and a D2H copy
    step_graphs = None
    if graphed:
        try:
            cap = torch.cuda.Stream()
            cap.wait_stream(main)
            with torch.cuda.stream(cap):
                for _ in range(3):
                    step_math(dev_in[0])
            main.wait_stream(cap)
            torch.cuda.synchronize()
            step_graphs = []
            for slot in range(2):
                gph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gph, stream=cap):
                    flat_static = step_math(dev_in[slot])
                step_graphs.append((gph, flat_static))
            torch.cuda.synchronize()
            api = (("stem_train_forward(conv1, conv2, x, gelu=(True, False)) on two QuantumConv1d nn.Modules" if stem_helper else
                    "QuantumConv1d nn.Module x2 + GELU (nn.Sequential)") +
                   ", loss and torch.autograd backward captured as one CUDA graph per input slot (torch.cuda.graph)")
        except Exception as e:
            step_graphs = None
            api += f" [whole-step capture failed: {type(e).__name__}: {str(e)[:80]}]"
            torch.cuda.synchronize()
            try:  # second choice: PyTorch's per-module graphing utility
                call = torch.cuda.make_graphed_callables(stem, (torch.randn(B, N_MELS, N_FRAMES, device=dev),))
                api += " -> torch.cuda.make_graphed_callables"
            except Exception as e2:
                call = stem
                api += f" -> eager ({type(e2).__name__})"

    def compute(i):
        slot = i % 2
        main.wait_event(ev_copied[slot])
        if step_graphs is not None:
            gph, flat = step_graphs[slot]
            gph.replay()
        else:
            flat = step_math(dev_in[slot])
        ev_free[slot].record(main)
        if world > 1 and not fused_grads:
            torch.distributed.all_reduce(flat, op=torch.distributed.ReduceOp.AVG)
        host_out[i % nhost].copy_(flat.detach(), non_blocking=True)

    def loop(n):
        for s in range(2):
            ev_free[s].record(main)
        issue_copy(0)
        for i in range(n):
            if i + 1 < n:
                issue_copy(i + 1)
            compute(i)

    loop(Wm)
    barrier(world)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    loop(K)
    b.record()
    torch.cuda.synchronize()
    ms = max_over_ranks(a.elapsed_time(b), world, dev)
    barrier(world)

    loss_value = float(host_out[(K - 1) % nhost][0])
    # copy-only control leg: the SAME loop with the compute taken out (H2D of every step's input from pinned memory, D2H of the
    # result vector): the ceiling the host link of this box puts on the e2e number at this GPU count
    flat_dummy = torch.zeros(1 + n_grad, device=dev)

    def copy_loop(n):
        for s_ in range(2):
            ev_free[s_].record(main)
        issue_copy(0)
        for i in range(n):
            if i + 1 < n:
                issue_copy(i + 1)
            slot = i % 2
            main.wait_event(ev_copied[slot])
            ev_free[slot].record(main)
            host_out[i % nhost].copy_(flat_dummy, non_blocking=True)

    copy_loop(Wm)
    barrier(world)
    a.record()
    copy_loop(K)
    b.record()
    torch.cuda.synchronize()
    ms_copy = max_over_ranks(a.elapsed_time(b), world, dev)
    barrier(world)
    h2d = B * N_MELS * N_FRAMES * 4
    val = world * B * WINDOWS_PER_UTT * K / (ms * 1e-3)
    return {"value": round(val, 1), "unit": UNIT, "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": 4 * (1 + n_grad), "ms_per_step": round(ms / K, 5),
            "copy_only_ms": round(ms_copy / K, 5), "h2d_GBps": round(h2d / (ms_copy / K * 1e-3) / 1e9, 2),
            "copy_bound_frac": round(ms_copy / ms, 4), "host": host_topology(dev),
            "api": api + " + MSE-style loss, torch.autograd, pinned host in/out, H2D double-buffered on a copy stream" +
                   dp_note,
            "loss": loss_value}


def run_encoder_fwd(B, K, dev, world):
    """BASELINE.json configs[1]: Quantum Whisper-Tiny encoder forward + 35-class head on Speech-Commands-shaped
    clips (1 s of 0.1*randn audio zero-padded to 30 s), log-mel computed on the GPU by qw_log_mel."""
    from qasr_ijcnlp_b200 import QuantumWhisper, QuantumWhisperClassifier, get_whisper_tiny_dims
    from qasr_ijcnlp_b200 import audio as qa

    torch.manual_seed(1)
    model = QuantumWhisperClassifier(QuantumWhisper(get_whisper_tiny_dims(), n_qubits=Q), 35).to(dev).eval()
    clips = (0.1 * torch.randn(B, 16000)).pin_memory()   # the 1 s clips as they come out of the dataset
    audio = torch.zeros(B, 480000)
    audio[:, :16000] = clips
    audio = audio.pin_memory()                            # ... and as the reference's Dataset.__getitem__ pads them (:62-77)

    def step(_):
        # batched data path: only the 16 000 stored samples cross PCIe; pad_or_trim happens inside qw_log_mel
        with torch.no_grad():
            mel = qa.log_mel_spectrogram(clips.to(dev, non_blocking=True), pad_to=480000)
            return model(mel)

    def step_host_padded(_):
        with torch.no_grad():
            mel = qa.log_mel_spectrogram(audio.to(dev, non_blocking=True))
            return model(mel)

    def front(_):
        return qa.log_mel_spectrogram(clips.to(dev, non_blocking=True), pad_to=480000)

    def front_host_padded(_):
        return qa.log_mel_spectrogram(audio.to(dev, non_blocking=True))

    with torch.no_grad():
        same = bool(torch.equal(front(0), front_host_padded(0)))
    for i in range(3):
        step(i), step_host_padded(i)
    ms = max_over_ranks(time_events(step, K), world, dev)
    ms_hp = max_over_ranks(time_events(step_host_padded, K), world, dev)
    ms_f, ms_fh = time_events(front, K) / K, time_events(front_host_padded, K) / K
    return {"value": round(world * B * K / (ms * 1e-3), 2), "unit": "utt/s", "ms_per_step": round(ms / K, 4),
            "host_padded": {"value": round(world * B * K / (ms_hp * 1e-3), 2), "ms_per_step": round(ms_hp / K, 4),
                            "h2d_bytes_per_step": B * 480000 * 4},
            "h2d_bytes_per_step": B * 16000 * 4,
            "front_end_ms": {"fused_pad": round(ms_f, 4), "host_padded": round(ms_fh, 4), "bit_identical": same,
                             "what": "H2D of the clips + log-mel only"},
            "workload": f"log-mel + QuantumWhisper-Tiny encoder fwd + Linear(384,35), batch {B}, host audio in: 1 s clips shipped "
                        f"as 16 000 samples, padded to 30 s inside qw_log_mel (host_padded = the reference's 480 000-sample rows)"}


def run_stem_infer(B, K, dev):
    """inference stem mel (B,80,3000) -> (B,1500,384): gelu(conv2(gelu(conv1(x)))).permute(0,2,1) + pos, fused
    (qw_stem_forward: two kernels, the (B,384,3000) intermediate never written) vs operator by operator (two QuantumConv1d
    forwards + torch GELU / permute / add).  Inputs resident in HBM, rotating sets."""
    import torch.nn.functional as F

    from qasr_ijcnlp_b200 import QuantumConv1d, fused_stem_forward
    from qasr_ijcnlp_b200.encoder import sinusoids

    torch.manual_seed(0)
    c1 = QuantumConv1d(N_MELS, N_STATE, kernel_size=3, padding=1, n_qubits=Q).to(dev)
    c2 = QuantumConv1d(N_STATE, N_STATE, kernel_size=3, stride=2, padding=1, n_qubits=Q).to(dev)
    pos = sinusoids(1500, N_STATE).to(dev)
    xs = [torch.rand(B, N_MELS, 3000, device=dev) * 3 - 1.5 for _ in range(4)]

    def fused(i):
        return fused_stem_forward(c1, c2, xs[i % 4], pos)

    def unfused(i):
        with torch.no_grad():
            return F.gelu(c2(F.gelu(c1(xs[i % 4])))).permute(0, 2, 1) + pos

    diff = (fused(0) - unfused(0)).abs().max().item()
    for i in range(3):
        fused(i), unfused(i)
    t_f, t_u = time_events(fused, K) / K, time_events(unfused, K) / K
    return {"fused_ms": round(t_f, 5), "unfused_ms": round(t_u, 5), "fused_utt_per_s": round(B / t_f * 1e3, 1),
            "unfused_utt_per_s": round(B / t_u * 1e3, 1), "max_abs_diff_fused_vs_unfused": diff,
            "workload": f"inference stem, batch {B}: mel (B,80,3000) -> (B,1500,384) incl. both GELUs, permute, positional embedding"}


def run_config1(dev, K):
    """BASELINE.json configs[0]: QuantumConv1d(80, 384, 3, padding=1, n_qubits=4) FORWARD ALONE on batch 2 x 80 mel x 3000 frames
    (6 000 windows) -- the reference's own CPU-runnable parity case (tests/test_qconv_gpu.py::test_config1_full_size_parity is
    its parity check).  GPU: the C-ABI forward on HBM-resident input, CUDA events.  CPU: the literal-loop restatement of the
    reference forward on a bounded number of output columns of the same input."""
    import torch.nn.functional as F  # noqa: F401

    from oracle import qconv_oracle as qo
    from qasr_ijcnlp_b200 import QuantumConv1d

    torch.manual_seed(0)
    m = QuantumConv1d(N_MELS, N_STATE, 3, padding=1, n_qubits=Q)
    x = torch.randn(2, N_MELS, N_FRAMES)
    md = m.to(dev)
    xs = [x.to(dev).clone() for _ in range(4)]
    with torch.no_grad():
        for i in range(4):
            md(xs[i])
        ms = time_events(lambda i: md(xs[i % 4]), max(K, 50)) / max(K, 50)
    p64 = [t.detach().cpu().double() for t in (m.pre_conv.weight, m.pre_conv.bias, m.quantum_weights, m.post_conv.weight, m.post_conv.bias)]
    cols = 24
    t0 = time.perf_counter()
    with torch.no_grad():
        qo.qconv1d_literal(x.double(), *p64, K=3, S=1, P=1, max_windows=cols)
    dt = time.perf_counter() - t0
    return {"gpu_ms": round(ms, 5), "gpu_windows_per_s": round(6000 / (ms * 1e-3), 1),
            "cpu_windows_per_s": round(2 * cols / dt, 2), "cpu_kind": _reference_kind(),
            "cpu_sample": f"literal-loop forward, first {cols} of 3000 output columns x batch 2 in {dt:.2f} s, torch threads={torch.get_num_threads()}",
            "workload": "QuantumConv1d(80,384,3,padding=1,n_qubits=4) forward alone, batch 2 x 80 x 3000 (6 000 windows), nn.Module call"}


def run_stem_train(B, K, dev, nsets=4):
    """SURVEY.md 8-f1 at training time: the encoder stem AS THE MODEL RUNS IT (whisper/whisper/model.py:193-194:
    gelu(conv1(x)), gelu(conv2(x))) forward + backward through the nn.Module API, inputs resident in HBM, whole step captured as
    one CUDA graph per input set.  `fused_gelu`: QuantumConv1d.forward_gelu (GELU in the forward epilogue, gelu' in the backward's gy
    pass: no activation pass of its own in either direction); `op_by_op`: the plain operators + ATen GELU (three extra passes over
    the (B,384,3000) / (B,384,1500) tensors per layer and direction-pair)."""
    import torch.nn.functional as F

    from qasr_ijcnlp_b200 import QuantumConv1d, stem_train_forward

    torch.manual_seed(0)
    c1 = QuantumConv1d(N_MELS, N_STATE, kernel_size=3, padding=1, n_qubits=Q).to(dev)
    c2 = QuantumConv1d(N_STATE, N_STATE, kernel_size=3, stride=2, padding=1, n_qubits=Q).to(dev)
    params = list(c1.parameters()) + list(c2.parameters())
    xs = [torch.rand(B, N_MELS, 3000, device=dev) * 3 - 1.5 for _ in range(nsets)]
    keys = {2: "one_kernel_forward", True: "fused_gelu", False: "op_by_op"}

    def step(fused, x):
        if fused == 2:    # what QuantumAudioEncoder.forward does in training: both layers + GELUs in one forward kernel
            y = stem_train_forward(c1, c2, x, gelu=True)
        else:
            y = c2.forward_gelu(c1.forward_gelu(x)) if fused else F.gelu(c2(F.gelu(c1(x))))
        loss = y.square().mean()
        return loss, torch.autograd.grad(loss, params)

    out = {}
    ref = None
    for fused in (True, 2, False):
        cap = torch.cuda.Stream()
        cap.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cap):
            for _ in range(3):
                step(fused, xs[0])
        torch.cuda.current_stream().wait_stream(cap)
        torch.cuda.synchronize()
        graphs = []
        for x in xs:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=cap):
                res = step(fused, x)
            graphs.append((g, res))
        torch.cuda.synchronize()
        for i in range(4):
            graphs[i % nsets][0].replay()
        ms = time_events(lambda i: graphs[i % nsets][0].replay(), K) / K
        graphs[0][0].replay()
        torch.cuda.synchronize()
        loss, grads = graphs[0][1]
        flat = torch.cat([loss.reshape(1)] + [t.reshape(-1) for t in grads]).clone()
        if ref is None:
            ref = flat
        out[keys[fused]] = {"ms_per_step": round(ms, 5), "utt_per_s": round(B / ms * 1e3, 1)}
        if fused is not True:
            err = float(((flat - ref).abs() / ref.abs().clamp(min=1.0)).max().detach())
            out["max_rel_diff_loss_and_grads"] = max(err, out.get("max_rel_diff_loss_and_grads", 0.0))
    out["speedup"] = round(out["op_by_op"]["ms_per_step"] / out["one_kernel_forward"]["ms_per_step"], 3)
    out["workload"] = (f"stem training step incl. both GELUs, batch {B}: mel (B,80,3000) -> gelu(conv1) -> gelu(conv2) -> loss, backward to all "
                       f"ten parameter gradients; nn.Module API, one CUDA graph per input set, {nsets} rotating sets")
    return out


def run_encoder_train(B, K, dev, world, rank):
    """BASELINE.json configs[2]: Quantum Whisper-Tiny + LSTM char decoder ASR training step on LibriSpeech-shaped synthetic
    30 s audio, batch B per GPU: host audio -> H2D -> log-mel (qw_log_mel) -> encoder (quantum stem + 4 blocks) -> char
    decoder -> CE(ignore 0) -> backward -> gradient all-reduce (NCCL, one flat bucket) -> clip 1.0 -> AdamW.
    Trainables as freeze_non_quantum_layers selects them (quantum_whisper.py:325-333): conv1, conv2, asr_head."""
    import qasr_ijcnlp_b200 as qw
    from qasr_ijcnlp_b200 import audio as qa
    from qasr_ijcnlp_b200 import dp

    torch.manual_seed(0)
    model = qw.QuantumWhisperASR(qw.QuantumWhisper(qw.get_whisper_tiny_dims(), n_qubits=Q)).to(dev)
    qw.freeze_non_quantum_layers(model)
    dp.broadcast_parameters(model)
    train = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(train, lr=1e-3, weight_decay=0.01)
    bucket = dp.GradBucket(train)
    g = torch.Generator().manual_seed(2 + rank)
    audio = (0.1 * torch.randn(B, 480000, generator=g)).pin_memory()
    V = len(qw.CHAR_VOCAB)
    tokens = torch.randint(4, V, (B, 100), generator=g)
    tokens[:, 0] = 2
    tokens = tokens.to(dev)
    lossf = torch.nn.CrossEntropyLoss(ignore_index=0)

    def step(_):
        mel = qa.log_mel_spectrogram(audio.to(dev, non_blocking=True))
        logits = model(mel, tokens[:, :-1])
        loss = lossf(logits.reshape(-1, V), tokens[:, 1:].reshape(-1))
        loss.backward()
        bucket.allreduce_mean()
        torch.nn.utils.clip_grad_norm_(train, 1.0)
        opt.step()
        for p in train:
            p.grad = None
        return loss

    for i in range(3):
        step(i)
    barrier(world)
    ms = max_over_ranks(time_events(step, K), world, dev)
    barrier(world)
    return {"value": round(world * B * K / (ms * 1e-3), 2), "unit": "utt/s", "ms_per_step": round(ms / K, 3),
            "trainable_floats": bucket.numel,
            "workload": f"ASR training step (30 s synthetic audio, batch {B}/GPU, quantum stem + 4 frozen blocks + LSTM char decoder), "
                        f"fwd+bwd+allreduce+clip+AdamW, host audio in"}


# --------------------------------------------------------------------------------------------- CPU legs
_PENNYLANE_REF = "unprobed"


def _reference_kind():
    """"reference" when the unmodified /root/reference QuantumConv1d can run here (PennyLane importable: SURVEY.md section 7 step 1),
    else "port" (the literal-loop restatement)."""
    global _PENNYLANE_REF
    if _PENNYLANE_REF == "unprobed":
        import oracle

        _PENNYLANE_REF = oracle.pennylane_reference()
    return "reference" if _PENNYLANE_REF is not None else "port"


def _literal_step(layers_cols, seed=0):
    """One bounded sample of the stem through the reference's own layer when PennyLane is importable, else through the
    literal-loop restatement of it (fwd + bwd)."""
    from oracle import qconv_oracle as qo

    n = 0
    if _reference_kind() == "reference":
        for name, cols in layers_cols.items():
            cfg = LAYERS[name]
            torch.manual_seed(seed)
            m = _PENNYLANE_REF(cfg["C"], cfg["O"], cfg["K"], stride=cfg["S"], padding=cfg["P"], n_qubits=Q)
            Lneed = (cols - 1) * cfg["S"] + cfg["K"] - 2 * cfg["P"]
            x = torch.randn(2, cfg["C"], max(1, Lneed), requires_grad=cfg["need_gx"])
            y = m(x)
            y.square().sum().backward()
            n += 2 * y.shape[-1]
        return n
    for name, cols in layers_cols.items():
        cfg = LAYERS[name]
        params = [p.requires_grad_(True) for p in qo.make_params(cfg["C"], cfg["O"], cfg["K"], Q, seed=seed)]
        g = torch.Generator().manual_seed(seed + 1)
        Lneed = cols * cfg["S"] + cfg["K"]
        x = torch.randn(2, cfg["C"], Lneed, generator=g, dtype=torch.float64, requires_grad=cfg["need_gx"])
        y = qo.qconv1d_literal(x, *params, K=cfg["K"], S=cfg["S"], P=cfg["P"], max_windows=cols)
        y[:, :, :cols].square().sum().backward()
        n += 2 * cols
    return n


def cpu_baseline(target_s=12.0):
    cores = torch.get_num_threads()
    t0 = time.perf_counter()
    n = _literal_step({"conv1": 16, "conv2": 8})
    dt = time.perf_counter() - t0
    rate = n / dt
    scale = max(1, int(target_s * rate / 48))
    cols = {"conv1": 16 * scale, "conv2": 8 * scale}
    t0 = time.perf_counter()
    n = _literal_step(cols)
    dt = time.perf_counter() - t0
    out = {"value": round(n / dt, 2), "unit": UNIT, "cores": cores, "kind": _reference_kind(),
           "sample": f"literal-loop fp64 restatement of quantum_whisper.py:107-126 (PennyLane unavailable), fwd+bwd, "
                     f"batch 2, first {cols['conv1']} conv1 + {cols['conv2']} conv2 output columns = {n} windows in {dt:.1f} s; "
                     f"torch threads={cores}, os.cpu_count()={os.cpu_count()} (the loop is scalar Python: ~1 core busy)"}
    # stronger secondary CPU baseline: vectorised fp64 oracle (unfold + batched statevector), all torch threads
    try:
        from oracle import qconv_oracle as qo

        cfg = LAYERS["conv1"]
        params = [p.requires_grad_(True) for p in qo.make_params(cfg["C"], cfg["O"], 3, Q, seed=0)]
        x = torch.randn(2, cfg["C"], 3000, dtype=torch.float64)
        t0 = time.perf_counter()
        y = qo.qconv1d_forward(x, *params, K=3, S=1, P=1)
        y.square().sum().backward()
        dtv = time.perf_counter() - t0
        out["vectorised_oracle"] = {"value": round(6000 / dtv, 1), "unit": UNIT, "cores": cores,
                                    "sample": f"conv1 geometry, batch 2 x 3000 windows fwd+bwd in {dtv:.2f} s"}
    except Exception as e:
        out["vectorised_oracle"] = {"error": repr(e)[:200]}
    return out


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    K, Wm = args.steps, args.warmup
    # torchrun exports OMP_NUM_THREADS=1 for every rank; the other ranks have exited, so rank 0 may use the whole host
    try:
        torch.set_num_threads(max(torch.get_num_threads(), len(os.sched_getaffinity(0))))
    except (AttributeError, RuntimeError):
        pass
    cores = torch.get_num_threads()
    cols = {"conv1": 32, "conv2": 16}
    # keep the whole run within a few minutes: calibrate the per-step sample on the first warm-up step
    t0 = time.perf_counter()
    n = _literal_step(cols)
    dt = time.perf_counter() - t0
    budget = 150.0 / max(1, K + Wm)
    if dt > budget:
        f = max(1, int(32 * budget / dt))
        cols = {"conv1": 2 * max(1, f // 2), "conv2": max(1, f // 2)}
    for _ in range(max(0, Wm - 1)):
        _literal_step(cols)
    t0 = time.perf_counter()
    n = 0
    for i in range(K):
        n += _literal_step(cols, seed=i)
    dt = time.perf_counter() - t0
    val = n / dt
    sample = (f"literal-loop fp64 restatement of quantum_whisper.py:107-126 (PennyLane unavailable), fwd+bwd, batch 2, "
              f"{cols['conv1']} conv1 + {cols['conv2']} conv2 output columns per step = {n // max(1, K)} windows/step: a bounded "
              f"sample of the batch-{args.batch} workload (the cost per window is constant in the reference's Python loop)")
    line = {
        "impl": "reference", "metric": METRIC, "value": round(val, 2), "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": Wm, "ms_per_step": round(dt / max(1, K) * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": make_config(args.batch, world, args.nsets),
        "cpu_baseline": {"value": round(val, 2), "unit": UNIT, "cores": cores, "kind": _reference_kind(), "sample": sample},
        "e2e": {"value": round(val, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def main():
    args = parse_args()
    # the contract is ONE JSON line on stdout: libraries (NCCL prints its version banner) write to fd 1 too, so point fd 1 at
    # stderr for the duration of the run and emit the line through a private duplicate of the real stdout
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
