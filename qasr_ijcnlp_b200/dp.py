"""Batch-sharded data parallelism for the quantum stem.

The reference has no distributed code at all (single process, `train_quantum_whisper.py:195-214`); `north_star`
asks for the batch of utterances to be partitioned across the GPUs of one box: windows and utterances are
independent, so forward / inference need NO collective, and a training step needs exactly one gradient
all-reduce (NCCL over NVLink/NVSwitch; `gloo` in the CPU tests) over the trainable parameters
(`freeze_non_quantum_layers` regime: 2 x QuantumConv1d = 9 440 floats + the task head).

One process per GPU; this module only holds the plumbing (shard arithmetic, one flat gradient bucket,
parameter broadcast).  The gradient partials of the quantum layers are already reduced to their final
per-GPU value inside the backward kernels' finalize step, so the bucket is ready as soon as autograd returns.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of the utterances rank `rank` owns: contiguous, sizes differ by at most one, the first
    `n_items % world` ranks get the extra item.  Integer arithmetic only."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Make every rank start from rank `src`'s parameters and buffers (one flat broadcast per dtype)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    tensors = [p.data for p in module.parameters()] + [b.data for b in module.buffers()]
    by_dtype = {}
    for t in tensors:
        by_dtype.setdefault((t.dtype, t.device), []).append(t)
    for (_, _), ts in by_dtype.items():
        flat = torch.cat([t.reshape(-1) for t in ts])
        dist.broadcast(flat, src=src, group=group)
        off = 0
        for t in ts:
            t.copy_(flat[off:off + t.numel()].view_as(t))
            off += t.numel()


class GradBucket:
    """One flat fp32 buffer holding the gradients of the trainable parameters, all-reduced once per step.

    `allreduce_mean()` packs `p.grad` (zeros for parameters that received none), issues a single
    `all_reduce(SUM)`, scales by 1/world and points every `p.grad` at its slice of the bucket (so the optimizer
    reads the reduced values with no copy back)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        for p in self.params:
            if p.device != dev or p.dtype != dt:
                raise ValueError("GradBucket needs all trainable parameters on one device with one dtype")
        self.group = group
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, device=dev, dtype=dt)
        self.views = []
        off = 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    @property
    def nbytes(self) -> int:
        return self.numel * self.flat.element_size()

    def pack(self, grads: Optional[Iterable[Optional[torch.Tensor]]] = None) -> None:
        src = [p.grad for p in self.params] if grads is None else list(grads)
        for v, g in zip(self.views, src):
            if g is None:
                v.zero_()
            elif g.data_ptr() != v.data_ptr():
                v.copy_(g)

    def allreduce_mean(self, grads: Optional[Iterable[Optional[torch.Tensor]]] = None, async_op: bool = False):
        self.pack(grads)
        world = dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1
        work = None
        if world > 1:
            work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
            if not async_op:
                self.flat.mul_(1.0 / world)
        for p, v in zip(self.params, self.views):
            p.grad = v
        if async_op and work is not None:
            return _Pending(work, self.flat, world)
        return None


class _Pending:
    def __init__(self, work, flat, world):
        self.work, self.flat, self.world = work, flat, world

    def wait(self):
        self.work.wait()
        self.flat.mul_(1.0 / self.world)


class FusedLayerGradAllReduce:
    """Per-layer context of `qw_conv1d_backward_dp`: the mean over ranks of a QuantumConv1d layer's five parameter gradients is
    taken INSIDE the backward's last kernel (every CTA of the finalize kernel stores (epoch, value) words straight into the peers'
    receive buffers over NVLink and polls its own: one write latency, no fence, no separate collective, no bucket packing;
    CUDA-graph capturable; bitwise identical on every rank).

    Owns the layer's symmetric data / flag buffers (sizes depend on C, O, n_layers only).  torch symmetric memory does the
    rendezvous.  world == 1 degenerates to the plain backward.  Attach to a module with `QuantumConv1d.fuse_grad_allreduce()`
    or pass `.args()` to the raw entry point."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int, stride: int, padding: int, n_qubits: int = 4,
                 n_layers: int = 1, device=None, group=None):
        import ctypes

        from . import _lib

        self._ctypes = ctypes
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else "cuda")
        inited = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if inited else 1
        self.rank = dist.get_rank(group) if inited else 0
        dims = (1, in_channels, 64, kernel_size, stride, padding, out_channels, n_qubits, n_layers)
        nbuf = self.lib.qw_conv1d_dp_buffer_bytes(*dims, self.world) // 4
        nflag = self.lib.qw_conv1d_dp_flag_bytes(*dims, self.world) // 4
        if nbuf == 0 or nflag == 0:
            raise ValueError("the fused gradient all-reduce needs n_qubits == 4 and a valid layer geometry")
        self.nbuf, self.nflag = nbuf, nflag
        if self.world == 1:
            self.buf = torch.zeros(nbuf, device=self.device, dtype=torch.float32)
            self.flags = torch.zeros(nflag, device=self.device, dtype=torch.int32)
            buf_ptrs, flag_ptrs = [self.buf.data_ptr()], [self.flags.data_ptr()]
        else:
            import torch.distributed._symmetric_memory as symm_mem

            grp = group if group is not None else dist.group.WORLD
            self.buf = symm_mem.empty(nbuf, dtype=torch.float32, device=self.device)
            self.flags = symm_mem.empty(nflag, dtype=torch.int32, device=self.device)
            self.buf.zero_()
            self.flags.zero_()
            self._hb = symm_mem.rendezvous(self.buf, grp)
            self._hf = symm_mem.rendezvous(self.flags, grp)
            buf_ptrs, flag_ptrs = list(self._hb.buffer_ptrs), list(self._hf.buffer_ptrs)
            torch.cuda.synchronize(self.device)
            dist.barrier(group)  # every rank's flags are zero before anybody signals
        P = ctypes.c_void_p
        self._bufs = (P * self.world)(*[P(p) for p in buf_ptrs])
        self._flags = (P * self.world)(*[P(p) for p in flag_ptrs])

    def args(self):
        """(peer_bufs, peer_flags, rank, world, scale) for qw_conv1d_backward_dp."""
        return self._bufs, self._flags, self.rank, self.world, self._ctypes.c_float(1.0 / self.world)

    def status(self) -> int:
        """0 ok; 1 if a backward trapped after waiting DP_TIMEOUT_MS for a peer (the CUDA context is dead by then: this is only
        readable from a fresh mapping; kept for diagnostics).  Synchronises."""
        return int(self.flags[self.nflag - 1].item())


class P2PGradAllReduce:
    """One-shot NVLink peer-memory all-reduce (mean) of a small flat fp32 gradient bucket: `qw_grads_allreduce_p2p`.

    torch symmetric memory does the rendezvous (it maps every rank's buffer into every process); the data movement and
    the reduction are the library's own kernel -- no NCCL call on the step's critical path, and the launch is CUDA-graph
    capturable.  world == 1 works with plain device tensors (used by the single-GPU test).  Raises if symmetric memory is
    unavailable; callers then fall back to `GradBucket` (NCCL)."""

    def __init__(self, numel: int, device: torch.device, group=None):
        import ctypes

        from . import _lib

        self._ctypes, self._lib_mod = ctypes, _lib
        self.lib = _lib.load()
        self.numel = int(numel)
        self.device = torch.device(device)
        inited = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if inited else 1
        self.rank = dist.get_rank(group) if inited else 0
        nbuf = self.lib.qw_grads_allreduce_p2p_buffer_bytes(self.numel, self.world) // 4
        nflag = self.lib.qw_grads_allreduce_p2p_flag_bytes(self.world) // 4
        self.nflag = nflag
        if self.world == 1:
            self.buf = torch.zeros(nbuf, device=self.device, dtype=torch.float32)
            self.flags = torch.zeros(nflag, device=self.device, dtype=torch.int32)
            buf_ptrs, flag_ptrs = [self.buf.data_ptr()], [self.flags.data_ptr()]
        else:
            import torch.distributed._symmetric_memory as symm_mem

            grp = group if group is not None else dist.group.WORLD
            self.buf = symm_mem.empty(nbuf, dtype=torch.float32, device=self.device)
            self.flags = symm_mem.empty(nflag, dtype=torch.int32, device=self.device)
            self.buf.zero_()
            self.flags.zero_()
            self._hb = symm_mem.rendezvous(self.buf, grp)
            self._hf = symm_mem.rendezvous(self.flags, grp)
            buf_ptrs, flag_ptrs = list(self._hb.buffer_ptrs), list(self._hf.buffer_ptrs)
            torch.cuda.synchronize(self.device)
            dist.barrier(group)  # every rank's flags are zero before anybody signals
        P = ctypes.c_void_p
        self._bufs = (P * self.world)(*[P(p) for p in buf_ptrs])
        self._flags = (P * self.world)(*[P(p) for p in flag_ptrs])

    def __call__(self, flat: torch.Tensor) -> torch.Tensor:
        """In-place mean over ranks of `flat` (contiguous fp32, numel == self.numel) on the current stream."""
        if flat.numel() != self.numel or flat.dtype != torch.float32 or not flat.is_contiguous():
            raise ValueError("bucket must be a contiguous fp32 tensor of the size given at construction")
        P = self._ctypes.c_void_p
        with torch.cuda.device(self.device):
            st = self.lib.qw_grads_allreduce_p2p(P(flat.data_ptr()), self.numel, self._bufs, self._flags, self.rank, self.world,
                                                 1.0 / self.world, P(torch.cuda.current_stream().cuda_stream))
        self._lib_mod.check(st, "qw_grads_allreduce_p2p")
        return flat

    def epoch(self) -> int:
        """Number of completed calls (read from the first chunk's epoch word; synchronises)."""
        return int(self.flags[self.nflag - 1 - 64].item())

    def status(self) -> int:
        """0 ok; 1 if a call trapped after waiting DP_TIMEOUT_MS for a peer (diagnostics; synchronises)."""
        return int(self.flags[self.nflag - 1].item())
