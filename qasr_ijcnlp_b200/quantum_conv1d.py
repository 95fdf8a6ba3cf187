"""Drop-in `QuantumConv1d` (host mirror of /root/reference/quantum_whisper.py:45-128).

Same constructor, attributes, forward signature and state_dict layout as the reference:
``pre_conv.weight (q, C*K)``, ``pre_conv.bias (q)``, ``post_conv.weight (O, q)``, ``post_conv.bias (O)``,
``quantum_weights (q, 3)``.  The forward/backward are hand-written sm_100a kernels reached through the C ABI
in ``include/qw.h`` (no PennyLane, no Python loop, no CPU fallback).

Extensions (keyword-only, defaults == the reference): ``n_layers`` (``quantum_weights`` becomes
``(n_layers, q, 3)`` when > 1) and ``embedding`` ("amplitude" | "angle").
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn

from . import _lib

EMBEDDINGS = {"amplitude": 0, "angle": 1}


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def out_length(L: int, K: int, S: int, P: int) -> int:
    """quantum_whisper.py:103"""
    return (L + 2 * P - K) // S + 1


class _QuantumConv1dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w_pre, b_pre, qw, w_post, b_post, K, S, P, n_layers, emb, dpctx=None):
        lib = _lib.load()
        B, C, L = x.shape
        O, q = w_post.shape
        Lout = out_length(L, K, S, P)
        if Lout <= 0:
            raise ValueError(f"kernel_size {K} does not fit the padded input length {L + 2 * P}")
        f64 = x.dtype == torch.float64
        x = x.contiguous()
        params = [t.contiguous() for t in (w_pre, b_pre, qw, w_post, b_post)]
        y = torch.empty(B, O, Lout, device=x.device, dtype=x.dtype)
        need_bwd = any(ctx.needs_input_grad[:6])
        general = q > 4 or emb != 0  # composed path: pre_save doubles as the scratch between its kernels
        pre_save = torch.empty(2, B * Lout, q, device=x.device, dtype=x.dtype) if (need_bwd or general) else None
        fn = lib.qw_conv1d_forward_f64 if f64 else lib.qw_conv1d_forward
        with torch.cuda.device(x.device):
            st = fn(_ptr(x), *[_ptr(p) for p in params], _ptr(y), _ptr(pre_save),
                    B, C, L, K, S, P, O, q, n_layers, emb, _stream())
        _lib.check(st, "qw_conv1d_forward")
        if need_bwd:
            ctx.save_for_backward(x, pre_save, *params)
            ctx.cfg = (B, C, L, K, S, P, O, q, n_layers, emb, f64)
            ctx.dpctx = dpctx
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = _lib.load()
        x, pre_save, w_pre, b_pre, qw, w_post, b_post = ctx.saved_tensors
        B, C, L, K, S, P, O, q, n_layers, emb, f64 = ctx.cfg
        gy = gy.contiguous()
        dev, dt = x.device, x.dtype
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gw_pre = torch.empty_like(w_pre)
        gb_pre = torch.empty_like(b_pre)
        gqw = torch.empty_like(qw)
        gw_post = torch.empty_like(w_post)
        gb_post = torch.empty_like(b_post)
        nbytes = lib.qw_conv1d_workspace_bytes(B, C, L, K, S, P, O, q, n_layers, 8 if f64 else 4)
        ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        dpctx = getattr(ctx, "dpctx", None)
        with torch.cuda.device(dev):
            if dpctx is not None and dpctx.world > 1:
                if f64:
                    raise RuntimeError("the fused gradient all-reduce is fp32 only")
                # parameter gradients come back already averaged over the ranks (qw_conv1d_backward_dp); grad_x stays local
                st = lib.qw_conv1d_backward_dp(_ptr(gy), _ptr(x), _ptr(pre_save), _ptr(w_pre), _ptr(qw), _ptr(w_post), _ptr(gx),
                                               _ptr(gw_pre), _ptr(gb_pre), _ptr(gqw), _ptr(gw_post), _ptr(gb_post), _ptr(ws), nbytes,
                                               B, C, L, K, S, P, O, q, n_layers, emb, *dpctx.args(), _stream())
            else:
                fn = lib.qw_conv1d_backward_f64 if f64 else lib.qw_conv1d_backward
                st = fn(_ptr(gy), _ptr(x), _ptr(pre_save), _ptr(w_pre), _ptr(qw), _ptr(w_post), _ptr(gx), _ptr(gw_pre),
                        _ptr(gb_pre), _ptr(gqw), _ptr(gw_post), _ptr(gb_post), _ptr(ws), nbytes,
                        B, C, L, K, S, P, O, q, n_layers, emb, _stream())
        _lib.check(st, "qw_conv1d_backward")
        return gx, gw_pre, gb_pre, gqw, gw_post, gb_post, None, None, None, None, None, None


class _QuantumConv1dGeluFn(torch.autograd.Function):
    """y = gelu(QuantumConv1d(x)) with the activation fused into the layer's kernels (qw_conv1d_forward_act / _backward_act):
    the pre-activation tensor never exists in HBM and neither GELU pass (forward, backward) runs on its own."""

    @staticmethod
    def forward(ctx, x, w_pre, b_pre, qw, w_post, b_post, K, S, P, n_layers):
        lib = _lib.load()
        B, C, L = x.shape
        O, q = w_post.shape
        Lout = out_length(L, K, S, P)
        x = x.contiguous()
        params = [t.contiguous() for t in (w_pre, b_pre, qw, w_post, b_post)]
        y = torch.empty(B, O, Lout, device=x.device, dtype=x.dtype)
        need_bwd = any(ctx.needs_input_grad[:6])
        pre_save = torch.empty(2, B * Lout, q, device=x.device, dtype=x.dtype) if need_bwd else None
        with torch.cuda.device(x.device):
            st = lib.qw_conv1d_forward_act(_ptr(x), *[_ptr(p) for p in params], _ptr(y), _ptr(pre_save),
                                           B, C, L, K, S, P, O, q, n_layers, 0, 1, _stream())
        _lib.check(st, "qw_conv1d_forward_act")
        if need_bwd:
            ctx.save_for_backward(x, pre_save, *params)
            ctx.cfg = (B, C, L, K, S, P, O, q, n_layers)
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = _lib.load()
        x, pre_save, w_pre, b_pre, qw, w_post, b_post = ctx.saved_tensors
        B, C, L, K, S, P, O, q, n_layers = ctx.cfg
        gy = gy.contiguous()
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        grads = [torch.empty_like(t) for t in (w_pre, b_pre, qw, w_post, b_post)]
        nbytes = lib.qw_conv1d_workspace_bytes(B, C, L, K, S, P, O, q, n_layers, 4)
        ws = torch.empty(nbytes, device=x.device, dtype=torch.uint8)
        with torch.cuda.device(x.device):
            st = lib.qw_conv1d_backward_act(_ptr(gy), _ptr(x), _ptr(pre_save), _ptr(w_pre), _ptr(qw), _ptr(w_post), _ptr(b_post), _ptr(gx),
                                            *[_ptr(g) for g in grads], _ptr(ws), nbytes, B, C, L, K, S, P, O, q, n_layers, 0, 1, _stream())
        _lib.check(st, "qw_conv1d_backward_act")
        return (gx, *grads, None, None, None, None)


def gelu_fusion_eligible(x, w_post, kernel_size, stride, padding, n_layers, embedding) -> bool:
    """Shapes the activation-fused kernels cover (the fast-path regime of csrc/qw_conv1d_fast.cu, forward needs padding 1)."""
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 3 and embedding == "amplitude"):
        return False
    O, q = w_post.shape
    L = x.shape[2]
    Lout = out_length(L, kernel_size, stride, padding)
    return (q == 4 and kernel_size == 3 and stride in (1, 2) and padding == 1 and n_layers <= 4 and L % 4 == 0 and Lout % 4 == 0
            and O % 4 == 0 and O <= 576 and x.shape[1] * 3 * 16 <= 96 * 1024 and _lib.load().qw_get_option(b"GY_MMA") != 0
            and _lib.load().qw_get_option(b"FAST_PATH") != 0)


def quantum_conv1d(x, w_pre, b_pre, quantum_weights, w_post, b_post, kernel_size, stride=1, padding=0,
                   n_layers=1, embedding="amplitude", grad_allreduce=None):
    """Functional form.  x: (B, C, L) CUDA float32 (or float64 for the validation build)."""
    if x.dim() != 3:
        raise ValueError(f"expected input of shape (batch, channels, length), got {tuple(x.shape)}")
    if not x.is_cuda:
        raise RuntimeError("QuantumConv1d (B200 build) has no CPU path: move the module and input to a CUDA device")
    if x.dtype not in (torch.float32, torch.float64):
        raise ValueError(f"QuantumConv1d supports float32 (and float64 validation) inputs, got {x.dtype}")
    q = w_post.shape[1]
    C = x.shape[1]
    if w_pre.shape != (q, C * kernel_size):
        raise ValueError(f"pre_conv.weight has shape {tuple(w_pre.shape)}, expected {(q, C * kernel_size)} "
                         f"for {C} input channels and kernel_size {kernel_size}")
    for name, p in (("pre_conv.weight", w_pre), ("pre_conv.bias", b_pre), ("quantum_weights", quantum_weights),
                    ("post_conv.weight", w_post), ("post_conv.bias", b_post)):
        if p.device != x.device or p.dtype != x.dtype:
            raise RuntimeError(f"{name} is {p.dtype} on {p.device} but the input is {x.dtype} on {x.device}")
    if quantum_weights.numel() != n_layers * q * 3:
        raise ValueError(f"quantum_weights has {quantum_weights.numel()} elements, expected {n_layers}*{q}*3")
    return _QuantumConv1dFn.apply(x, w_pre, b_pre, quantum_weights, w_post, b_post, int(kernel_size), int(stride),
                                  int(padding), int(n_layers), EMBEDDINGS[embedding], grad_allreduce)


class QuantumConv1d(nn.Module):
    """Quantum 1D convolution replacing Whisper's Conv1d (quantum_whisper.py:45-128)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int, stride: int = 1, padding: int = 0,
                 n_qubits: int = 4, *, n_layers: int = 1, embedding: str = "amplitude"):
        super().__init__()
        if embedding not in EMBEDDINGS:
            raise ValueError(f"embedding must be one of {sorted(EMBEDDINGS)}, got {embedding!r}")
        if n_layers < 1:
            raise ValueError("n_layers must be >= 1")
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = kernel_size
        self.stride = stride
        self.padding = padding
        self.n_qubits = min(n_qubits, in_channels * kernel_size)  # :55
        self.n_layers = n_layers
        self.embedding = embedding
        # same construction order as the reference so the same seed gives the same parameters (:58-59,:88)
        self.pre_conv = nn.Linear(in_channels * kernel_size, self.n_qubits)
        self.post_conv = nn.Linear(self.n_qubits, out_channels)
        shape = (self.n_qubits, 3) if n_layers == 1 else (n_layers, self.n_qubits, 3)
        self.quantum_weights = nn.Parameter(torch.randn(*shape))
        self._grad_allreduce = None  # dp.FusedLayerGradAllReduce: parameter gradients leave backward already averaged

    def fuse_grad_allreduce(self, group=None):
        """Data-parallel training: average this layer's parameter gradients over the process group INSIDE its backward
        (`qw_conv1d_backward_dp`, one NVLink round trip in the last kernel) instead of in a separate collective.  Call once,
        on every rank, after `.to(device)`.  The parameters must then be left out of any other gradient all-reduce.  Needs the
        fast-path regime (n_qubits=4, amplitude embedding, k=3, stride 1|2, fp32)."""
        from . import dp

        self._grad_allreduce = dp.FusedLayerGradAllReduce(self.in_channels, self.out_channels, self.kernel_size, self.stride,
                                                          self.padding, self.n_qubits, self.n_layers,
                                                          device=self.pre_conv.weight.device, group=group)
        return self

    def to(self, *args, **kwargs):  # the reference overrides .to(device) and returns self (:90-93)
        super().to(*args, **kwargs)
        return self

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() == 3 and x.shape[1] != self.in_channels:
            raise ValueError(f"expected {self.in_channels} input channels, got {x.shape[1]}")
        return quantum_conv1d(x, self.pre_conv.weight, self.pre_conv.bias, self.quantum_weights,
                              self.post_conv.weight, self.post_conv.bias, self.kernel_size, self.stride,
                              self.padding, self.n_layers, self.embedding, self._grad_allreduce)

    def forward_gelu(self, x: torch.Tensor) -> torch.Tensor:
        """``F.gelu(self(x))`` -- what the encoder does with both stem layers (whisper/whisper/model.py:193-194) -- with the GELU
        fused into the layer's forward epilogue and its derivative into the backward's gy pass when the shape is in the fast-path
        regime (same result up to the 3e-7 error of the fused erf GELU; saves two full passes over the (B, O, L_out) tensor in
        the forward and three in the backward); otherwise exactly ``F.gelu(self(x))``."""
        if (self._grad_allreduce is None and x.dim() == 3 and x.shape[1] == self.in_channels
                and gelu_fusion_eligible(x, self.post_conv.weight, self.kernel_size, self.stride, self.padding, self.n_layers, self.embedding)):
            return _QuantumConv1dGeluFn.apply(x, self.pre_conv.weight, self.pre_conv.bias, self.quantum_weights, self.post_conv.weight,
                                              self.post_conv.bias, self.kernel_size, self.stride, self.padding, self.n_layers)
        return torch.nn.functional.gelu(self.forward(x))

    def extra_repr(self) -> str:
        return (f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, stride={self.stride}, "
                f"padding={self.padding}, n_qubits={self.n_qubits}, n_layers={self.n_layers}, "
                f"embedding={self.embedding!r}")


class _StemTrainFn(torch.autograd.Function):
    """y2 = act2(conv2(act1(conv1(x)))) for TRAINING: one forward kernel for both layers (qw_stem_train_forward: the (B, hidden, L)
    activation is written for the backward but never read back by the forward), then the CHAINED backward: conv2's
    qw_conv1d_backward_act / _dp without grad_x, conv1's qw_conv1d_backward_chained, whose gy kernel rebuilds conv2's input gradient
    from conv2's gpre rows (the STEM_CHAIN option set to 0 restores the two independent backward calls with grad_x through HBM)."""

    @staticmethod
    def forward(ctx, x, act, n_layers, dpctx, *params):
        lib = _lib.load()
        act1, act2 = (bool(act[0]), bool(act[1])) if isinstance(act, tuple) else (bool(act), bool(act))
        ctx.dpctx = dpctx  # None, or (conv1's, conv2's) dp.FusedLayerGradAllReduce: gradients leave the backward averaged over the ranks
        B, C, L = x.shape
        p1 = [t.contiguous() for t in params[:5]]
        p2 = [t.contiguous() for t in params[5:]]
        H, O = p1[3].shape[0], p2[3].shape[0]
        x = x.contiguous()
        y1 = torch.empty(B, H, L, device=x.device, dtype=x.dtype)
        y2 = torch.empty(B, O, L // 2, device=x.device, dtype=x.dtype)
        ps1 = torch.empty(2, B * L, 4, device=x.device, dtype=x.dtype)
        ps2 = torch.empty(2, B * (L // 2), 4, device=x.device, dtype=x.dtype)
        with torch.cuda.device(x.device):
            if act1 == act2 and lib.qw_stem_train_forward_preferred(B, L):
                st = lib.qw_stem_train_forward(_ptr(x), *[_ptr(t) for t in p1], *[_ptr(t) for t in p2], _ptr(y1), _ptr(ps1), _ptr(y2),
                                               _ptr(ps2), B, C, L, H, O, n_layers, 1 if act1 else 0, _stream())
                _lib.check(st, "qw_stem_train_forward")
            else:  # many tiles per CTA (two leaner forward kernels win) or different activations after the two layers: the saved
                   # tensors are the same, so is the chained backward
                st = lib.qw_conv1d_forward_act(_ptr(x), *[_ptr(t) for t in p1], _ptr(y1), _ptr(ps1), B, C, L, 3, 1, 1, H, 4, n_layers, 0,
                                               1 if act1 else 0, _stream())
                _lib.check(st, "qw_conv1d_forward_act")
                st = lib.qw_conv1d_forward_act(_ptr(y1), *[_ptr(t) for t in p2], _ptr(y2), _ptr(ps2), B, H, L, 3, 2, 1, O, 4, n_layers, 0,
                                               1 if act2 else 0, _stream())
                _lib.check(st, "qw_conv1d_forward_act")
        ctx.save_for_backward(x, y1, ps1, ps2, *p1, *p2)
        ctx.cfg = (B, C, L, H, O, n_layers, act1, act2)
        return y2

    @staticmethod
    def backward(ctx, gy2):
        lib = _lib.load()
        x, y1, ps1, ps2, *pp = ctx.saved_tensors
        p1, p2 = pp[:5], pp[5:]
        B, C, L, H, O, n_layers, act1, act2 = ctx.cfg
        dev = x.device

        def layer_backward(gy, xin, ps, prm, Cin, Lin, S, Oout, need_gx, act):
            w_pre, b_pre, qw, w_post, b_post = prm
            gx = torch.empty_like(xin) if need_gx else None
            grads = [torch.empty_like(t) for t in prm]
            nbytes = lib.qw_conv1d_workspace_bytes(B, Cin, Lin, 3, S, 1, Oout, 4, n_layers, 4)
            ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
            with torch.cuda.device(dev):
                st = lib.qw_conv1d_backward_act(_ptr(gy), _ptr(xin), _ptr(ps), _ptr(w_pre), _ptr(qw), _ptr(w_post), _ptr(b_post), _ptr(gx),
                                                *[_ptr(g) for g in grads], _ptr(ws), nbytes, B, Cin, Lin, 3, S, 1, Oout, 4, n_layers, 0,
                                                1 if act else 0, _stream())
            _lib.check(st, "qw_conv1d_backward_act")
            return gx, grads

        dpctx = ctx.dpctx
        if lib.qw_get_option(b"STEM_CHAIN") == 0 and dpctx is None:
            g1, grads2 = layer_backward(gy2.contiguous(), y1, ps2, p2, H, L, 2, O, True, act2)
            gx, grads1 = layer_backward(g1, x, ps1, p1, C, L, 1, H, ctx.needs_input_grad[0], act1)
            return (gx, None, None, None, *grads1, *grads2)
        # chained: conv2's backward writes NO gradient for its input; conv1's gy kernel rebuilds that gradient tile by tile from conv2's
        # gpre rows (inside conv2's workspace) and pre_conv weights (qw_conv1d_backward_chained)
        w_pre2, b_pre2, qw2, w_post2, b_post2 = p2
        grads2 = [torch.empty_like(t) for t in p2]
        n2 = lib.qw_conv1d_workspace_bytes(B, H, L, 3, 2, 1, O, 4, n_layers, 4)
        ws2 = torch.empty(n2, device=dev, dtype=torch.uint8)
        w_pre1, b_pre1, qw1, w_post1, b_post1 = p1
        grads1 = [torch.empty_like(t) for t in p1]
        n1 = lib.qw_conv1d_workspace_bytes(B, C, L, 3, 1, 1, H, 4, n_layers, 4)
        ws1 = torch.empty(n1, device=dev, dtype=torch.uint8)
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        a = 1 if act1 else 0
        with torch.cuda.device(dev):
            if dpctx is not None:  # (act2 is False here: the data-parallel backward has no activation form)
                st = lib.qw_conv1d_backward_dp(_ptr(gy2.contiguous()), _ptr(y1), _ptr(ps2), _ptr(w_pre2), _ptr(qw2), _ptr(w_post2), None,
                                               *[_ptr(g) for g in grads2], _ptr(ws2), n2, B, H, L, 3, 2, 1, O, 4, n_layers, 0,
                                               *dpctx[1].args(), _stream())
                _lib.check(st, "qw_conv1d_backward_dp")
            else:
                st = lib.qw_conv1d_backward_act(_ptr(gy2.contiguous()), _ptr(y1), _ptr(ps2), _ptr(w_pre2), _ptr(qw2), _ptr(w_post2),
                                                _ptr(b_post2), None, *[_ptr(g) for g in grads2], _ptr(ws2), n2, B, H, L, 3, 2, 1, O, 4,
                                                n_layers, 0, 1 if act2 else 0, _stream())
                _lib.check(st, "qw_conv1d_backward_act")
            dp1 = dpctx[0].args() if dpctx is not None else (None, None, 0, 1, 1.0)
            st = lib.qw_conv1d_backward_chained(_ptr(ws2), _ptr(w_pre2), O, _ptr(x), _ptr(ps1), _ptr(w_pre1), _ptr(qw1), _ptr(w_post1),
                                                _ptr(b_post1), _ptr(gx), *[_ptr(g) for g in grads1], _ptr(ws1), n1, B, C, L, 3, 1, 1, H, 4,
                                                n_layers, 0, a, *dp1, _stream())
            _lib.check(st, "qw_conv1d_backward_chained")
        return (gx, None, None, None, *grads1, *grads2)


def stem_train_eligible(conv1: "QuantumConv1d", conv2: "QuantumConv1d", x: torch.Tensor, allow_dp: bool = False) -> bool:
    """Shapes qw_stem_train_forward covers: the Whisper stem (conv1 K=3,S=1,P=1; conv2 K=3,S=2,P=1; n_qubits 4, amplitude).
    ``allow_dp``: layers with ``fuse_grad_allreduce()`` on BOTH of them qualify too (the chained backward then averages the
    gradients over the ranks inside its finalize kernels)."""
    def ok(m, S):
        return (isinstance(m, QuantumConv1d) and m.kernel_size == 3 and m.stride == S and m.padding == 1 and m.n_qubits == 4
                and m.embedding == "amplitude" and m.n_layers <= 4)
    if not (ok(conv1, 1) and ok(conv2, 2) and conv1.n_layers == conv2.n_layers and conv2.in_channels == conv1.out_channels):
        return False
    has_dp = (conv1._grad_allreduce is not None, conv2._grad_allreduce is not None)
    if any(has_dp) and not (allow_dp and all(has_dp)):
        return False
    lib = _lib.load()
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 3 and x.shape[1] == conv1.in_channels and conv1.in_channels <= 96
            and conv1.in_channels % 4 == 0 and x.shape[2] % 8 == 0 and conv1.out_channels % 4 == 0 and conv2.out_channels % 8 == 0 and conv1.out_channels <= 384
            and conv2.out_channels <= 384 and lib.qw_get_option(b"FAST_PATH") != 0 and lib.qw_get_option(b"GY_MMA") != 0)


def stem_train_forward(conv1: "QuantumConv1d", conv2: "QuantumConv1d", x: torch.Tensor, gelu=True) -> torch.Tensor:
    """``gelu(conv2(gelu(conv1(x))))`` (or without the activations) for training, differentiable in x and all ten parameters:
    the first two lines of AudioEncoder.forward (whisper/whisper/model.py:193-194) as one forward kernel (when that is the faster
    form for the batch) and a backward in which the gradient between the two layers never touches HBM.  ``gelu`` may be a pair
    ``(after_conv1, after_conv2)``, e.g. ``(True, False)`` for conv1 -> GELU -> conv2."""
    g1, g2 = (bool(gelu[0]), bool(gelu[1])) if isinstance(gelu, (tuple, list)) else (bool(gelu), bool(gelu))
    if not stem_train_eligible(conv1, conv2, x, allow_dp=not g2):
        h = conv1.forward_gelu(x) if g1 else conv1(x)
        return conv2.forward_gelu(h) if g2 else conv2(h)
    prm = [conv1.pre_conv.weight, conv1.pre_conv.bias, conv1.quantum_weights, conv1.post_conv.weight, conv1.post_conv.bias,
           conv2.pre_conv.weight, conv2.pre_conv.bias, conv2.quantum_weights, conv2.post_conv.weight, conv2.post_conv.bias]
    dpctx = (conv1._grad_allreduce, conv2._grad_allreduce) if conv1._grad_allreduce is not None else None
    return _StemTrainFn.apply(x, (g1, g2), conv1.n_layers, dpctx, *prm)


def fused_stem_eligible(conv1: "QuantumConv1d", conv2: "QuantumConv1d", x: torch.Tensor) -> bool:
    """True when ``gelu(conv2(gelu(conv1(x))))`` can run through ``qw_stem_forward`` (inference, fast-path regime)."""
    def ok(m, K, S, P):
        return (isinstance(m, QuantumConv1d) and m.kernel_size == K and m.stride == S and m.padding == P and m.n_qubits == 4
                and m.embedding == "amplitude" and m.n_layers <= 4)
    return (ok(conv1, 3, 1, 1) and ok(conv2, 3, 2, 1) and conv1.n_layers == conv2.n_layers and conv2.in_channels == conv1.out_channels
            and x.is_cuda and x.dtype == torch.float32 and x.dim() == 3 and x.shape[1] == conv1.in_channels and x.shape[2] % 4 == 0
            and conv1.out_channels % 4 == 0 and conv2.out_channels % 4 == 0 and conv1.out_channels <= 576
            and conv2.out_channels <= 576 and conv1.in_channels * 3 * 16 <= 96 * 1024)


@torch.no_grad()
def fused_stem_forward(conv1: "QuantumConv1d", conv2: "QuantumConv1d", x: torch.Tensor,
                       positional_embedding: torch.Tensor | None = None) -> torch.Tensor:
    """Inference forward of the encoder stem in two kernels:

        gelu(conv2(gelu(conv1(x)))).permute(0, 2, 1) + positional_embedding      # whisper/whisper/model.py:193-198

    x (B, n_mels, L) -> (B, L // 2, n_state).  The (B, n_state, L) activation between the layers is never written.
    No autograd graph is recorded: use the two modules for training."""
    if not fused_stem_eligible(conv1, conv2, x):
        raise ValueError("fused_stem_forward needs the Whisper stem regime (n_qubits=4, k3/s1/p1 + k3/s2/p1, fp32 CUDA input, "
                         "L % 4 == 0, channel counts % 4 == 0 and <= 576)")
    lib = _lib.load()
    B, C, L = x.shape
    hidden, O = conv1.out_channels, conv2.out_channels
    x = x.contiguous()
    p1 = [t.detach().contiguous() for t in (conv1.pre_conv.weight, conv1.pre_conv.bias, conv1.quantum_weights,
                                            conv1.post_conv.weight, conv1.post_conv.bias)]
    p2 = [t.detach().contiguous() for t in (conv2.pre_conv.weight, conv2.pre_conv.bias, conv2.quantum_weights,
                                            conv2.post_conv.weight, conv2.post_conv.bias)]
    for t in p1 + p2:
        if t.device != x.device or t.dtype != torch.float32:
            raise RuntimeError("stem parameters must be float32 on the input's device")
    pos = None
    if positional_embedding is not None:
        if tuple(positional_embedding.shape) != (L // 2, O):
            raise AssertionError("incorrect audio shape")  # whisper/model.py:197
        pos = positional_embedding.to(device=x.device, dtype=torch.float32).contiguous()
    out = torch.empty(B, L // 2, O, device=x.device, dtype=torch.float32)
    nbytes = lib.qw_stem_workspace_bytes(B, L)
    ws = torch.empty(nbytes, device=x.device, dtype=torch.uint8)
    with torch.cuda.device(x.device):
        st = lib.qw_stem_forward(_ptr(x), *[_ptr(t) for t in p1], *[_ptr(t) for t in p2], _ptr(pos), _ptr(out), _ptr(ws), nbytes,
                                 B, C, L, hidden, O, conv1.n_layers, _stream())
    _lib.check(st, "qw_stem_forward")
    return out


class _CircuitFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pre, qw, n_layers, emb):
        lib = _lib.load()
        W, q = pre.shape
        f64 = pre.dtype == torch.float64
        pre = pre.contiguous()
        qw = qw.contiguous()
        out = torch.empty_like(pre)
        fn = lib.qw_circuit_forward_f64 if f64 else lib.qw_circuit_forward
        with torch.cuda.device(pre.device):
            st = fn(_ptr(pre), _ptr(qw), _ptr(out), W, q, n_layers, emb, _stream())
        _lib.check(st, "qw_circuit_forward")
        ctx.save_for_backward(pre, qw)
        ctx.cfg = (W, q, n_layers, emb, f64)
        return out

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.load()
        pre, qw = ctx.saved_tensors
        W, q, n_layers, emb, f64 = ctx.cfg
        gout = gout.contiguous()
        gpre = torch.empty_like(pre)
        gqw = torch.empty_like(qw)
        nbytes = lib.qw_circuit_workspace_bytes(W, q, n_layers, 8 if f64 else 4)
        ws = torch.empty(max(nbytes, 16), device=pre.device, dtype=torch.uint8)
        fn = lib.qw_circuit_backward_f64 if f64 else lib.qw_circuit_backward
        with torch.cuda.device(pre.device):
            st = fn(_ptr(pre), _ptr(qw), _ptr(gout), _ptr(gpre), _ptr(gqw), _ptr(ws), nbytes, W, q, n_layers, emb,
                    _stream())
        _lib.check(st, "qw_circuit_backward")
        return gpre, gqw, None, None


class _CollapsedFn(torch.autograd.Function):
    """out_i = xh^T M_i xh per window (qw_circuit_forward_collapsed); backward returns gpre and gM."""

    @staticmethod
    def forward(ctx, pre, M):
        lib = _lib.load()
        W, q = pre.shape
        f64 = pre.dtype == torch.float64
        pre = pre.contiguous()
        M = M.contiguous()
        out = torch.empty_like(pre)
        fn = lib.qw_circuit_forward_collapsed_f64 if f64 else lib.qw_circuit_forward_collapsed
        with torch.cuda.device(pre.device):
            st = fn(_ptr(pre), _ptr(M), _ptr(out), W, q, _stream())
        _lib.check(st, "qw_circuit_forward_collapsed")
        ctx.save_for_backward(pre, M)
        ctx.cfg = (W, q, f64)
        return out

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.load()
        pre, M = ctx.saved_tensors
        W, q, f64 = ctx.cfg
        gout = gout.contiguous()
        gpre = torch.empty_like(pre)
        gM = torch.empty_like(M)
        nbytes = lib.qw_circuit_collapsed_workspace_bytes(W, q, 8 if f64 else 4)
        ws = torch.empty(max(nbytes, 16), device=pre.device, dtype=torch.uint8)
        fn = lib.qw_circuit_backward_collapsed_f64 if f64 else lib.qw_circuit_backward_collapsed
        with torch.cuda.device(pre.device):
            st = fn(_ptr(pre), _ptr(M), _ptr(gout), _ptr(gpre), _ptr(gM), _ptr(ws), nbytes, W, q, _stream())
        _lib.check(st, "qw_circuit_backward_collapsed")
        return gpre, gM


def collapsed_matrices(quantum_weights: torch.Tensor, n_qubits: int, n_layers: int = 1) -> torch.Tensor:
    """M (q, q, q), M[i] = Re(U[:, :q]^H Z'_i U[:, :q]) of SURVEY.md 8a (iii), read off the STATEVECTOR kernel: it is evaluated on the
    q (q + 1) / 2 probe windows e_a and (e_a + e_b) / sqrt 2 and M_aa = f(e_a), M_ab = f((e_a + e_b)/sqrt 2) - (M_aa + M_bb) / 2.
    Differentiable: gradients wrt the weights flow back through the statevector adjoint kernel on those probes."""
    q = n_qubits
    dev, dt = quantum_weights.device, quantum_weights.dtype
    ia, ib = torch.triu_indices(q, q, device=dev)
    probes = torch.zeros(ia.numel(), q, device=dev, dtype=dt)
    rows = torch.arange(ia.numel(), device=dev)
    probes[rows, ia] = 1.0
    probes[rows, ib] = 1.0  # diagonal probes: e_a (normalised by the embedding); off-diagonal: e_a + e_b -> (e_a + e_b)/sqrt 2
    f = _CircuitFn.apply(probes, quantum_weights, int(n_layers), EMBEDDINGS["amplitude"])  # (n_probes, q): f[p, i]
    diag = f[ia == ib]                                    # (q, q): diag[a, i] = M_i[a][a]
    off = f - 0.5 * (diag[ia] + diag[ib])                 # for a == b this is 0 and is replaced below
    vals = torch.where((ia == ib)[:, None], f, off)       # (n_probes, q)
    M = torch.zeros(q, q, q, device=dev, dtype=dt)
    M = M.index_put((torch.arange(q, device=dev)[None, :].expand(ia.numel(), q), ia[:, None].expand(-1, q), ib[:, None].expand(-1, q)), vals)
    M = M + M.transpose(1, 2) - torch.diag_embed(torch.diagonal(M, dim1=1, dim2=2))
    return M


def quantum_circuit(pre: torch.Tensor, quantum_weights: torch.Tensor, n_layers: int = 1,
                    embedding: str = "amplitude", simulator: str = "statevector") -> torch.Tensor:
    """The QNode alone, batched: pre (W, q) -> <Z_i> (W, q)  (quantum_whisper.py:64-85).

    ``simulator="statevector"`` (default) is the batched statevector simulator + adjoint differentiation.
    ``simulator="collapsed"`` (opt-in, amplitude embedding only) evaluates the same circuit through the quadratic forms
    ``xh^T M_i xh`` (SURVEY.md 8a iii) with M read off the statevector kernel on q(q+1)/2 probe windows: a separately reported
    mode and the device-side second oracle of the tests -- never the default."""
    if not pre.is_cuda:
        raise RuntimeError("quantum_circuit (B200 build) has no CPU path")
    if pre.dim() != 2:
        raise ValueError("pre must be (windows, n_qubits)")
    if simulator == "statevector":
        return _CircuitFn.apply(pre, quantum_weights, int(n_layers), EMBEDDINGS[embedding])
    if simulator != "collapsed":
        raise ValueError(f"simulator must be 'statevector' or 'collapsed', got {simulator!r}")
    if embedding != "amplitude":
        raise ValueError("the collapsed quadratic-form evaluation exists for amplitude embedding only")
    M = collapsed_matrices(quantum_weights, pre.shape[1], n_layers)
    return _CollapsedFn.apply(pre, M)
