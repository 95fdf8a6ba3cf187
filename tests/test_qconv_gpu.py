"""Parity of the CUDA QuantumConv1d / circuit kernels (through the C ABI) against the fp64 oracle.

Tolerances (BASELINE.json north_star): fp32 <= 1e-5 abs on <Z_i> and per-window gradients, fp64 validation
build <= 1e-10, window indexing bit-exact.  For quantities that are sums over many windows (layer outputs
after post_conv are not, weight gradients are) the bound is relative to the largest reference entry and is
written next to each assert (SURVEY.md 8c "Tolerances").
"""
import ctypes
import json
import os

import numpy as np
import pytest
import torch

from oracle import qconv_oracle as qo

pytestmark = pytest.mark.gpu


def _mods():
    import importlib

    from qasr_ijcnlp_b200 import _lib
    return _lib, importlib.import_module("qasr_ijcnlp_b200.quantum_conv1d")


def _kat(golden_dir):
    with open(os.path.join(golden_dir, "qconv_kat.json")) as fh:
        return json.load(fh)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.float64, 1e-10)])
def test_circuit_known_answers(cuda, golden_dir, dtype, tol):
    _, qc = _mods()
    kat = _kat(golden_dir)
    pre = torch.tensor([kat["pre"]], dtype=dtype, device=cuda, requires_grad=True)
    for name in ("kat0", "kat1"):
        k = kat[name]
        w = torch.tensor(k["quantum_weights"], dtype=dtype, device=cuda, requires_grad=True)
        out = qc.quantum_circuit(pre, w)
        assert np.abs(out.detach().cpu().numpy()[0] - np.array(k["out"])).max() <= tol
        if "cotangent" in k:
            g = torch.tensor([k["cotangent"]], dtype=dtype, device=cuda)
            gp, gw = torch.autograd.grad(out, [pre, w], g)
            assert np.abs(gp.cpu().numpy()[0] - np.array(k["grad_pre"])).max() <= tol
            assert np.abs(gw.cpu().numpy() - np.array(k["grad_quantum_weights"])).max() <= tol


@pytest.mark.parametrize("q", [1, 2, 3, 4])
@pytest.mark.parametrize("n_layers", [1, 2, 3])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.float64, 1e-10)])
def test_circuit_random_vs_oracle(cuda, q, n_layers, dtype, tol):
    _, qc = _mods()
    g = torch.Generator().manual_seed(100 * q + n_layers)
    W = 3001  # ragged vs the 128-thread blocks
    pre64 = torch.randn(W, q, generator=g, dtype=torch.float64)
    shape = (q, 3) if n_layers == 1 else (n_layers, q, 3)
    w64 = torch.randn(*shape, generator=g, dtype=torch.float64)
    cot64 = torch.randn(W, q, generator=g, dtype=torch.float64)
    # the oracle sees exactly the values the kernel sees (rounded to the kernel dtype)
    pre_o = pre64.to(dtype).double().requires_grad_(True)
    w_o = w64.to(dtype).double().requires_grad_(True)
    out_o = qo.circuit_expvals(pre_o, w_o)
    gp_o, gw_o = torch.autograd.grad(out_o, [pre_o, w_o], cot64.to(dtype).double())
    pre = pre64.to(dtype).to(cuda).requires_grad_(True)
    w = w64.to(dtype).to(cuda).requires_grad_(True)
    out = qc.quantum_circuit(pre, w, n_layers=n_layers)
    gp, gw = torch.autograd.grad(out, [pre, w], cot64.to(dtype).to(cuda))
    assert (out.detach().cpu().double() - out_o.detach()).abs().max().item() <= tol
    # per-window gradients: |d out / d pre| scales with 1/||pre||; bound relative to that scale
    scale = (1.0 / pre_o.detach().norm(dim=1, keepdim=True)).clamp(min=1.0)
    assert ((gp.cpu().double() - gp_o) / scale).abs().max().item() <= tol
    # weight gradient is a sum over W windows: bound relative to its largest entry
    assert (gw.cpu().double() - gw_o).abs().max().item() <= tol * max(1.0, gw_o.abs().max().item()) * 4


def test_circuit_identities(cuda):
    """SURVEY.md 8c identities: omega-invariance, scale invariance, product-state channels."""
    _, qc = _mods()
    g = torch.Generator().manual_seed(5)
    pre = torch.randn(513, 4, generator=g).to(cuda)
    w = torch.randn(4, 3, generator=g).to(cuda)
    out = qc.quantum_circuit(pre, w)
    w2 = w.clone()
    w2[:, 2] += torch.randn(4, generator=g).to(cuda)  # omega on every wire
    w2[0, 0] += 0.7  # phi on wires < q - m
    w2[1, 0] -= 1.3
    assert (qc.quantum_circuit(pre, w2) - out).abs().max().item() <= 1e-5
    assert (qc.quantum_circuit(-3.7 * pre, w) - out).abs().max().item() <= 1e-5
    c0 = torch.cos(w[0, 1])
    c1 = torch.cos(w[1, 1])
    assert (out[:, 0] - c0).abs().max().item() <= 1e-5
    assert (out[:, 1] - c0 * c1).abs().max().item() <= 1e-5


GEOMS = [
    # (B, C, L, K, S, P, O, q) -- Whisper-Tiny stem geometries at reduced length + ragged/edge cases
    (2, 80, 200, 3, 1, 1, 384, 4),
    (2, 384, 203, 3, 2, 1, 384, 4),
    (3, 5, 77, 3, 1, 1, 9, 4),
    (1, 7, 64, 3, 2, 0, 33, 3),
    (2, 4, 50, 5, 3, 2, 6, 2),
    (2, 3, 31, 1, 1, 0, 5, 1),
    (1, 33, 130, 4, 2, 3, 65, 4),
]


def _run_layer(cuda, geom, dtype, n_layers=1, seed=0):
    _, qc = _mods()
    B, C, L, K, S, P, O, q = geom
    params64 = qo.make_params(C, O, K, q, n_layers=n_layers, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    x64 = torch.randn(B, C, L, generator=g, dtype=torch.float64)
    Lo = qo.out_length(L, K, S, P)
    gy64 = torch.randn(B, O, Lo, generator=g, dtype=torch.float64)
    # oracle on the dtype-rounded values
    xr = x64.to(dtype).double()
    pr = [p.to(dtype).double() for p in params64]
    ref = qo.qconv1d_grads(xr, pr, gy64.to(dtype).double(), K, S, P)
    x = x64.to(dtype).to(cuda).requires_grad_(True)
    ps = [p.to(dtype).to(cuda).requires_grad_(True) for p in params64]
    y = qc.quantum_conv1d(x, *ps, kernel_size=K, stride=S, padding=P, n_layers=n_layers)
    grads = torch.autograd.grad(y, [x] + ps, gy64.to(dtype).to(cuda))
    got = dict(zip(["x", "w_pre", "b_pre", "qweights", "w_post", "b_post"], [t.cpu().double() for t in grads]))
    got["y"] = y.detach().cpu().double()
    return got, ref


def _rel(a, b):
    return (a - b).abs().max().item() / max(1.0, b.abs().max().item())


@pytest.mark.parametrize("geom", GEOMS)
def test_layer_f64_validation_build(cuda, geom):
    got, ref = _run_layer(cuda, geom, torch.float64)
    for k in ("y", "x", "w_pre", "b_pre", "qweights", "w_post", "b_post"):
        assert _rel(got[k], ref[k]) <= 1e-10, k  # fp64 validation build: 1e-10 (relative to max(1,|ref|))


FAST_GEOMS = [
    # shapes that qualify for the TMA fast path (fp32, q=4, K=3, S in {1,2}, L%4==0, L_out%4==0, O%4==0)
    (2, 80, 200, 3, 1, 1, 384, 4),
    (2, 384, 208, 3, 2, 1, 384, 4),
    (3, 80, 260, 3, 1, 1, 384, 4),
    (2, 40, 136, 3, 2, 1, 64, 4),
    (1, 100, 1028, 3, 1, 1, 8, 4),
    (5, 33, 72, 3, 2, 1, 200, 4),
    (2, 130, 392, 3, 2, 3, 512, 4),
]


@pytest.fixture(params=[True, False], ids=["fast", "generic"])
def fast_path(request):
    from qasr_ijcnlp_b200 import _lib
    lib = _lib.load()
    lib.qw_set_fast_path(1 if request.param else 0)
    yield request.param
    lib.qw_set_fast_path(1)


@pytest.mark.parametrize("geom", FAST_GEOMS)
@pytest.mark.parametrize("n_layers", [1, 2])
def test_layer_f32_fast_and_generic(cuda, geom, n_layers, fast_path):
    """Same tolerances as test_layer_f32, run once through the TMA fast path and once through the generic kernels."""
    got, ref = _run_layer(cuda, geom, torch.float32, n_layers=n_layers, seed=5)
    assert (got["y"] - ref["y"]).abs().max().item() <= 5e-5
    for k in ("x", "w_pre", "b_pre", "qweights", "w_post", "b_post"):
        assert _rel(got[k], ref[k]) <= 5e-5, k


def test_fast_and_generic_agree_full_size(cuda):
    """Whisper-Tiny stem shapes at batch 4: fast path vs generic kernels (both fp32) on identical inputs."""
    _lib, qc = _mods()
    lib = _lib.load()
    for (C, S) in ((80, 1), (384, 2)):
        torch.manual_seed(21)
        m = qc.QuantumConv1d(C, 384, 3, stride=S, padding=1, n_qubits=4).to(cuda)
        x = torch.randn(4, C, 3000, device=cuda, requires_grad=True)
        res = []
        for fast in (1, 0):
            lib.qw_set_fast_path(fast)
            y = m(x)
            gy = torch.ones_like(y) * torch.linspace(-1, 1, y.shape[-1], device=cuda)
            grads = torch.autograd.grad(y, [x] + list(m.parameters()), gy)
            res.append([y.detach()] + [g for g in grads])
        lib.qw_set_fast_path(1)
        # y and grad_x are per-window quantities; the five parameter gradients are fp32 sums over N = 12 000 / 6 000 windows
        # (this cotangent makes grad post_conv.bias an almost exactly cancelling sum of terms of magnitude <= 1), so they get
        # the random-walk rounding allowance 2e-6 * sqrt(N) on top -- the two paths only differ in summation order.
        N = y.shape[0] * y.shape[2]
        for i, (a, b) in enumerate(zip(*res)):
            tol = 2e-5 * max(1.0, b.abs().max().item()) + (2e-6 * N ** 0.5 if i >= 2 else 0.0)
            assert (a - b).abs().max().item() <= tol, i


def _set_opt(lib, name, value):
    assert lib.qw_set_option(name.encode(), value) == 0, name


@pytest.mark.parametrize("B,C,L,O", [(16, 80, 3000, 384), (2, 80, 3000, 384), (3, 80, 260, 384), (1, 40, 1028, 8), (5, 96, 72, 200),
                                     (7, 33, 64, 64)])
def test_fused_backward_data_layer(cuda, B, C, L, O):
    """A data layer (its input needs no gradient: conv1 of the stem) at small batch can run its whole backward in ONE kernel
    (gy pass on the tensor pipe + adjoint warp + pre_conv^T, `fast_bwd_gy3_kernel<.., 1>`, option BWD_FUSED=1; measured slower
    than the three kernels on B200 and therefore off by default).  All five parameter gradients against the
    fp64 oracle (same bound as test_layer_f32: 5e-5 relative to max(1, |ref|)) and against the three-kernel path."""
    _lib, qc = _mods()
    lib = _lib.load()
    params64 = qo.make_params(C, O, 3, 4, seed=9)
    g = torch.Generator().manual_seed(10)
    x64 = torch.randn(B, C, L, generator=g, dtype=torch.float64)
    gy64 = torch.randn(B, O, L, generator=g, dtype=torch.float64)
    xr = x64.float().double()
    pr = [p.float().double() for p in params64]
    ref = qo.qconv1d_grads(xr, pr, gy64.float().double(), 3, 1, 1, need_gx=False)
    res = {}
    launches = {}
    try:
        for fused in (1, 0):
            _set_opt(lib, "BWD_FUSED", fused)
            x = x64.float().to(cuda)  # no requires_grad: grad_x is not computed
            ps = [p.float().to(cuda).requires_grad_(True) for p in params64]
            y = qc.quantum_conv1d(x, *ps, kernel_size=3, stride=1, padding=1)
            n0 = _lib.launch_count()
            grads = torch.autograd.grad(y, ps, gy64.float().to(cuda))
            torch.cuda.synchronize()
            launches[fused] = _lib.launch_count() - n0
            res[fused] = dict(zip(["w_pre", "b_pre", "qweights", "w_post", "b_post"], [t.cpu().double() for t in grads]))
    finally:
        _set_opt(lib, "BWD_FUSED", 0)
    assert launches == {1: 2, 0: 4}, launches  # fused + finalize vs gy / adjoint / pre_conv^T / finalize
    for k in ("w_pre", "b_pre", "qweights", "w_post", "b_post"):
        assert _rel(res[1][k], ref[k]) <= 5e-5, k
        assert _rel(res[0][k], ref[k]) <= 5e-5, k
        assert _rel(res[1][k], res[0][k]) <= 2e-5, k


def _abi_forward_backward(lib, _lib, cuda, x, params, gy, K, S, P, need_gx=True):
    """One forward + backward straight through the C ABI; returns y, pre_save (2, W, q) and the gradients."""
    B, C, L = x.shape
    O, q = params[3].shape
    Lo = qo.out_length(L, K, S, P)
    dev = [t.to(cuda).contiguous() for t in (x, *params)]
    y = torch.empty(B, O, Lo, device=cuda)
    pre_save = torch.empty(2, B * Lo, q, device=cuda)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    P_ = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    _lib.check(lib.qw_conv1d_forward(*[P_(t) for t in dev], P_(y), P_(pre_save), B, C, L, K, S, P, O, q, 1, 0, st), "qw_conv1d_forward")
    gyd = gy.to(cuda).contiguous()
    gx = torch.empty_like(dev[0]) if need_gx else None
    grads = [torch.empty_like(t) for t in dev[1:]]
    n = lib.qw_conv1d_workspace_bytes(B, C, L, K, S, P, O, q, 1, 4)
    ws = torch.empty(n, device=cuda, dtype=torch.uint8)
    _lib.check(lib.qw_conv1d_backward(P_(gyd), P_(dev[0]), P_(pre_save), P_(dev[1]), P_(dev[3]), P_(dev[4]), P_(gx), *[P_(t) for t in grads],
                                      P_(ws), n, B, C, L, K, S, P, O, q, 1, 0, st), "qw_conv1d_backward")
    torch.cuda.synchronize()
    out = dict(zip(["w_pre", "b_pre", "qweights", "w_post", "b_post"], [t.cpu().double() for t in grads]))
    if need_gx:
        out["x"] = gx.cpu().double()
    return y.cpu().double(), pre_save.cpu().double(), out


@pytest.mark.parametrize("layer", ["conv1", "conv2"])
def test_stem_layers_batch16_full_parity(cuda, layer):
    """BASELINE.json configs[1]/[2] stem shapes at the bench batch (16): every output of the fast path against the fp64 oracle at
    FULL size -- y, the saved pre_conv outputs (plane 0) and per-window <Z_i> (plane 1, <= 1e-5 abs: the north_star bound on
    expectation values), and all gradients (5e-5 relative to max(1, |ref|), summed quantities) -- for the default backward
    (tensor-pipe gy pass, adjoint kernel, pre_conv^T kernel) and for the two forms with the adjoint inside the gy kernel
    (BWD_FUSED = 1 / 2: measured slower at batch 16, kept as A/B switches)."""
    _lib, _ = _mods()
    lib = _lib.load()
    C, S, need_gx = (80, 1, False) if layer == "conv1" else (384, 2, True)
    B, L, O = 16, 3000, 384
    params64 = qo.make_params(C, O, 3, 4, seed=31)
    g = torch.Generator().manual_seed(32)
    x = torch.randn(B, C, L, generator=g)
    Lo = qo.out_length(L, 3, S, 1)
    gy = torch.randn(B, O, Lo, generator=g)
    params = [p.float() for p in params64]
    p64 = [p.double() for p in params]
    yo, preo, qouto = qo.qconv1d_forward(x.double(), *p64, K=3, S=S, P=1, return_intermediates=True)
    ref = qo.qconv1d_grads(x.double(), p64, gy.double(), 3, S, 1, need_gx=need_gx)
    try:
        for fused in (0, 1, 2):
            _set_opt(lib, "BWD_FUSED", fused)
            y, pre_save, got = _abi_forward_backward(lib, _lib, cuda, x, params, gy, 3, S, 1, need_gx=need_gx)
            assert (y - yo).abs().max().item() <= 5e-5
            # plane 0: pre_conv outputs (fp32 sums of C*K products of O(1) terms): 2e-6 * sqrt(C*K) abs
            assert (pre_save[0] - preo.reshape(-1, 4)).abs().max().item() <= 2e-6 * (3 * C) ** 0.5
            # plane 1: <Z_i> per window, evaluated by the kernel on ITS fp32 pre_conv output -> compare with the oracle circuit on
            # exactly those values: the north_star's 1e-5 bound on expectation values
            z_ref = qo.circuit_expvals(pre_save[0].contiguous(), p64[2])
            assert (pre_save[1] - z_ref).abs().max().item() <= 1e-5
            assert (pre_save[1] - qouto.reshape(-1, 4)).abs().max().item() <= 5e-5  # and end to end from x
            for k, v in got.items():
                assert _rel(v, ref[k]) <= 5e-5, (k, fused)
    finally:
        _set_opt(lib, "BWD_FUSED", 0)


@pytest.mark.parametrize("C,S,B,L", [(80, 1, 2, 3000), (384, 2, 2, 3000), (384, 2, 16, 3000), (40, 2, 3, 136), (96, 1, 1, 260)])
def test_gelu_fused_layer(cuda, C, S, B, L):
    """SURVEY.md 8-f1 at training time: `QuantumConv1d.forward_gelu` == F.gelu(layer(x)) (whisper/whisper/model.py:193-194) with
    the GELU inside the forward epilogue and gelu' inside the backward's gy pass.  Output and all six gradients against the fp64
    oracle composed with torch's exact GELU (5e-5 abs on y, 5e-5 relative to max(1, |ref|) on the gradients), and against the
    unfused device path (plain operator + ATen GELU); the fused path must take exactly the layer's own launches (no GELU kernel)."""
    _lib, qc = _mods()
    torch.manual_seed(41)
    m = qc.QuantumConv1d(C, 384, 3, stride=S, padding=1, n_qubits=4).to(cuda)
    g = torch.Generator().manual_seed(42)
    x = torch.randn(B, C, L, generator=g) * 1.5
    Lo = qo.out_length(L, 3, S, 1)
    gy = torch.randn(B, 384, Lo, generator=g)
    params = [m.pre_conv.weight, m.pre_conv.bias, m.quantum_weights, m.post_conv.weight, m.post_conv.bias]
    # oracle: fp64, exact erf GELU
    leaves = [p.detach().cpu().double().requires_grad_(True) for p in params]
    xo = x.double().requires_grad_(True)
    yo = torch.nn.functional.gelu(qo.qconv1d_forward(xo, *leaves, K=3, S=S, P=1))
    ref = torch.autograd.grad(yo, [xo] + leaves, gy.double())
    xc = x.to(cuda).requires_grad_(True)
    n0 = _lib.launch_count()
    y = m.forward_gelu(xc)
    n_fwd = _lib.launch_count() - n0
    got = torch.autograd.grad(y, [xc] + params, gy.to(cuda))
    torch.cuda.synchronize()
    assert n_fwd == 1 and _lib.launch_count() - n0 == 5  # forward kernel + gy / adjoint / pre_conv^T / finalize: the whole step
    assert (y.detach().cpu().double() - yo.detach()).abs().max().item() <= 5e-5
    for name, a, b in zip(["x", "w_pre", "b_pre", "qweights", "w_post", "b_post"], got, ref):
        assert _rel(a.cpu().double(), b) <= 5e-5, name
    # unfused device path
    y2 = torch.nn.functional.gelu(m(xc))
    got2 = torch.autograd.grad(y2, [xc] + params, gy.to(cuda))
    assert (y - y2).abs().max().item() <= 2e-6
    for name, a, b in zip(["x", "w_pre", "b_pre", "qweights", "w_post", "b_post"], got, got2):
        assert _rel(a.double(), b.double()) <= 2e-5, name


def test_gelu_fused_layer_falls_back_outside_the_fast_path(cuda):
    _lib, qc = _mods()
    torch.manual_seed(1)
    m = qc.QuantumConv1d(7, 33, 3, stride=2, padding=0, n_qubits=3).to(cuda)  # generic kernels: no fused activation
    x = torch.randn(2, 7, 65, device=cuda)
    assert torch.equal(m.forward_gelu(x), torch.nn.functional.gelu(m(x)))
    lib = _lib.load()
    st = lib.qw_conv1d_forward_act(None, None, None, None, None, None, None, None, 1, 1, 1, 1, 1, 0, 1, 1, 1, 0, 1, None)
    assert st == -1  # null pointers
    y = torch.empty(2, 33, 32, device=cuda)
    p = [m.pre_conv.weight, m.pre_conv.bias, m.quantum_weights, m.post_conv.weight, m.post_conv.bias]
    P_ = lambda t: ctypes.c_void_p(t.data_ptr())
    st = lib.qw_conv1d_forward_act(P_(x), *[P_(t) for t in p], P_(y), None, 2, 7, 65, 3, 2, 0, 33, 3, 1, 0, 1,
                                   ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert st == -2 and b"fast-path" in lib.qw_last_error()


@pytest.mark.parametrize("mma", [1, 0])
def test_gy_pass_tensor_pipe_vs_ffma(cuda, mma):
    """The gy pass of the backward exists in a tensor-pipe form (mma.sync m16n8k8, 3 x TF32 split, default) and an FFMA form:
    both within the layer tolerance of the oracle at a stem-shaped geometry with ragged last tiles (L_out = 1500 = 46 x 32 + 28)."""
    _lib, _ = _mods()
    lib = _lib.load()
    try:
        _set_opt(lib, "GY_MMA", mma)
        got, ref = _run_layer(cuda, (2, 384, 3000, 3, 2, 1, 384, 4), torch.float32, seed=13)
    finally:
        _set_opt(lib, "GY_MMA", 1)
    assert (got["y"] - ref["y"]).abs().max().item() <= 5e-5
    for k in ("x", "w_pre", "b_pre", "qweights", "w_post", "b_post"):
        assert _rel(got[k], ref[k]) <= 5e-5, k


@pytest.mark.parametrize("geom", GEOMS)
def test_layer_f32(cuda, geom):
    got, ref = _run_layer(cuda, geom, torch.float32)
    # y = post_conv(<Z>) with |w_post| <= 1/sqrt(q): 1e-5 on <Z> -> <= ~2e-5 on y; the fp32 pre_conv
    # reduction adds O(1e-6 / ||pre||) per window, so the bound used is 5e-5 abs.
    assert (got["y"] - ref["y"]).abs().max().item() <= 5e-5
    # gradients are sums over up to B*L_out windows / O channels: bound relative to the largest entry
    for k in ("x", "w_pre", "b_pre", "qweights", "w_post", "b_post"):
        assert _rel(got[k], ref[k]) <= 5e-5, k


@pytest.mark.parametrize("n_layers", [2, 3])
def test_layer_multilayer_extension(cuda, n_layers):
    got, ref = _run_layer(cuda, (2, 80, 150, 3, 1, 1, 384, 4), torch.float64, n_layers=n_layers, seed=3)
    for k in ("y", "x", "w_pre", "b_pre", "qweights", "w_post", "b_post"):
        assert _rel(got[k], ref[k]) <= 1e-10, k


@pytest.mark.parametrize("geom", [(2, 80, 3000, 3, 1, 1, 384, 4), (2, 384, 3000, 3, 2, 1, 384, 4), (3, 6, 100, 5, 3, 2, 4, 4)])
def test_window_indexing_bit_exact(cuda, geom):
    """One-hot pre_conv rows turn pre_save into a gather of x: must equal F.unfold bit for bit
    (quantum_whisper.py:99-111; feature f = c*K + k; zero padding both sides)."""
    _lib, qc = _mods()
    lib = _lib.load()
    B, C, L, K, S, P, O, q = geom
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, C, L, generator=g)
    Lo = qo.out_length(L, K, S, P)
    ref_win = torch.nn.functional.unfold(x[:, :, None, :], (1, K), padding=(0, P), stride=(1, S))  # (B, C*K, Lo)
    feats = torch.randint(0, C * K, (8, q), generator=g).tolist() + [[0, 1, C * K - 2, C * K - 1][:q]]
    for fsel in feats:
        w_pre = torch.zeros(q, C * K)
        for j, f in enumerate(fsel):
            w_pre[j, f] = 1.0
        dev = [t.to(cuda).contiguous() for t in (x, w_pre, torch.zeros(q), torch.randn(q, 3, generator=g),
                                                 torch.randn(O, q, generator=g), torch.zeros(O))]
        y = torch.empty(B, O, Lo, device=cuda)
        pre_save = torch.empty(2, B * Lo, q, device=cuda)
        st = lib.qw_conv1d_forward(*[ctypes.c_void_p(t.data_ptr()) for t in dev], ctypes.c_void_p(y.data_ptr()),
                                   ctypes.c_void_p(pre_save.data_ptr()), B, C, L, K, S, P, O, q, 1, 0,
                                   ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(st, "qw_conv1d_forward")
        got = pre_save[0].cpu().reshape(B, Lo, q)
        want = ref_win[:, fsel, :].permute(0, 2, 1)
        assert torch.equal(got, want)


def test_config1_full_size_parity(cuda):
    """BASELINE.json configs[0]: QuantumConv1d(80,384,3,padding=1,n_qubits=4), x = randn(2,80,3000), seed 0,
    module-default init; y, per-window <Z> (on the kernel's own fp32 pre_conv output) and all grads."""
    _, qc = _mods()
    torch.manual_seed(0)
    m = qc.QuantumConv1d(80, 384, 3, padding=1, n_qubits=4)
    x = torch.randn(2, 80, 3000)
    for xin in (x, (torch.rand(2, 80, 3000) * 3 - 1.5)):
        params = [m.pre_conv.weight, m.pre_conv.bias, m.quantum_weights, m.post_conv.weight, m.post_conv.bias]
        p64 = [p.detach().double() for p in params]
        gy = torch.randn(2, 384, 3000)
        ref = qo.qconv1d_grads(xin.double(), p64, gy.double(), 3, 1, 1)
        md = m.to(cuda)
        xc = xin.to(cuda).requires_grad_(True)
        y = md(xc)
        grads = torch.autograd.grad(y, [xc] + list(md.parameters()), gy.to(cuda))
        names = ["x"] + [n for n, _ in md.named_parameters()]
        assert names == ["x", "quantum_weights", "pre_conv.weight", "pre_conv.bias", "post_conv.weight", "post_conv.bias"]
        key = {"quantum_weights": "qweights", "pre_conv.weight": "w_pre", "pre_conv.bias": "b_pre",
               "post_conv.weight": "w_post", "post_conv.bias": "b_post", "x": "x"}
        assert (y.detach().cpu().double() - ref["y"]).abs().max().item() <= 5e-5
        for n, gr in zip(names, grads):
            assert _rel(gr.cpu().double(), ref[key[n]]) <= 5e-5, n
        m = m.to("cpu")
        # per-window <Z_i> (pre_save plane 1) on the kernel's own fp32 pre_conv output (plane 0): <= 1e-5 abs
        _lib, _ = _mods()
        _, pre_save, _ = _abi_forward_backward(_lib.load(), _lib, cuda, xin, [p.detach() for p in params], gy, 3, 1, 1, need_gx=True)
        yo, preo, qouto = qo.qconv1d_forward(xin.double(), *p64, K=3, S=1, P=1, return_intermediates=True)
        assert (pre_save[0] - preo.reshape(-1, 4)).abs().max().item() <= 2e-6 * 240 ** 0.5
        assert (pre_save[1] - qo.circuit_expvals(pre_save[0].contiguous(), p64[2])).abs().max().item() <= 1e-5
        assert (pre_save[1] - qouto.reshape(-1, 4)).abs().max().item() <= 5e-5


def test_full_size_properties_conv2(cuda):
    """Whisper-Tiny conv2 at batch 16 (BASELINE configs[1]/[2] stem shape): size-independent properties.
    (i) scale invariance of the embedding: x * grad_x sums to 0 per window set -> <x, gx> == 0 when b_pre = 0;
    (ii) linearity of the backward in gy; (iii) parity of a random subset of windows with the oracle."""
    _, qc = _mods()
    torch.manual_seed(3)
    m = qc.QuantumConv1d(384, 384, 3, stride=2, padding=1, n_qubits=4).to(cuda)
    with torch.no_grad():
        m.pre_conv.bias.zero_()
    x = torch.randn(16, 384, 3000, device=cuda, requires_grad=True)
    y = m(x)
    gy = torch.randn_like(y)
    gx, gw = torch.autograd.grad(y, [x, m.pre_conv.weight], gy, retain_graph=True)
    # (i) y is invariant under x -> (1+eps) x when b_pre = 0, so <x, dL/dx> = 0 (up to fp32 rounding of a 18M-term sum)
    dot = (x.detach().double() * gx.double()).sum().item()
    scale = (x.detach().double().abs() * gx.double().abs()).sum().item()
    assert abs(dot) <= 1e-5 * scale
    # same for the pre_conv weight: <W, dL/dW> = 0
    dotw = (m.pre_conv.weight.detach().double() * gw.double()).sum().item()
    scalew = (m.pre_conv.weight.detach().double().abs() * gw.double().abs()).sum().item()
    assert abs(dotw) <= 1e-4 * scalew
    # (ii) linearity in gy
    gx2, = torch.autograd.grad(y, [x], 2.0 * gy, retain_graph=True)
    assert (gx2 - 2.0 * gx).abs().max().item() <= 1e-6 * max(1.0, gx.abs().max().item())
    # (iii) oracle on utterances 0 and 15, first 256 input columns (128 windows; window 127 needs col 254)
    for b in (0, 15):
        xs = x.detach()[b:b + 1, :, :256].cpu().double()
        p64 = [m.pre_conv.weight, m.pre_conv.bias, m.quantum_weights, m.post_conv.weight, m.post_conv.bias]
        yo = qo.qconv1d_forward(xs, *[p.detach().cpu().double() for p in p64], K=3, S=2, P=1)
        assert (y.detach()[b, :, :127].cpu().double() - yo[0, :, :127]).abs().max().item() <= 5e-5


def test_errors(cuda):
    _lib, qc = _mods()
    m = qc.QuantumConv1d(8, 16, 3, padding=1, n_qubits=4)
    with pytest.raises(RuntimeError):
        m(torch.randn(1, 8, 10))  # CPU input: no CPU path
    m = m.to(cuda)
    with pytest.raises(ValueError):
        m(torch.randn(1, 7, 10, device=cuda))
    with pytest.raises(ValueError):
        m(torch.randn(1, 8, 10, device=cuda).half())
    # q is clamped like the reference (quantum_whisper.py:55)
    assert qc.QuantumConv1d(1, 4, 2, n_qubits=4).n_qubits == 2
    # non-contiguous input is accepted
    x = torch.randn(2, 10, 8, device=cuda).transpose(1, 2)
    assert m(x).shape == (2, 16, 10)
    # zero window -> NaN, exactly like the reference's unguarded normalisation (quantum_whisper.py:74)
    with torch.no_grad():
        m.pre_conv.bias.zero_()
    y = m(torch.zeros(1, 8, 10, device=cuda))
    assert torch.isnan(y).all()
    lib = _lib.load()
    assert lib.qw_conv1d_forward(None, None, None, None, None, None, None, None, 1, 1, 1, 1, 1, 0, 1, 1, 1, 0, None) == -1
    assert b"null" in lib.qw_last_error()
