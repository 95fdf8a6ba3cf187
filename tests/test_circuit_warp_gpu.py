"""Parity of the thread-group cooperative circuit kernels (n_qubits 1..12, amplitude / angle embedding, n_layers) and of
the composed general QuantumConv1d path (q > 4 or angle embedding) with the fp64 oracle.  BASELINE.json config 4 regime.
Tolerances: fp32 <= 1e-5 abs on <Z_i> and per-window gradients (gradients wrt pre scale with 1/||pre|| under amplitude
embedding: the bound is relative to that factor), fp64 validation build <= 1e-10."""
import pytest
import torch

from oracle import qconv_oracle as qo

pytestmark = pytest.mark.gpu

EMB = {"amplitude": qo.EMB_AMPLITUDE, "angle": qo.EMB_ANGLE}


def _circuit_case(cuda, q, n_layers, embedding, dtype, tol, W):
    from qasr_ijcnlp_b200 import quantum_circuit
    g = torch.Generator().manual_seed(1000 * q + 10 * n_layers + len(embedding))
    pre64 = torch.randn(W, q, generator=g, dtype=torch.float64)
    w64 = torch.randn(n_layers, q, 3, generator=g, dtype=torch.float64)
    cot64 = torch.randn(W, q, generator=g, dtype=torch.float64)
    pre_o = pre64.to(dtype).double().requires_grad_(True)
    w_o = w64.to(dtype).double().requires_grad_(True)
    out_o = qo.circuit_expvals(pre_o, w_o, EMB[embedding])
    gp_o, gw_o = torch.autograd.grad(out_o, [pre_o, w_o], cot64.to(dtype).double())
    pre = pre64.to(dtype).to(cuda).requires_grad_(True)
    w = w64.to(dtype).to(cuda).requires_grad_(True)
    out = quantum_circuit(pre, w, n_layers=n_layers, embedding=embedding)
    gp, gw = torch.autograd.grad(out, [pre, w], cot64.to(dtype).to(cuda))
    assert (out.detach().cpu().double() - out_o.detach()).abs().max().item() <= tol
    scale = (1.0 / pre_o.detach().norm(dim=1, keepdim=True)).clamp(min=1.0) if embedding == "amplitude" else 1.0
    assert ((gp.cpu().double() - gp_o) / scale).abs().max().item() <= tol * (1 if dtype == torch.float64 else 2)
    # weight gradient is a sum over W windows: bound relative to its largest entry
    assert (gw.cpu().double() - gw_o).abs().max().item() <= tol * max(1.0, gw_o.abs().max().item()) * 4


@pytest.mark.parametrize("q", list(range(1, 13)))
@pytest.mark.parametrize("embedding", ["amplitude", "angle"])
def test_circuit_f64_all_qubit_counts(cuda, q, embedding):
    W = 333 if q <= 9 else 41  # ragged against every group size (128, 64, ..., 4, 1 windows per CTA iteration)
    _circuit_case(cuda, q, 2, embedding, torch.float64, 1e-10, W)


@pytest.mark.parametrize("q", [4, 5, 6, 8, 9, 10, 11, 12])
@pytest.mark.parametrize("n_layers", [1, 4])
@pytest.mark.parametrize("embedding", ["amplitude", "angle"])
def test_circuit_f32(cuda, q, n_layers, embedding):
    W = 515 if q <= 9 else 37
    _circuit_case(cuda, q, n_layers, embedding, torch.float32, 1e-5, W)


def test_thread_and_group_kernels_agree_q4(cuda):
    """q = 4 amplitude goes through the per-thread kernels, q = 4 angle through the group kernels; zero-angle trainable
    layer + angle embedding has the closed form <Z_0> = cos(pre_0) (RZ RY |0>)."""
    from qasr_ijcnlp_b200 import quantum_circuit
    pre = torch.linspace(-3, 3, 64, device=cuda)[:, None].repeat(1, 4).contiguous()
    out = quantum_circuit(pre, torch.zeros(4, 3, device=cuda), embedding="angle")
    # CNOT chain turns wire i into the parity of wires 0..i: <Z_0 ... Z_i> = cos(pre)^(i+1)
    for i in range(4):
        assert (out[:, i] - torch.cos(pre[:, 0]) ** (i + 1)).abs().max().item() <= 1e-5


def test_large_batch_property_q10(cuda):
    """Config-4 sized batch at q = 10: scale invariance of the amplitude embedding and linearity of the backward."""
    from qasr_ijcnlp_b200 import quantum_circuit
    g = torch.Generator().manual_seed(4)
    W = 1 << 14
    pre = torch.randn(W, 10, generator=g).to(cuda).requires_grad_(True)
    w = torch.randn(2, 10, 3, generator=g).to(cuda).requires_grad_(True)
    out = quantum_circuit(pre, w, n_layers=2)
    assert out.abs().max().item() <= 1.0 + 1e-5
    assert (quantum_circuit(pre.detach() * -2.5, w.detach(), n_layers=2) - out.detach()).abs().max().item() <= 1e-5
    cot = torch.randn(W, 10, generator=g).to(cuda)
    gp, gw = torch.autograd.grad(out, [pre, w], cot, retain_graph=True)
    gp2, gw2 = torch.autograd.grad(out, [pre, w], 3.0 * cot)
    assert (gp2 - 3 * gp).abs().max().item() <= 1e-5 * max(1.0, gp.abs().max().item())
    assert (gw2 - 3 * gw).abs().max().item() <= 1e-4 * max(1.0, gw.abs().max().item())
    assert ((pre.detach() * gp).sum(dim=1)).abs().max().item() <= 1e-4  # pre . dL/dpre = 0


GEN_GEOMS = [
    # (B, C, L, K, S, P, O, q, n_layers, embedding)
    (2, 80, 96, 3, 1, 1, 384, 6, 1, "amplitude"),
    (2, 48, 101, 3, 2, 1, 96, 8, 2, "amplitude"),
    (1, 9, 64, 5, 3, 2, 33, 10, 1, "amplitude"),
    (2, 16, 50, 3, 1, 1, 40, 4, 1, "angle"),
    (2, 7, 37, 4, 2, 3, 130, 5, 2, "angle"),
    (1, 6, 40, 3, 2, 0, 12, 12, 1, "amplitude"),
    (2, 40, 70, 3, 3, 1, 24, 6, 1, "amplitude"),    # kernel_size 3 with a stride the one-pass pre_conv^T kernel divides at run time
    (1, 384, 260, 3, 2, 1, 384, 7, 1, "amplitude"),  # conv2 channel counts: 12 channel chunks, 3 position tiles (one ragged)
]


@pytest.mark.parametrize("geom", GEN_GEOMS)
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 5e-5)])
def test_general_layer_vs_oracle(cuda, geom, dtype, tol):
    from qasr_ijcnlp_b200 import quantum_conv1d
    B, C, L, K, S, P, O, q, n_layers, embedding = geom
    params64 = qo.make_params(C, O, K, q, n_layers=n_layers, seed=7)
    g = torch.Generator().manual_seed(8)
    x64 = torch.randn(B, C, L, generator=g, dtype=torch.float64)
    Lo = qo.out_length(L, K, S, P)
    gy64 = torch.randn(B, O, Lo, generator=g, dtype=torch.float64)
    ref = qo.qconv1d_grads(x64.to(dtype).double(), [p.to(dtype).double() for p in params64], gy64.to(dtype).double(), K, S, P,
                           embedding=EMB[embedding])
    x = x64.to(dtype).to(cuda).requires_grad_(True)
    ps = [p.to(dtype).to(cuda).requires_grad_(True) for p in params64]
    y = quantum_conv1d(x, *ps, kernel_size=K, stride=S, padding=P, n_layers=n_layers, embedding=embedding)
    grads = torch.autograd.grad(y, [x] + ps, gy64.to(dtype).to(cuda))
    rel = lambda a, b: (a - b).abs().max().item() / max(1.0, b.abs().max().item())
    assert rel(y.detach().cpu().double(), ref["y"]) <= tol
    for name, gr in zip(["x", "w_pre", "b_pre", "qweights", "w_post", "b_post"], grads):
        assert rel(gr.cpu().double(), ref[name]) <= tol, name


def test_module_with_six_qubits_trains(cuda):
    """The reference's only quantum knob is --n_qubits: the drop-in module must take any value, not just 4."""
    from qasr_ijcnlp_b200 import QuantumConv1d
    torch.manual_seed(0)
    m = QuantumConv1d(80, 384, 3, padding=1, n_qubits=6).to(cuda)
    assert tuple(m.quantum_weights.shape) == (6, 3) and tuple(m.pre_conv.weight.shape) == (6, 240)
    x = torch.randn(2, 80, 300, device=cuda)
    y = m(x)
    assert y.shape == (2, 384, 300)
    y.square().mean().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    with torch.no_grad():
        assert m(x).shape == (2, 384, 300)  # inference (no autograd) goes through the same composed path


def test_general_layer_full_size_conv2_geometry(cuda):
    """The one-pass pre_conv^T kernel, the prefetched gy pass and the 8-way row reduction of the general path at the stem's conv2
    geometry (384 -> 384, stride 2, 3000 -> 1500 positions: 12 channel chunks x 24 position tiles, 592 + partial rows), n_qubits 6:
    output and all six gradients against the fp64 oracle."""
    from qasr_ijcnlp_b200 import quantum_conv1d
    B, C, L, K, S, P, O, q = 2, 384, 3000, 3, 2, 1, 384, 6
    params64 = qo.make_params(C, O, K, q, n_layers=1, seed=11)
    g = torch.Generator().manual_seed(12)
    x64 = torch.randn(B, C, L, generator=g, dtype=torch.float64)
    Lo = qo.out_length(L, K, S, P)
    gy64 = torch.randn(B, O, Lo, generator=g, dtype=torch.float64)
    ref = qo.qconv1d_grads(x64.float().double(), [p.float().double() for p in params64], gy64.float().double(), K, S, P)
    x = x64.float().to(cuda).requires_grad_(True)
    ps = [p.float().to(cuda).requires_grad_(True) for p in params64]
    y = quantum_conv1d(x, *ps, kernel_size=K, stride=S, padding=P)
    grads = torch.autograd.grad(y, [x] + ps, gy64.float().to(cuda))
    rel = lambda a, b: (a - b).abs().max().item() / max(1.0, b.abs().max().item())
    assert rel(y.detach().cpu().double(), ref["y"]) <= 5e-5
    for name, gr in zip(["x", "w_pre", "b_pre", "qweights", "w_post", "b_post"], grads):
        assert rel(gr.cpu().double(), ref[name]) <= 5e-5, name


def test_general_layer_batch_512_last_utterance_matches_single(cuda):
    """General path (n_qubits 6) at conv1's geometry and batch 512: the (B, 384, 3000) output is 2.36 GB, so element byte offsets pass
    2^31 in every streaming kernel of the path.  Output and grad_x of utterances 0 and 511 must equal those of the utterance alone."""
    from qasr_ijcnlp_b200 import QuantumConv1d
    torch.manual_seed(512)
    m = QuantumConv1d(80, 384, 3, padding=1, n_qubits=6).to(cuda)
    B, L = 512, 3000
    x = torch.randn(B, 80, L, device=cuda, requires_grad=True)
    y = m(x)
    (gx,) = torch.autograd.grad(y, [x], torch.ones_like(y))
    for b in (0, B - 1):
        xb = x[b:b + 1].detach().clone().requires_grad_(True)
        yb = m(xb)
        (gxb,) = torch.autograd.grad(yb, [xb], torch.ones_like(yb))
        assert (y[b:b + 1] - yb).abs().max().item() <= 2e-6 * max(1.0, yb.abs().max().item()), b
        assert (gx[b:b + 1] - gxb).abs().max().item() <= 2e-6 * max(1.0, gxb.abs().max().item()), b
