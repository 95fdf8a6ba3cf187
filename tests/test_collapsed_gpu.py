"""The collapsed quadratic-form evaluation of the amplitude-embedded circuit (SURVEY.md 8a iii; opt-in `simulator="collapsed"`)
against the statevector kernels ON THE DEVICE and against the fp64 oracle: the second, independent device-side check of the
simulator (the suite's other oracles run on the CPU).  out_i = xh^T M_i xh with M read off the statevector kernel on q(q+1)/2
probe windows; weight gradients flow back through the statevector adjoint kernel on those probes."""
import pytest
import torch

from oracle import qconv_oracle as qo

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("q", [1, 2, 3, 4, 6, 8, 12])
@pytest.mark.parametrize("n_layers", [1, 2])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.float64, 1e-10)])
def test_collapsed_equals_statevector(cuda, q, n_layers, dtype, tol):
    from qasr_ijcnlp_b200 import quantum_circuit
    g = torch.Generator().manual_seed(1000 * q + n_layers)
    W = 1537 if q <= 8 else 259  # ragged vs the 128-thread blocks
    pre = torch.randn(W, q, generator=g, dtype=torch.float64).to(dtype).to(cuda).requires_grad_(True)
    shape = (q, 3) if n_layers == 1 else (n_layers, q, 3)
    w = torch.randn(*shape, generator=g, dtype=torch.float64).to(dtype).to(cuda).requires_grad_(True)
    cot = torch.randn(W, q, generator=g, dtype=torch.float64).to(dtype).to(cuda)
    out_sv = quantum_circuit(pre, w, n_layers=n_layers)
    gp_sv, gw_sv = torch.autograd.grad(out_sv, [pre, w], cot)
    out_c = quantum_circuit(pre, w, n_layers=n_layers, simulator="collapsed")
    gp_c, gw_c = torch.autograd.grad(out_c, [pre, w], cot)
    assert (out_c - out_sv).abs().max().item() <= 2 * tol
    scale = (1.0 / pre.detach().norm(dim=1, keepdim=True)).clamp(min=1.0)
    assert ((gp_c - gp_sv) / scale).abs().max().item() <= 4 * tol
    # weight gradients are sums over W windows: relative to the largest entry
    assert (gw_c - gw_sv).abs().max().item() <= 8 * tol * max(1.0, gw_sv.abs().max().item())
    if q <= 6:  # and against the CPU oracle
        pre_o = pre.detach().cpu().double().requires_grad_(True)
        w_o = w.detach().cpu().double().requires_grad_(True)
        out_o = qo.circuit_expvals(pre_o, w_o)
        gp_o, gw_o = torch.autograd.grad(out_o, [pre_o, w_o], cot.cpu().double())
        assert (out_c.detach().cpu().double() - out_o.detach()).abs().max().item() <= 2 * tol
        assert (gw_c.cpu().double() - gw_o).abs().max().item() <= 8 * tol * max(1.0, gw_o.abs().max().item())


def test_collapsed_matrices_are_the_quadratic_forms_of_the_oracle(cuda):
    """M[i] = Re(U[:, :q]^H Z'_i U[:, :q]): compare with the dense-unitary oracle (oracle/qconv_oracle.py::dense_unitary)."""
    import numpy as np

    from qasr_ijcnlp_b200.quantum_conv1d import collapsed_matrices
    q = 4
    w = torch.randn(2, q, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(3))
    M = collapsed_matrices(w.to(cuda), q, 2).cpu().numpy()
    U = qo.dense_unitary(w.numpy())[:, :q]
    ks = np.arange(1 << q)
    for i in range(q):
        z = 1.0 - 2.0 * ((ks >> (q - 1 - i)) & 1)
        assert np.abs(M[i] - np.real(U.conj().T @ (z[:, None] * U))).max() <= 1e-12


def test_collapsed_rejects_angle_embedding(cuda):
    from qasr_ijcnlp_b200 import quantum_circuit
    pre = torch.randn(8, 4, device=cuda)
    w = torch.randn(4, 3, device=cuda)
    with pytest.raises(ValueError):
        quantum_circuit(pre, w, embedding="angle", simulator="collapsed")
    with pytest.raises(ValueError):
        quantum_circuit(pre, w, simulator="tensor-network")
