"""Run-to-run bitwise reproducibility = the suite's race check (compute-sanitizer's racecheck is not available on the GPU pool): every
reduction in the library has a fixed order (per-CTA partial rows summed by the finalize kernels, no floating-point atomics), so the
same call on the same inputs must give the same bits however the CTAs are scheduled.  Each case runs several times with unrelated
work in between (different L2 / scheduling state) on two streams."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _perturb(cuda, i):
    a = torch.randn(512 + 64 * i, 512, device=cuda)
    (a @ a.t()).sum().item()
    torch.empty(48 << 20, device=cuda, dtype=torch.uint8).fill_(i)  # push earlier lines out of part of L2


def _repeat(cuda, fn, n=5):
    side = torch.cuda.Stream(device=cuda)
    ref = None
    for i in range(n):
        _perturb(cuda, i)
        if i % 2:
            side.wait_stream(torch.cuda.current_stream(cuda))
            with torch.cuda.stream(side):
                got = fn()
            torch.cuda.current_stream(cuda).wait_stream(side)
        else:
            got = fn()
        torch.cuda.synchronize()
        got = [g.clone() for g in got]
        if ref is None:
            ref = got
        else:
            for k, (a, b) in enumerate(zip(got, ref)):
                assert torch.equal(a, b), f"run {i}, result {k}: {(a != b).sum().item()} elements differ"


@pytest.mark.parametrize("B,L", [(16, 3000), (3, 96), (5, 1000)])
def test_training_stem_step_is_bitwise_reproducible(cuda, B, L):
    import qasr_ijcnlp_b200 as qw
    from qasr_ijcnlp_b200.quantum_conv1d import stem_train_forward
    torch.manual_seed(B + L)
    c1 = qw.QuantumConv1d(80, 384, 3, padding=1, n_qubits=4).to(cuda)
    c2 = qw.QuantumConv1d(384, 384, 3, stride=2, padding=1, n_qubits=4).to(cuda)
    x = torch.randn(B, 80, L, device=cuda, requires_grad=True)
    cot = torch.randn(B, 384, L // 2, device=cuda)
    prm = list(c1.parameters()) + list(c2.parameters())

    def step():
        y = stem_train_forward(c1, c2, x, gelu=True)
        return [y.detach()] + list(torch.autograd.grad(y, [x] + prm, cot))

    _repeat(cuda, step)


@pytest.mark.parametrize("geom", [(2, 80, 3000, 3, 1, 1, 384, 4), (2, 384, 3000, 3, 2, 1, 384, 4), (2, 12, 77, 5, 3, 2, 20, 4),
                                  (2, 80, 300, 3, 1, 1, 384, 6), (1, 16, 200, 3, 2, 1, 40, 10)])
def test_layer_is_bitwise_reproducible(cuda, geom):
    """Fast path (both stem geometries), generic thread-per-window path, general path (q = 6, 10)."""
    import qasr_ijcnlp_b200 as qw
    B, C, L, K, S, P, O, q = geom
    torch.manual_seed(sum(geom))
    m = qw.QuantumConv1d(C, O, K, stride=S, padding=P, n_qubits=q).to(cuda)
    x = torch.randn(B, C, L, device=cuda, requires_grad=True)
    y0 = m(x)
    cot = torch.randn_like(y0)

    def step():
        y = m(x)
        return [y.detach()] + list(torch.autograd.grad(y, [x] + list(m.parameters()), cot))

    _repeat(cuda, step)


def test_log_mel_and_inference_stem_are_bitwise_reproducible(cuda):
    import qasr_ijcnlp_b200 as qw
    from qasr_ijcnlp_b200 import audio as qa
    from qasr_ijcnlp_b200.quantum_conv1d import fused_stem_forward
    torch.manual_seed(9)
    audio = torch.randn(4, 16000, device=cuda) * 0.1
    lengths = torch.tensor([16000, 15000, 801, 12345], device=cuda, dtype=torch.int32)
    c1 = qw.QuantumConv1d(80, 384, 3, padding=1, n_qubits=4).to(cuda)
    c2 = qw.QuantumConv1d(384, 384, 3, stride=2, padding=1, n_qubits=4).to(cuda)
    pos = torch.randn(1500, 384, device=cuda)

    def run():
        with torch.no_grad():
            mel = qa.log_mel_spectrogram(audio, pad_to=qa.N_SAMPLES, lengths=lengths)
            return [mel, fused_stem_forward(c1, c2, mel, pos)]

    _repeat(cuda, run, n=4)
