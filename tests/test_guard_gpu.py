"""Out-of-bounds stores: every buffer the C ABI writes (outputs, saved tensors, gradients, workspaces) is carved out of one arena
filled with a sentinel byte, with 4 KiB guard bands on both sides; after the calls every guard band must still hold the sentinel.
(compute-sanitizer is not available on the GPU pool, so this is the suite's memcheck for stores; the parity tests cover the values.)
Shapes are the ragged ones: tile ranges that end inside an utterance, L % 32 != 0, a single tile per utterance."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu

SENTINEL = 0x5A
GUARD = 4096


class Arena:
    def __init__(self, dev, nbytes):
        self.buf = torch.full((nbytes,), SENTINEL, device=dev, dtype=torch.uint8)
        self.top = GUARD
        self.used = []

    def take(self, nbytes, align=1024):
        base = self.buf.data_ptr()
        start = (base + self.top + align - 1) // align * align - base
        assert start + nbytes + GUARD <= self.buf.numel(), "arena too small"
        self.used.append((start, start + nbytes))
        self.top = start + nbytes + GUARD
        return self.buf[start:start + nbytes]

    def f32(self, *shape):
        n = 1
        for s in shape:
            n *= s
        return self.take(4 * n).view(torch.float32).view(*shape)

    def assert_guards_intact(self, what):
        torch.cuda.synchronize()
        mask = torch.ones(self.buf.numel(), device=self.buf.device, dtype=torch.bool)
        for a, b in self.used:
            mask[a:b] = False
        bad = ((self.buf != SENTINEL) & mask).nonzero()
        if bad.numel():
            first = int(bad[0])
            owner = min(self.used, key=lambda ab: min(abs(first - ab[0]), abs(first - ab[1])))
            raise AssertionError(f"{what}: {bad.numel()} guard bytes overwritten, first at arena offset {first} "
                                 f"(nearest buffer [{owner[0]}, {owner[1]}), #{self.used.index(owner)})")


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("B,L,fused", [(3, 96, True), (2, 40, True), (5, 1000, True), (1, 3000, True), (2, 200, False), (7, 136, False)])
def test_stem_forward_and_chained_backward_stay_inside_their_buffers(cuda, B, L, fused):
    from qasr_ijcnlp_b200 import _lib
    import qasr_ijcnlp_b200 as qw
    lib = _lib.load()
    torch.manual_seed(B * 1000 + L)
    C, H, O, nl = 80, 384, 384, 1
    c1 = qw.QuantumConv1d(C, H, 3, padding=1, n_qubits=4).to(cuda)
    c2 = qw.QuantumConv1d(H, O, 3, stride=2, padding=1, n_qubits=4).to(cuda)
    prm = lambda m: [t.detach().contiguous() for t in (m.pre_conv.weight, m.pre_conv.bias, m.quantum_weights, m.post_conv.weight,
                                                       m.post_conv.bias)]
    p1, p2 = prm(c1), prm(c2)
    n1 = lib.qw_conv1d_workspace_bytes(B, C, L, 3, 1, 1, H, 4, nl, 4)
    n2 = lib.qw_conv1d_workspace_bytes(B, H, L, 3, 2, 1, O, 4, nl, 4)
    ar = Arena(cuda, 4 * (B * C * L * 2 + B * H * L + 2 * B * O * (L // 2) + 8 * B * L * 2) + n1 + n2 + (4 << 20))
    x = ar.f32(B, C, L)
    x.copy_(torch.randn(B, C, L, device=cuda))
    y1, ps1 = ar.f32(B, H, L), ar.f32(2, B * L, 4)
    y2, ps2 = ar.f32(B, O, L // 2), ar.f32(2, B * (L // 2), 4)
    with torch.cuda.device(cuda):
        if fused:
            st = lib.qw_stem_train_forward(_p(x), *[_p(t) for t in p1], *[_p(t) for t in p2], _p(y1), _p(ps1), _p(y2), _p(ps2),
                                           B, C, L, H, O, nl, 1, _stream())
            _lib.check(st, "qw_stem_train_forward")
        else:
            st = lib.qw_conv1d_forward_act(_p(x), *[_p(t) for t in p1], _p(y1), _p(ps1), B, C, L, 3, 1, 1, H, 4, nl, 0, 1, _stream())
            _lib.check(st, "qw_conv1d_forward_act")
            st = lib.qw_conv1d_forward_act(_p(y1), *[_p(t) for t in p2], _p(y2), _p(ps2), B, H, L, 3, 2, 1, O, 4, nl, 0, 1, _stream())
            _lib.check(st, "qw_conv1d_forward_act")
    ar.assert_guards_intact("forward")
    assert torch.isfinite(y2).all() and torch.isfinite(y1).all()

    gy2 = ar.f32(B, O, L // 2)
    gy2.copy_(torch.randn(B, O, L // 2, device=cuda))
    gx = ar.f32(B, C, L)
    grads1 = [ar.f32(*t.shape) for t in p1]
    grads2 = [ar.f32(*t.shape) for t in p2]
    ws1, ws2 = ar.take(n1), ar.take(n2)
    with torch.cuda.device(cuda):
        st = lib.qw_conv1d_backward_act(_p(gy2), _p(y1), _p(ps2), _p(p2[0]), _p(p2[2]), _p(p2[3]), _p(p2[4]), None,
                                        *[_p(g) for g in grads2], _p(ws2), n2, B, H, L, 3, 2, 1, O, 4, nl, 0, 1, _stream())
        _lib.check(st, "qw_conv1d_backward_act")
        st = lib.qw_conv1d_backward_chained(_p(ws2), _p(p2[0]), O, _p(x), _p(ps1), _p(p1[0]), _p(p1[2]), _p(p1[3]), _p(p1[4]), _p(gx),
                                            *[_p(g) for g in grads1], _p(ws1), n1, B, C, L, 3, 1, 1, H, 4, nl, 0, 1,
                                            None, None, 0, 1, 1.0, _stream())
        _lib.check(st, "qw_conv1d_backward_chained")
    ar.assert_guards_intact("chained backward")
    for g in [gx] + grads1 + grads2:
        assert torch.isfinite(g).all()
        assert g.abs().max().item() < 1e12  # the sentinel pattern 0x5A5A5A5A is the finite float 1.5e16: an unwritten element shows up here


@pytest.mark.parametrize("B,C,L,K,S,P,O,q", [(2, 80, 100, 3, 1, 1, 384, 4), (3, 384, 100, 3, 2, 1, 384, 4), (2, 12, 77, 5, 3, 2, 20, 4),
                                             (2, 40, 70, 3, 2, 1, 24, 6)])
def test_layer_forward_backward_stay_inside_their_buffers(cuda, B, C, L, K, S, P, O, q):
    """The per-layer entry points (fast path, generic thread-per-window path, general q = 6 path) with grad_x requested."""
    from qasr_ijcnlp_b200 import _lib
    from oracle import qconv_oracle as qo
    lib = _lib.load()
    nl = 1
    Lo = qo.out_length(L, K, S, P)
    p = [t.float().to(cuda).contiguous() for t in qo.make_params(C, O, K, q, n_layers=nl, seed=5)]
    n = lib.qw_conv1d_workspace_bytes(B, C, L, K, S, P, O, q, nl, 4)
    ar = Arena(cuda, 4 * (2 * B * C * L + 2 * B * O * Lo + 2 * B * Lo * q + 2 * sum(t.numel() for t in p)) + n + (2 << 20))
    x = ar.f32(B, C, L)
    x.copy_(torch.randn(B, C, L, device=cuda))
    y, ps = ar.f32(B, O, Lo), ar.f32(2, B * Lo, q)
    with torch.cuda.device(cuda):
        st = lib.qw_conv1d_forward(_p(x), *[_p(t) for t in p], _p(y), _p(ps), B, C, L, K, S, P, O, q, nl, 0, _stream())
        _lib.check(st, "qw_conv1d_forward")
    ar.assert_guards_intact("layer forward")
    gy = ar.f32(B, O, Lo)
    gy.copy_(torch.randn(B, O, Lo, device=cuda))
    gx = ar.f32(B, C, L)
    grads = [ar.f32(*t.shape) for t in p]
    ws = ar.take(n)
    with torch.cuda.device(cuda):
        st = lib.qw_conv1d_backward(_p(gy), _p(x), _p(ps), _p(p[0]), _p(p[2]), _p(p[3]), _p(gx), *[_p(g) for g in grads], _p(ws), n,
                                    B, C, L, K, S, P, O, q, nl, 0, _stream())
        _lib.check(st, "qw_conv1d_backward")
    ar.assert_guards_intact("layer backward")
    for g in [gx] + grads:
        assert g.abs().max().item() < 1e12


@pytest.mark.parametrize("B,n_in,n", [(1, 16000, 480000), (3, 16000, 16000), (2, 12345, 32000), (5, 480000, 480000)])
def test_log_mel_stays_inside_its_buffers(cuda, B, n_in, n):
    from qasr_ijcnlp_b200 import _lib
    from qasr_ijcnlp_b200.audio import _prepared_filters
    lib = _lib.load()
    prep = _prepared_filters(cuda, 80)
    T = n // 160
    nws = lib.qw_log_mel_call_workspace_bytes(B, n)
    ar = Arena(cuda, 4 * (B * n_in + B * 80 * T) + nws + 8 * B + (1 << 20))
    a = ar.f32(B, n_in)
    a.copy_(torch.randn(B, n_in, device=cuda) * 0.1)
    lengths = ar.take(4 * B).view(torch.int32)
    lengths.copy_(torch.tensor([max(201, n_in - 37 * i) for i in range(B)], dtype=torch.int32))
    mel = ar.f32(B, 80, T)
    ws = ar.take(max(nws, 16))
    with torch.cuda.device(cuda):
        st = lib.qw_log_mel_padded(_p(a), _p(lengths), _p(prep), _p(mel), _p(ws), nws, B, n_in, n, 80, _stream())
        _lib.check(st, "qw_log_mel_padded")
    ar.assert_guards_intact("log-mel")
    assert torch.isfinite(mel).all() and mel.abs().max().item() < 10.0


@pytest.mark.parametrize("B,L", [(1, 3000), (3, 96), (2, 40), (5, 1000)])
def test_inference_stem_stays_inside_its_buffers(cuda, B, L):
    from qasr_ijcnlp_b200 import _lib
    import qasr_ijcnlp_b200 as qw
    lib = _lib.load()
    torch.manual_seed(B + L)
    C, H, O, nl = 80, 384, 384, 1
    c1 = qw.QuantumConv1d(C, H, 3, padding=1, n_qubits=4).to(cuda)
    c2 = qw.QuantumConv1d(H, O, 3, stride=2, padding=1, n_qubits=4).to(cuda)
    prm = lambda m: [t.detach().contiguous() for t in (m.pre_conv.weight, m.pre_conv.bias, m.quantum_weights, m.post_conv.weight,
                                                       m.post_conv.bias)]
    n = lib.qw_stem_workspace_bytes(B, L)
    ar = Arena(cuda, 4 * (B * C * L + 2 * (L // 2) * O + B * (L // 2) * O) + n + (1 << 20))
    x = ar.f32(B, C, L)
    x.copy_(torch.randn(B, C, L, device=cuda))
    pos = ar.f32(L // 2, O)
    pos.copy_(torch.randn(L // 2, O, device=cuda))
    out = ar.f32(B, L // 2, O)
    ws = ar.take(max(n, 16))
    with torch.cuda.device(cuda):
        st = lib.qw_stem_forward(_p(x), *[_p(t) for t in prm(c1)], *[_p(t) for t in prm(c2)], _p(pos), _p(out), _p(ws), n,
                                 B, C, L, H, O, nl, _stream())
        _lib.check(st, "qw_stem_forward")
    ar.assert_guards_intact("inference stem")
    assert out.abs().max().item() < 1e12


@pytest.mark.parametrize("W,q,nl,emb", [(333, 4, 1, 0), (515, 6, 2, 0), (41, 10, 1, 0), (37, 12, 1, 1), (1000, 5, 4, 1), (1, 4, 1, 0)])
def test_circuit_entry_points_stay_inside_their_buffers(cuda, W, q, nl, emb):
    from qasr_ijcnlp_b200 import _lib
    lib = _lib.load()
    n = lib.qw_circuit_workspace_bytes(W, q, nl, 4)
    ar = Arena(cuda, 4 * (4 * W * q + 2 * nl * q * 3) + n + (1 << 20))
    pre = ar.f32(W, q)
    pre.copy_(torch.randn(W, q, device=cuda))
    w = ar.f32(nl, q, 3)
    w.copy_(torch.randn(nl, q, 3, device=cuda))
    out = ar.f32(W, q)
    with torch.cuda.device(cuda):
        _lib.check(lib.qw_circuit_forward(_p(pre), _p(w), _p(out), W, q, nl, emb, _stream()), "qw_circuit_forward")
    ar.assert_guards_intact("circuit forward")
    assert out.abs().max().item() <= 1.0 + 1e-5
    gout = ar.f32(W, q)
    gout.copy_(torch.randn(W, q, device=cuda))
    gpre, gw = ar.f32(W, q), ar.f32(nl, q, 3)
    ws = ar.take(max(n, 16))
    with torch.cuda.device(cuda):
        _lib.check(lib.qw_circuit_backward(_p(pre), _p(w), _p(gout), _p(gpre), _p(gw), _p(ws), n, W, q, nl, emb, _stream()),
                   "qw_circuit_backward")
    ar.assert_guards_intact("circuit backward")
    assert gpre.abs().max().item() < 1e12 and gw.abs().max().item() < 1e12
