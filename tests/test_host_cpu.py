"""Host-side mirror of the reference interface (no GPU): constructor, attributes, state_dict layout, name-based
freezing, the encoder mirror against the vendored whisper AudioEncoder (golden fixture), audio helpers."""
import os

import numpy as np
import pytest
import torch

import qasr_ijcnlp_b200 as qw
from qasr_ijcnlp_b200 import audio as qa
from qasr_ijcnlp_b200 import dp


def test_constructor_attributes_and_state_dict_layout():
    # quantum_whisper.py:136-137 call forms
    c1 = qw.QuantumConv1d(80, 384, kernel_size=3, padding=1, n_qubits=4)
    c2 = qw.QuantumConv1d(384, 384, kernel_size=3, stride=2, padding=1, n_qubits=4)
    assert (c1.in_channels, c1.out_channels, c1.kernel_size, c1.stride, c1.padding, c1.n_qubits) == (80, 384, 3, 1, 1, 4)
    assert (c2.stride, c2.padding) == (2, 1)
    sd = c2.state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == {
        "quantum_weights": (4, 3), "pre_conv.weight": (4, 1152), "pre_conv.bias": (4,),
        "post_conv.weight": (384, 4), "post_conv.bias": (384,)}
    # 9 440 trainable floats in the two layers
    assert sum(p.numel() for m in (c1, c2) for p in m.parameters()) == 9440
    assert qw.QuantumConv1d(1, 4, 2, n_qubits=4).n_qubits == 2  # :55 clamp
    assert c1.to("cpu") is c1  # :90-93 returns self
    # extensions keep reference defaults
    assert c1.n_layers == 1 and c1.embedding == "amplitude"
    assert tuple(qw.QuantumConv1d(8, 8, 3, n_qubits=4, n_layers=3).quantum_weights.shape) == (3, 4, 3)
    with pytest.raises(ValueError):
        qw.QuantumConv1d(8, 8, 3, embedding="basis")


def test_same_seed_same_parameters_as_reference_construction_order():
    """pre_conv, post_conv, then randn(q,3) (quantum_whisper.py:58-59,88): replay the RNG stream by hand."""
    torch.manual_seed(123)
    m = qw.QuantumConv1d(5, 7, 3, n_qubits=4)
    torch.manual_seed(123)
    pre = torch.nn.Linear(15, 4)
    post = torch.nn.Linear(4, 7)
    qwts = torch.randn(4, 3)
    assert torch.equal(m.pre_conv.weight, pre.weight) and torch.equal(m.post_conv.bias, post.bias)
    assert torch.equal(m.quantum_weights, qwts)


def test_encoder_names_and_freezing():
    dims = qw.ModelDimensions(n_mels=8, n_audio_ctx=10, n_audio_state=16, n_audio_head=2, n_audio_layer=1)
    model = qw.QuantumWhisperASR(qw.QuantumWhisper(dims, n_qubits=4))
    qw.freeze_non_quantum_layers(model)
    trainable = {n for n, p in model.named_parameters() if p.requires_grad}
    assert "quantum_whisper.encoder.conv1.quantum_weights" in trainable
    assert "quantum_whisper.encoder.conv2.pre_conv.weight" in trainable
    assert all(("conv1" in n) or ("conv2" in n) or ("asr_head" in n) for n in trainable)
    assert not any("blocks" in n for n in trainable)
    t = qw.get_whisper_tiny_dims()
    assert (t.n_mels, t.n_audio_ctx, t.n_audio_state, t.n_audio_head, t.n_audio_layer) == (80, 1500, 384, 6, 4)


def test_classical_encoder_mirror_matches_vendored_whisper(golden_dir):
    g = np.load(os.path.join(golden_dir, "audio_encoder.npz"))
    enc = qw.AudioEncoder(n_mels=8, n_ctx=10, n_state=16, n_head=2, n_layer=2).eval()
    sd = {k.replace("__", "."): torch.from_numpy(g[k]) for k in g.files if k not in ("x", "y")}
    enc.load_state_dict(sd, strict=True)  # same module / parameter names as whisper/model.py
    with torch.no_grad():
        y = enc(torch.from_numpy(g["x"]))
    assert (y - torch.from_numpy(g["y"])).abs().max().item() <= 2e-5
    with pytest.raises(AssertionError):
        enc(torch.zeros(1, 8, 24))  # whisper/model.py:197 "incorrect audio shape"


def test_audio_helpers(golden_dir):
    assert (qa.N_SAMPLES, qa.N_FRAMES, qa.HOP_LENGTH, qa.N_FFT) == (480000, 3000, 160, 400)
    a = torch.arange(10.0)
    assert qa.pad_or_trim(a, 4).tolist() == [0, 1, 2, 3]
    assert qa.pad_or_trim(a, 12).tolist() == list(range(10)) + [0, 0]
    assert qa.pad_or_trim(np.ones((2, 3), np.float32), 5).shape == (2, 5)
    assert qa.pad_or_trim(torch.ones(3, 2), 4, axis=0).shape == (4, 2)
    ref = np.load(os.path.join(golden_dir, "mel_filters_ref.npz"))
    for n in (80, 128):
        fb = qa.mel_filters("cpu", n)
        dense = np.zeros((n, 201), np.float32)
        dense[ref[f"rows_{n}"].astype(int), ref[f"cols_{n}"].astype(int)] = ref[f"vals_{n}"]
        assert fb.dtype == torch.float32 and np.abs(fb.numpy() - dense).max() <= 1e-7
        assert np.array_equal(fb.numpy() != 0, dense != 0)  # same support as the reference asset
    with pytest.raises(AssertionError):
        qa.mel_filters("cpu", 64)


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 16, 128, 1001):
        for world in (1, 2, 3, 8):
            spans = [dp.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        dp.shard_range(4, 2, 2)


def test_grad_bucket_single_process():
    lin = torch.nn.Linear(3, 2)
    frozen = torch.nn.Linear(2, 2)
    for p in frozen.parameters():
        p.requires_grad = False
    b = dp.GradBucket(list(lin.parameters()) + list(frozen.parameters()))
    assert b.numel == 8 and b.nbytes == 32
    lin(torch.ones(1, 3)).sum().backward()
    want = [p.grad.clone() for p in lin.parameters()]
    b.allreduce_mean()
    for p, w in zip(lin.parameters(), want):
        assert torch.equal(p.grad, w) and p.grad.data_ptr() >= b.flat.data_ptr()


def test_reference_checkpoint_formats_load(tmp_path):
    """a reference-style checkpoint (bare state_dict or utils.save_model dict, with the text decoder's
    extra keys) loads into the B200 modules key for key."""
    from qasr_ijcnlp_b200 import checkpoint as ck
    dims = qw.ModelDimensions(n_mels=8, n_audio_ctx=10, n_audio_state=16, n_audio_head=2, n_audio_layer=1)
    torch.manual_seed(3)
    src = qw.QuantumWhisperClassifier(qw.QuantumWhisper(dims, n_qubits=4), 35)
    sd = {k: v.clone() for k, v in src.state_dict().items()}
    sd["quantum_whisper.decoder.token_embedding.weight"] = torch.zeros(4, 16)  # reference-only (text decoder)
    torch.manual_seed(4)
    dst = qw.QuantumWhisperClassifier(qw.QuantumWhisper(dims, n_qubits=4), 35)
    missing, ref_only = ck.load_reference_state_dict(dst, sd)
    assert missing == [] and ref_only == ["quantum_whisper.decoder.token_embedding.weight"]
    for (ka, a), (kb, b) in zip(src.state_dict().items(), dst.state_dict().items()):
        assert ka == kb and torch.equal(a, b)
    # utils.save_model format, through a file
    path = str(tmp_path / "ckpt.pth")
    torch.save({"model_state_dict": sd, "model_info": {"epoch": 3}, "training_history": {"loss": [1.0]}}, path)
    torch.manual_seed(5)
    dst2 = qw.QuantumWhisperClassifier(qw.QuantumWhisper(dims, n_qubits=4), 35)
    _, hist, info = ck.load_reference_checkpoint(dst2, path)
    assert info["epoch"] == 3 and hist["loss"] == [1.0]
    assert torch.equal(dst2.quantum_whisper.encoder.conv2.quantum_weights, src.quantum_whisper.encoder.conv2.quantum_weights)
    # a checkpoint trained with another n_qubits does not fit: strict load says so
    other = qw.QuantumWhisperClassifier(qw.QuantumWhisper(dims, n_qubits=3), 35)
    with pytest.raises(RuntimeError, match="shape mismatch"):
        ck.load_reference_state_dict(other, sd)
