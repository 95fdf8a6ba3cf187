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


def test_quantum_encoder_state_dict_matches_the_vendored_subclass(golden_dir):
    """`quantum_audio_encoder.npz` is the state_dict of the VENDORED whisper AudioEncoder subclassed exactly as
    quantum_whisper.py:130-137 does: every key and shape must load (strict) into the B200 QuantumAudioEncoder."""
    g = np.load(os.path.join(golden_dir, "quantum_audio_encoder.npz"))
    enc = qw.QuantumAudioEncoder(n_mels=8, n_ctx=12, n_state=16, n_head=2, n_layer=2, n_qubits=4)
    sd = {k.replace("__", "."): torch.from_numpy(g[k]) for k in g.files if k not in ("x", "y")}
    assert set(sd) == set(enc.state_dict())
    enc.load_state_dict(sd, strict=True)
    with pytest.raises(RuntimeError, match="no CPU path"):
        enc(torch.from_numpy(g["x"]))  # the product path has no CPU fallback


@pytest.mark.skipif(not os.path.isdir("/root/reference/whisper"), reason="vendored whisper only exists in the build container")
def test_drop_in_inside_the_real_vendored_audio_encoder():
    """north_star: "a drop-in inside the official whisper/ AudioEncoder".  Subclass the vendored class the way
    quantum_whisper.py:130-137 does, with the B200 QuantumConv1d: construction, positional call form, parameter names,
    name-based freezing and `.to()` behave as in the reference; its forward reaches the layer (which refuses CPU tensors)."""
    import sys
    sys.path.insert(0, "/root/reference/whisper")
    try:
        import whisper.model as wm
    finally:
        sys.path.pop(0)

    class QuantumAudioEncoder(wm.AudioEncoder):
        def __init__(self, n_mels, n_ctx, n_state, n_head, n_layer, n_qubits=4):
            super().__init__(n_mels, n_ctx, n_state, n_head, n_layer)
            self.conv1 = qw.QuantumConv1d(n_mels, n_state, kernel_size=3, padding=1, n_qubits=n_qubits)
            self.conv2 = qw.QuantumConv1d(n_state, n_state, kernel_size=3, stride=2, padding=1, n_qubits=n_qubits)

    torch.manual_seed(0)
    real = QuantumAudioEncoder(8, 12, 16, 2, 2)
    mirror = qw.QuantumAudioEncoder(8, 12, 16, 2, 2)
    assert [n for n, _ in real.named_parameters()] == [n for n, _ in mirror.named_parameters()]
    assert {k: tuple(v.shape) for k, v in real.state_dict().items()} == {k: tuple(v.shape) for k, v in mirror.state_dict().items()}
    mirror.load_state_dict(real.state_dict(), strict=True)
    qw.freeze_non_quantum_layers(real)
    assert {n for n, p in real.named_parameters() if p.requires_grad} == {
        f"conv{i}.{s}" for i in (1, 2) for s in ("quantum_weights", "pre_conv.weight", "pre_conv.bias", "post_conv.weight", "post_conv.bias")}
    assert real.to("cpu") is real
    with pytest.raises(RuntimeError, match="no CPU path"):
        real(torch.randn(1, 8, 24))  # whisper/model.py:193 calls self.conv1(x): the B200 layer is what runs


def test_char_asr_harness_vocab_shapes_and_one_step():
    """SURVEY.md 8-f2 (config-3 harness): vocabulary ids as librispeech_asr.py:102-117, teacher-forcing shapes,
    CE(ignore_index=0) as train_quantum_whisper_asr.py:133, and one optimisation step lowers the loss."""
    assert qw.CHAR_VOCAB[:4] == ["<PAD>", "<UNK>", "<START>", "<END>"]
    assert len(qw.CHAR_VOCAB) == 32 and len(set(qw.CHAR_VOCAB)) == 32
    idx = {c: i for i, c in enumerate(qw.CHAR_VOCAB)}
    text = "the cat's hat"
    ids = [idx["<START>"]] + [idx.get(ch, idx["<UNK>"]) for ch in text] + [idx["<END>"]]
    assert min(ids) >= 2 and idx.get("~", idx["<UNK>"]) == 1  # unknown characters map to <UNK>
    T = 100  # librispeech_asr.py:44 max_text_length
    tokens = torch.zeros(2, T, dtype=torch.long)
    tokens[0, :len(ids)] = torch.tensor(ids)
    tokens[1, :5] = torch.tensor([2, 10, 11, 12, 3])
    torch.manual_seed(0)
    head = qw.CharASRHead(n_state=16, hidden=32, num_layers=2)
    feats = torch.randn(2, 12, 16)
    logits = head(feats, tokens[:, :-1])
    assert logits.shape == (2, T - 1, len(qw.CHAR_VOCAB))
    crit = torch.nn.CrossEntropyLoss(ignore_index=0)
    opt = torch.optim.AdamW(head.parameters(), lr=1e-2, weight_decay=0.01)

    def loss_fn():
        return crit(head(feats, tokens[:, :-1]).reshape(-1, len(qw.CHAR_VOCAB)), tokens[:, 1:].reshape(-1))

    l0 = loss_fn()
    assert torch.isfinite(l0) and abs(l0.item() - np.log(32)) < 1.0
    for _ in range(5):
        opt.zero_grad()
        loss = loss_fn()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(head.parameters(), 1.0)  # train_quantum_whisper_asr.py:175
        opt.step()
    assert loss_fn().item() < l0.item() - 0.05
    # gradient does not flow from padded positions: the embedding row of <PAD> stays zero (padding_idx=0)
    assert head.embed.weight[0].abs().max().item() == 0.0
