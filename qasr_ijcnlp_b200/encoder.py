"""Host-side callers of the hot path: the Whisper audio encoder with a quantum stem.

Mirrors, with the same module / parameter names (so reference checkpoints load with ``strict=True`` and
``freeze_non_quantum_layers`` selects the same tensors):

  * ``AudioEncoder``           /root/reference/whisper/whisper/model.py:174-204 (caller context)
  * ``QuantumAudioEncoder``    /root/reference/quantum_whisper.py:130-144
  * ``QuantumWhisper``         /root/reference/quantum_whisper.py:146-165 (encoder side only: ``embed_audio``;
                               the text decoder is downstream of the path and out of scope)
  * ``get_whisper_tiny_dims``  /root/reference/quantum_whisper.py:167-181
  * ``freeze_non_quantum_layers``  /root/reference/quantum_whisper.py:320-341
  * ``QuantumWhisperClassifier``   /root/reference/train_quantum_whisper.py:146-169 (config-2 harness)
  * ``CharASRHead``            fresh char-level LSTM decoder for the config-3 harness (the reference's
                               ``librispeech_asr.py:132-184`` head is non-functional as shipped)

Everything after the two QuantumConv1d layers is plain PyTorch (cuBLAS / SDPA), exactly as in the reference.
Checked against the vendored whisper AudioEncoder through ``tests/golden/audio_encoder.npz``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn as nn
import torch.nn.functional as F

from .quantum_conv1d import QuantumConv1d, fused_stem_eligible, fused_stem_forward, stem_train_forward


@dataclass
class ModelDimensions:
    n_mels: int
    n_audio_ctx: int
    n_audio_state: int
    n_audio_head: int
    n_audio_layer: int
    n_vocab: int = 51865
    n_text_ctx: int = 448
    n_text_state: int = 384
    n_text_head: int = 6
    n_text_layer: int = 4


def get_whisper_tiny_dims() -> ModelDimensions:
    return ModelDimensions(n_mels=80, n_audio_ctx=1500, n_audio_state=384, n_audio_head=6, n_audio_layer=4,
                           n_vocab=51865, n_text_ctx=448, n_text_state=384, n_text_head=6, n_text_layer=4)


def sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> torch.Tensor:
    """Fixed positional table: [sin(t * w_i) | cos(t * w_i)], w_i geometric in 1..1/max_timescale."""
    if channels % 2:
        raise ValueError("channels must be even")
    half = channels // 2
    step = math.log(max_timescale) / (half - 1)
    freqs = torch.exp(-step * torch.arange(half))
    angle = torch.arange(length)[:, None] * freqs[None, :]
    return torch.cat([angle.sin(), angle.cos()], dim=1)


class LayerNorm(nn.LayerNorm):
    def forward(self, x):
        return super().forward(x.float()).type(x.dtype)


class Linear(nn.Linear):
    def forward(self, x):
        b = None if self.bias is None else self.bias.to(x.dtype)
        return F.linear(x, self.weight.to(x.dtype), b)


class MultiHeadAttention(nn.Module):
    def __init__(self, n_state: int, n_head: int):
        super().__init__()
        self.n_head = n_head
        self.query = Linear(n_state, n_state)
        self.key = Linear(n_state, n_state, bias=False)
        self.value = Linear(n_state, n_state)
        self.out = Linear(n_state, n_state)

    def forward(self, x):
        B, T, D = x.shape
        split = lambda t: t.view(B, T, self.n_head, D // self.n_head).transpose(1, 2)
        a = F.scaled_dot_product_attention(split(self.query(x)), split(self.key(x)), split(self.value(x)))
        return self.out(a.transpose(1, 2).reshape(B, T, D))


class ResidualAttentionBlock(nn.Module):
    def __init__(self, n_state: int, n_head: int):
        super().__init__()
        self.attn = MultiHeadAttention(n_state, n_head)
        self.attn_ln = LayerNorm(n_state)
        self.mlp = nn.Sequential(Linear(n_state, 4 * n_state), nn.GELU(), Linear(4 * n_state, n_state))
        self.mlp_ln = LayerNorm(n_state)

    def forward(self, x):
        x = x + self.attn(self.attn_ln(x))
        return x + self.mlp(self.mlp_ln(x))


class AudioEncoder(nn.Module):
    """Classical Whisper encoder (stem = two nn.Conv1d); the quantum subclass swaps the stem."""

    def __init__(self, n_mels: int, n_ctx: int, n_state: int, n_head: int, n_layer: int):
        super().__init__()
        self.conv1 = nn.Conv1d(n_mels, n_state, kernel_size=3, padding=1)
        self.conv2 = nn.Conv1d(n_state, n_state, kernel_size=3, stride=2, padding=1)
        self.register_buffer("positional_embedding", sinusoids(n_ctx, n_state))
        self.blocks = nn.ModuleList([ResidualAttentionBlock(n_state, n_head) for _ in range(n_layer)])
        self.ln_post = LayerNorm(n_state)

    def forward(self, x):
        x = F.gelu(self.conv1(x))
        x = F.gelu(self.conv2(x))
        x = x.permute(0, 2, 1)
        if x.shape[1:] != self.positional_embedding.shape:
            raise AssertionError("incorrect audio shape")  # whisper/model.py:197
        x = (x + self.positional_embedding).to(x.dtype)
        for blk in self.blocks:
            x = blk(x)
        return self.ln_post(x)


class QuantumAudioEncoder(AudioEncoder):
    def __init__(self, n_mels: int, n_ctx: int, n_state: int, n_head: int, n_layer: int, n_qubits: int = 4, **qkw):
        super().__init__(n_mels, n_ctx, n_state, n_head, n_layer)
        self.conv1 = QuantumConv1d(n_mels, n_state, kernel_size=3, padding=1, n_qubits=n_qubits, **qkw)
        self.conv2 = QuantumConv1d(n_state, n_state, kernel_size=3, stride=2, padding=1, n_qubits=n_qubits, **qkw)

        self.fused_stem = True  # inference only: both layers + GELUs + permute + positional embedding in two kernels

    def to(self, *args, **kwargs):
        super().to(*args, **kwargs)
        return self

    def forward(self, x):
        # when no autograd graph is being recorded the stem runs fused (qw_stem_forward); training and any
        # shape outside the fast-path regime take the operator-by-operator path of AudioEncoder.forward
        if self.fused_stem and not torch.is_grad_enabled() and fused_stem_eligible(self.conv1, self.conv2, x):
            x = fused_stem_forward(self.conv1, self.conv2, x, self.positional_embedding)
            for blk in self.blocks:
                x = blk(x)
            return self.ln_post(x)
        if not self.fused_stem:
            return super().forward(x)
        # training (or a shape outside the fused inference kernel): both layers with their GELUs in ONE forward kernel
        # (qw_stem_train_forward) when the shape is the Whisper stem's, else the two operators with their GELUs fused into them
        # (QuantumConv1d.forward_gelu: no separate activation pass in either direction); the rest as AudioEncoder.forward
        x = stem_train_forward(self.conv1, self.conv2, x, gelu=True)
        x = x.permute(0, 2, 1)
        if x.shape[1:] != self.positional_embedding.shape:
            raise AssertionError("incorrect audio shape")  # whisper/model.py:197
        x = (x + self.positional_embedding).to(x.dtype)
        for blk in self.blocks:
            x = blk(x)
        return self.ln_post(x)


class QuantumWhisper(nn.Module):
    """Encoder side of the reference's QuantumWhisper (embed_audio is what both task heads call)."""

    def __init__(self, dims: ModelDimensions, n_qubits: int = 4, **qkw):
        super().__init__()
        self.dims = dims
        self.encoder = QuantumAudioEncoder(dims.n_mels, dims.n_audio_ctx, dims.n_audio_state, dims.n_audio_head,
                                           dims.n_audio_layer, n_qubits, **qkw)

    def to(self, *args, **kwargs):
        super().to(*args, **kwargs)
        return self

    def embed_audio(self, mel):
        return self.encoder(mel)

    forward = embed_audio

    @property
    def device(self):
        return next(self.parameters()).device


def freeze_non_quantum_layers(model: nn.Module) -> nn.Module:
    """Train only parameters whose name contains conv1 / conv2 / asr_head (quantum_whisper.py:325-333)."""
    for name, p in model.named_parameters():
        p.requires_grad = any(tag in name for tag in ("conv1", "conv2", "asr_head"))
    return model


class QuantumWhisperClassifier(nn.Module):
    """mean-pool over time + Linear(n_state, n_classes) (train_quantum_whisper.py:146-169)."""

    def __init__(self, quantum_whisper: QuantumWhisper, num_classes: int = 35):
        super().__init__()
        self.quantum_whisper = quantum_whisper
        self.classifier = nn.Linear(quantum_whisper.dims.n_audio_state, num_classes)

    def forward(self, mel):
        feats = self.quantum_whisper.embed_audio(mel)
        return self.classifier(feats.mean(dim=1))


CHAR_VOCAB = ["<PAD>", "<UNK>", "<START>", "<END>", " ", "'"] + [chr(c) for c in range(ord("a"), ord("z") + 1)]


class CharASRHead(nn.Module):
    """Char-level LSTM decoder over mean+attention-pooled encoder states (config-3 harness).

    Vocabulary and special ids follow librispeech_asr.py:102-117 (<PAD>=0,<UNK>=1,<START>=2,<END>=3);
    README.md:48-51 of the reference describes an LSTM decoder, which is what is built here.  Teacher forcing:
    input tokens (B, T) -> logits (B, T, V); loss is CE(ignore_index=0) against the shifted targets
    (train_quantum_whisper_asr.py:133-134)."""

    def __init__(self, n_state: int = 384, vocab_size: int = len(CHAR_VOCAB), hidden: int = 256, num_layers: int = 2):
        super().__init__()
        self.embed = nn.Embedding(vocab_size, hidden, padding_idx=0)
        self.ctx_q = nn.Linear(hidden, n_state)
        self.lstm = nn.LSTM(hidden + n_state, hidden, num_layers=num_layers, batch_first=True)
        self.proj = nn.Linear(hidden, vocab_size)

    def forward(self, audio_feats, tokens):
        e = self.embed(tokens)  # (B,T,H)
        att = torch.softmax(self.ctx_q(e) @ audio_feats.transpose(1, 2) / math.sqrt(audio_feats.shape[-1]), dim=-1)
        ctx = att @ audio_feats  # (B,T,D)
        h, _ = self.lstm(torch.cat([e, ctx], dim=-1))
        return self.proj(h)


class QuantumWhisperASR(nn.Module):
    def __init__(self, quantum_whisper: QuantumWhisper, vocab_size: int = len(CHAR_VOCAB), num_layers: int = 2):
        super().__init__()
        self.quantum_whisper = quantum_whisper
        self.asr_head = CharASRHead(quantum_whisper.dims.n_audio_state, vocab_size, num_layers=num_layers)

    def forward(self, mel, tokens):
        return self.asr_head(self.quantum_whisper.embed_audio(mel), tokens)
