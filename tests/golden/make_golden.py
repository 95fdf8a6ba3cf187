"""Generates the golden fixtures in this directory.  Run ONLY in the build container, where
/root/reference exists:  python tests/golden/make_golden.py

Fixtures (all small):
  logmel_short.npz     vendored whisper.log_mel_spectrogram on 1 s / 0.37 s clips (full output)
  logmel_30s.npz       vendored output on a 30 s clip, every 61st frame + global stats
  mel_filters_ref.npz  the reference asset's nonzero pattern + values (sparse) for 80/128 mels
  audio_encoder.npz    vendored whisper AudioEncoder (tiny dims) state_dict + input + output
  quantum_audio_encoder.npz  the vendored AudioEncoder SUBCLASSED exactly as quantum_whisper.py:130-137 does (conv1 / conv2 swapped
                       for a QuantumConv1d), with the fp64 oracle standing in for the PennyLane layer: state_dict + input + output.
                       The GPU test loads that state_dict (strict) into qasr_ijcnlp_b200.QuantumAudioEncoder and must reproduce y.
  qconv_kat.json       the known-answer vectors of SURVEY.md 8c (scratch fp64 restatement by the
                       surveyor; PennyLane itself is not installable -> "parity unpinned")
Inputs are drawn with numpy RandomState so they can be regenerated anywhere.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, "/root/reference/whisper")
from oracle import qconv_oracle as qo  # noqa: E402
import whisper.audio as wa  # noqa: E402
import whisper.model as wm  # noqa: E402


class OracleQuantumConv1d(torch.nn.Module):
    """CPU stand-in for the reference layer (quantum_whisper.py:45-128): same constructor, construction order and parameter
    names; forward = the fp64 oracle (PennyLane is not installable)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, n_qubits=4):
        super().__init__()
        self.kernel_size, self.stride, self.padding = kernel_size, stride, padding
        self.n_qubits = min(n_qubits, in_channels * kernel_size)
        self.pre_conv = torch.nn.Linear(in_channels * kernel_size, self.n_qubits)
        self.post_conv = torch.nn.Linear(self.n_qubits, out_channels)
        self.quantum_weights = torch.nn.Parameter(torch.randn(self.n_qubits, 3))

    def forward(self, x):
        p = [t.detach().double() for t in (self.pre_conv.weight, self.pre_conv.bias, self.quantum_weights,
                                           self.post_conv.weight, self.post_conv.bias)]
        return qo.qconv1d_forward(x.double(), *p, K=self.kernel_size, S=self.stride, P=self.padding, cast_fp32=True).to(x.dtype)


def quantum_audio_encoder_fixture():
    class QuantumAudioEncoder(wm.AudioEncoder):  # quantum_whisper.py:130-137, verbatim structure
        def __init__(self, n_mels, n_ctx, n_state, n_head, n_layer, n_qubits=4):
            super().__init__(n_mels, n_ctx, n_state, n_head, n_layer)
            self.conv1 = OracleQuantumConv1d(n_mels, n_state, kernel_size=3, padding=1, n_qubits=n_qubits)
            self.conv2 = OracleQuantumConv1d(n_state, n_state, kernel_size=3, stride=2, padding=1, n_qubits=n_qubits)

    torch.manual_seed(11)
    enc = QuantumAudioEncoder(n_mels=8, n_ctx=12, n_state=16, n_head=2, n_layer=2).eval()
    x = torch.randn(2, 8, 24)
    with torch.no_grad():
        y = enc(x)
    sd = {k.replace(".", "__"): v.numpy() for k, v in enc.state_dict().items()}
    np.savez_compressed(os.path.join(HERE, "quantum_audio_encoder.npz"), x=x.numpy(), y=y.numpy(), **sd)


def main():
    quantum_audio_encoder_fixture()
    rs = np.random.RandomState(1234)
    # ---- log-mel, short clips, full output
    a1 = (0.1 * rs.standard_normal(16000)).astype(np.float32)
    a2 = (0.3 * np.sin(2 * np.pi * 440 * np.arange(5920) / 16000) + 0.01 * rs.standard_normal(5920)).astype(np.float32)
    m1 = wa.log_mel_spectrogram(torch.from_numpy(a1)).numpy()
    m2 = wa.log_mel_spectrogram(torch.from_numpy(a2)).numpy()
    m1_128 = wa.log_mel_spectrogram(torch.from_numpy(a1), n_mels=128).numpy()
    np.savez_compressed(os.path.join(HERE, "logmel_short.npz"), a1=a1, a2=a2, m1=m1, m2=m2, m1_128=m1_128)
    # ---- log-mel, 30 s (speech-commands style: 1 s of signal, zero-padded by pad_or_trim)
    rs = np.random.RandomState(4321)
    full = (0.1 * rs.standard_normal(480000)).astype(np.float32)
    sc = np.asarray(wa.pad_or_trim(torch.from_numpy(full[:16000]))).astype(np.float32)
    mf = wa.log_mel_spectrogram(torch.from_numpy(full)).numpy()
    ms = wa.log_mel_spectrogram(torch.from_numpy(sc)).numpy()
    np.savez_compressed(
        os.path.join(HERE, "logmel_30s.npz"),
        seed=4321, stride=61,
        full_sub=mf[:, ::61], sc_sub=ms[:, ::61],
        full_stats=np.array([mf.min(), mf.max(), mf.mean(), np.abs(mf).sum()], dtype=np.float64),
        sc_stats=np.array([ms.min(), ms.max(), ms.mean(), np.abs(ms).sum()], dtype=np.float64),
        full_edges=np.concatenate([mf[:, :4], mf[:, -4:]], axis=1),
    )
    # ---- mel filter asset, sparse
    f = np.load("/root/reference/whisper/whisper/assets/mel_filters.npz")
    out = {}
    for n in (80, 128):
        m = f[f"mel_{n}"]
        r, c = np.nonzero(m)
        out[f"rows_{n}"] = r.astype(np.int16)
        out[f"cols_{n}"] = c.astype(np.int16)
        out[f"vals_{n}"] = m[r, c]
    np.savez_compressed(os.path.join(HERE, "mel_filters_ref.npz"), **out)
    # ---- vendored AudioEncoder at tiny dims
    torch.manual_seed(7)
    enc = wm.AudioEncoder(n_mels=8, n_ctx=10, n_state=16, n_head=2, n_layer=2).eval()
    x = torch.randn(2, 8, 20)
    with torch.no_grad():
        y = enc(x)
    sd = {k.replace(".", "__"): v.numpy() for k, v in enc.state_dict().items()}
    np.savez_compressed(os.path.join(HERE, "audio_encoder.npz"), x=x.numpy(), y=y.numpy(), **sd)
    # ---- QuantumConv1d known answers
    kat = {
        "source": "SURVEY.md 8c (scratch fp64 restatement; PennyLane not installable -> parity unpinned)",
        "pre": [0.5, -1.0, 2.0, 0.25],
        "kat0": {"quantum_weights": [[0.0] * 3] * 4,
                 "out": [1.0, 1.0, -0.5294117647058822, -0.8823529411764706]},
        "kat1": {"quantum_weights": [[0.1, 0.2, 0.3], [-0.4, 0.5, -0.6], [0.7, -0.8, 0.9], [1.0, 1.1, -1.2]],
                 "out": [0.980066577841, 0.860089338205, -0.183997000514, 0.368386588580],
                 "cotangent": [1.0, -2.0, 3.0, 0.5],
                 "grad_pre": [1.637374648515, -0.865828803050, -0.779861932685, -0.499169047749],
                 "grad_quantum_weights": [[0.0, 0.2245844360784, 0.0], [0.0, 1.140666697130, 0.0],
                                          [-0.3051022628793, -1.678297288399, 0.0],
                                          [-0.1329337373520, 0.2657969494641, 0.0]]},
    }
    with open(os.path.join(HERE, "qconv_kat.json"), "w") as fh:
        json.dump(kat, fh, indent=1)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
