"""BASELINE.json config 5: fused log-mel + QuantumConv1d stem forward, batch sweep over 30 s synthetic audio.
One JSON line per batch size: kernel times (in-library CUDA events), utt/s of log-mel alone, of mel -> conv1 -> GELU ->
conv2 (operator by operator) and of mel -> fused stem (conv1, GELU, conv2, GELU, permute, positional embedding), HBM fraction of the log-mel kernels (algorithmic bytes 4*(n + 3*80*3000) per utterance).

    python tools/sweep_logmel.py [--out profiles/r1_config5_sweep.jsonl] [--max-batch 512]
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from qasr_ijcnlp_b200 import QuantumConv1d, _lib, fused_stem_forward
from qasr_ijcnlp_b200 import audio as qa
from qasr_ijcnlp_b200.encoder import sinusoids

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=None)
ap.add_argument("--max-batch", type=int, default=512)
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
dev = torch.device("cuda:0")
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pk = os.path.join(root, "MEASURED_PEAKS.json")
HBM = (json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0) * 1e9
torch.manual_seed(0)
conv1 = QuantumConv1d(80, 384, 3, padding=1, n_qubits=4).to(dev)
conv2 = QuantumConv1d(384, 384, 3, stride=2, padding=1, n_qubits=4).to(dev)
pos = sinusoids(1500, 384).to(dev)
rows = []
B = 1
while B <= a.max_batch:
    g = torch.Generator(device=dev).manual_seed(4)
    audio = 0.1 * torch.randn(B, 480000, device=dev, generator=g)
    def stem():
        with torch.no_grad():
            mel = qa.log_mel_spectrogram(audio)
            return conv2(F.gelu(conv1(mel)))
    for _ in range(2): stem()
    torch.cuda.synchronize()
    _lib.profile_read(True); _lib.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters): stem()
    e1.record(); torch.cuda.synchronize()
    _lib.profile_enable(False)
    prof = {k: v[0] / v[1] for k, v in _lib.profile_read(True).items()}
    # with profiling on, events serialise nothing but add small gaps: time the un-instrumented loop too
    e0.record()
    for _ in range(a.iters): stem()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    # the whole encoder front (model.py:193-198 incl. the second GELU, permute, positional embedding) through the fused stem
    def front_fused():
        return fused_stem_forward(conv1, conv2, qa.log_mel_spectrogram(audio), pos)
    for _ in range(2): front_fused()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(a.iters): front_fused()
    e1.record(); torch.cuda.synchronize()
    ms_fused = e0.elapsed_time(e1) / a.iters
    t_mel = prof["logmel_stft_kernel"] + prof["logmel_finish_kernel"]
    bytes_mel = B * 4.0 * (480000 + 3 * 80 * 3000)
    row = {"config": 5, "batch": B, "logmel_stft_ms": round(prof["logmel_stft_kernel"], 4),
           "logmel_finish_ms": round(prof["logmel_finish_kernel"], 4), "conv_fwd_ms_per_launch": round(prof["qconv_fwd_kernel"], 4),
           "logmel_utt_per_s": round(B / (t_mel * 1e-3), 1), "logmel_hbm_frac": round(bytes_mel / (t_mel * 1e-3) / HBM, 4),
           "stem_ms": round(ms, 4), "stem_utt_per_s": round(B / (ms * 1e-3), 1),
           "fused_front_ms": round(ms_fused, 4), "fused_front_utt_per_s": round(B / (ms_fused * 1e-3), 1)}
    rows.append(row)
    print(json.dumps(row), flush=True)
    del audio
    B *= 2
if a.out:
    with open(a.out, "a") as fh:
        for r in rows: fh.write(json.dumps(r) + "\n")
