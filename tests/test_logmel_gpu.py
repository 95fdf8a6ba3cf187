"""Parity of the fused log-mel kernel (qw_log_mel through the C ABI / qasr_ijcnlp_b200.audio) with the fp64 oracle and
with the golden outputs of the vendored whisper.log_mel_spectrogram (tests/golden/logmel_*.npz).

Tolerance: the reference pins nothing tighter than `mel.max()-mel.min() <= 2` (whisper/tests/test_audio.py:8-19);
SURVEY.md 8d asks <= 1e-4 abs on the (x+4)/4 scale.  Near the 1e-10 power floor a bin is the difference of large
fp32 FFT terms, so a few bins of a loud frame move more; the bound used is 1e-4 on >= 99.9 % of the entries and
2e-3 on all of them against the fp64 oracle (the vendored fp32 implementation itself differs from fp64 by that much),
and the frame / sample indexing checks are exact."""
import os

import numpy as np
import pytest
import torch

from oracle import logmel_oracle as lo

pytestmark = pytest.mark.gpu


def _close(got, ref, tight=1e-4, loose=2e-3, frac=0.999):
    err = np.abs(np.asarray(got, dtype=np.float64) - ref)
    assert err.max() <= loose, err.max()
    assert (err <= tight).mean() >= frac, (err <= tight).mean()


def test_short_clips_vs_golden_and_oracle(cuda, golden_dir):
    from qasr_ijcnlp_b200 import audio as qa
    g = np.load(os.path.join(golden_dir, "logmel_short.npz"))
    for a, m, n in ((g["a1"], g["m1"], 80), (g["a2"], g["m2"], 80), (g["a1"], g["m1_128"], 128)):
        got = qa.log_mel_spectrogram(torch.from_numpy(a), n_mels=n, device=cuda).cpu().numpy()
        assert got.shape == m.shape and got.dtype == np.float32
        _close(got, m.astype(np.float64))                      # vendored whisper output
        _close(got, lo.log_mel_spectrogram(a, n_mels=n))       # fp64 oracle
        assert got.max() - got.min() <= 2.0                    # whisper/tests/test_audio.py:17


def test_30s_full_size_vs_golden(cuda, golden_dir):
    from qasr_ijcnlp_b200 import audio as qa
    g = np.load(os.path.join(golden_dir, "logmel_30s.npz"))
    rs = np.random.RandomState(int(g["seed"]))
    full = (0.1 * rs.standard_normal(480000)).astype(np.float32)
    stride = int(g["stride"])
    sc = qa.pad_or_trim(torch.from_numpy(full[:16000]))        # Speech-Commands-shaped: 1 s + zero padding
    batch = torch.stack([torch.from_numpy(full), sc]).to(cuda)
    mel = qa.log_mel_spectrogram(batch).cpu().numpy()
    assert mel.shape == (2, 80, 3000)
    _close(mel[0][:, ::stride], g["full_sub"].astype(np.float64))
    _close(np.concatenate([mel[0][:, :4], mel[0][:, -4:]], axis=1), g["full_edges"].astype(np.float64))  # reflect edges
    _close(mel[1][:, ::stride], g["sc_sub"].astype(np.float64))
    for k, st in ((0, g["full_stats"]), (1, g["sc_stats"])):
        assert abs(mel[k].min() - st[0]) <= 2e-4 and abs(mel[k].max() - st[1]) <= 2e-4 and abs(mel[k].mean() - st[2]) <= 2e-5
    # batched call == per-utterance calls, bit for bit (max taken per utterance)
    for k in range(2):
        assert np.array_equal(mel[k], qa.log_mel_spectrogram(batch[k]).cpu().numpy())


@pytest.mark.parametrize("n", [1600, 5920, 16000, 160 * 33, 1000, 16001, 5999, 16159, 250])
def test_random_lengths_vs_oracle(cuda, n):
    from qasr_ijcnlp_b200 import audio as qa
    rs = np.random.RandomState(n)
    a = (rs.standard_normal((3, n)) * np.array([[1.0], [0.01], [5.0]])).astype(np.float32)
    got = qa.log_mel_spectrogram(torch.from_numpy(a).to(cuda)).cpu().numpy()
    assert got.shape == (3, 80, n // 160)
    _close(got, lo.log_mel_spectrogram(a))


def test_frame_indexing_exact(cuda):
    """A single unit impulse at sample p lights exactly the frames whose 400-tap window (after reflect padding) contains
    p with a non-zero Hann weight: integer frame map == oracle's frame_sample_indices, incl. both reflected edges and
    the dropped last frame (audio.py:149)."""
    from qasr_ijcnlp_b200 import audio as qa
    n = 160 * 70
    T = n // 160
    idx = lo.frame_sample_indices(n, T)              # (T, 400)
    win = lo.hann_periodic()
    ps = [0, 1, 37, 199, 200, 201, 5000, n - 201, n - 200, n - 41, n - 40, n - 2, n - 1]
    a = np.zeros((len(ps), n), np.float32)
    for r, p in enumerate(ps):
        a[r, p] = 1.0
    got = qa.log_mel_spectrogram(torch.from_numpy(a).to(cuda)).cpu().numpy()
    for r, p in enumerate(ps):
        hit = ((idx == p) & (win[None, :] > 1e-6)).any(axis=1)      # frames that see the impulse
        floor = got[r].min()
        lit = (got[r] > floor + 1e-6).any(axis=0)
        # a frame seeing the impulse only through a ~0 Hann weight may sit on the floor; a frame not seeing it must
        assert not (lit & ~((idx == p).any(axis=1))).any()
        assert (lit | ~hit).all()
    _close(got, lo.log_mel_spectrogram(a), tight=2e-4)


def test_errors(cuda):
    from qasr_ijcnlp_b200 import audio as qa
    assert qa.log_mel_spectrogram(torch.zeros(1000, device=cuda)).shape == (80, 6)  # any length, like the reference (T = n // 160)
    with pytest.raises(ValueError):
        qa.log_mel_spectrogram(torch.zeros(1, 2, 1600, device=cuda))
    with pytest.raises(RuntimeError):
        qa.log_mel_spectrogram(torch.zeros(1600))                    # CPU tensor, no device: no CPU path
    with pytest.raises(Exception):
        qa.log_mel_spectrogram(torch.zeros(160, device=cuda))        # <= 200 samples: reflect padding undefined
    # all-zero audio: every bin on the 1e-10 floor -> (-10 + 4) / 4 = -1.5 everywhere, like the reference
    z = qa.log_mel_spectrogram(torch.zeros(3200, device=cuda))
    assert torch.all(z == -1.5)


def test_encoder_forward_config2_shape(cuda):
    """BASELINE configs[1] shape: Speech-Commands clips -> log-mel -> quantum encoder -> 35-class head, on the GPU."""
    import qasr_ijcnlp_b200 as qw
    from qasr_ijcnlp_b200 import _lib
    from qasr_ijcnlp_b200 import audio as qa
    torch.manual_seed(1)
    model = qw.QuantumWhisperClassifier(qw.QuantumWhisper(qw.get_whisper_tiny_dims(), n_qubits=4), 35).to(cuda).eval()
    audio = qa.pad_or_trim(0.1 * torch.randn(2, 16000), qa.N_SAMPLES).to(cuda)
    n0 = _lib.launch_count()
    with torch.no_grad():
        mel = qa.log_mel_spectrogram(audio)
        logits = model(mel)
    assert mel.shape == (2, 80, 3000) and logits.shape == (2, 35) and torch.isfinite(logits).all()
    assert _lib.launch_count() - n0 in (2 + 2, 3 + 2)  # (filterbank prep once per device) + stft + finish, conv1 fwd, conv2 fwd


def test_one_shot_entry_point_equals_prepared(cuda):
    """qw_log_mel (analysis + run in one call) and qw_log_mel_prepare / qw_log_mel_prepared give identical bits."""
    import ctypes
    from qasr_ijcnlp_b200 import _lib
    from qasr_ijcnlp_b200 import audio as qa
    lib = _lib.load()
    a = 0.1 * torch.randn(3, 16000, device=cuda)
    want = qa.log_mel_spectrogram(a)
    filt = qa.mel_filters(cuda, 80)
    mel = torch.empty(3, 80, 100, device=cuda)
    n = lib.qw_log_mel_workspace_bytes(3, 16000, 80)
    ws = torch.empty(n, device=cuda, dtype=torch.uint8)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    st = lib.qw_log_mel(p(a), p(filt), p(mel), p(ws), n, 3, 16000, 80, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(st, "qw_log_mel")
    assert torch.equal(mel, want)
    # a dense (non-triangular) filterbank takes the uncached path: compare with a plain matmul of the power spectrum
    dense = torch.rand(80, 201, device=cuda) * 0.01
    st = lib.qw_log_mel(p(a), p(dense), p(mel), p(ws), n, 3, 16000, 80, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(st, "qw_log_mel")
    win = torch.hann_window(400, device=cuda)
    spec = torch.stft(a, 400, 160, window=win, return_complex=True)[..., :-1].abs() ** 2
    ref = torch.clamp(dense @ spec, min=1e-10).log10()
    ref = torch.maximum(ref, ref.amax(dim=(1, 2), keepdim=True) - 8.0)
    ref = (ref + 4.0) / 4.0
    assert (mel - ref).abs().max().item() <= 2e-4


@pytest.mark.parametrize("n_in,n_out", [(16000, 480000), (16000, 16000 * 3), (5000, 8000), (48000, 16000), (16001, 480000), (333, 4000)])
def test_fused_pad_or_trim_equals_host_padding(cuda, n_in, n_out):
    """log_mel_spectrogram(audio, pad_to=N) == log_mel_spectrogram(pad_or_trim(audio, N)) BIT FOR BIT (zero padding
    / trimming done inside the kernel, tiles in the padding skip the FFT), and both match the oracle."""
    from qasr_ijcnlp_b200 import audio as qa
    rs = np.random.RandomState(n_in + n_out)
    a = (0.1 * rs.standard_normal((3, n_in))).astype(np.float32)
    dev = torch.from_numpy(a).to(cuda)
    fused = qa.log_mel_spectrogram(dev, pad_to=n_out)
    host = qa.log_mel_spectrogram(qa.pad_or_trim(dev, n_out))
    assert fused.shape == (3, 80, n_out // 160)
    assert torch.equal(fused, host)
    _close(fused.cpu().numpy(), lo.log_mel_spectrogram(lo.pad_or_trim(a, n_out)))


def test_ragged_batch_with_lengths_and_collate(cuda):
    """Variable-length clips: collate_clips ships (B, max_len) + lengths; every utterance equals its own pad_or_trim'd call."""
    from qasr_ijcnlp_b200 import audio as qa
    rs = np.random.RandomState(5)
    clips = [(0.1 * rs.standard_normal(n)).astype(np.float32) for n in (16000, 9000, 12345, 400)]
    batch, lens = qa.collate_clips(clips, cuda)
    assert batch.shape == (4, 16000) and lens.tolist() == [16000, 9000, 12345, 400]
    got = qa.log_mel_spectrogram(batch, pad_to=qa.N_SAMPLES, lengths=lens)
    assert got.shape == (4, 80, 3000)
    for i, c in enumerate(clips):
        want = qa.log_mel_spectrogram(qa.pad_or_trim(torch.from_numpy(c).to(cuda), qa.N_SAMPLES))
        assert torch.equal(got[i], want), i
    _close(got[1].cpu().numpy(), lo.log_mel_spectrogram(lo.pad_or_trim(clips[1])))
    with pytest.raises(ValueError):
        qa.log_mel_spectrogram(batch, pad_to=qa.N_SAMPLES, lengths=lens[:2])


def test_selective_second_pass_equals_unconditional(cuda):
    """audio.py:155-156 in two forms.  With the full per-call workspace the stft kernel stores (v + 4) / 4 and the second pass only
    touches tiles whose minimum lies below max - 8; with the B-float workspace of the first interface it stores v and an
    unconditional pass applies the formula.  Same bits -- on noise (no tile touched), on a clip whose quiet tail / zero padding
    sits below the floor (write-only and read-modify-write tiles), and against the oracle."""
    import ctypes
    from qasr_ijcnlp_b200 import _lib
    from qasr_ijcnlp_b200 import audio as qa
    lib = _lib.load()
    rs = np.random.RandomState(11)
    n = 160 * 320
    a = (0.1 * rs.standard_normal((4, n))).astype(np.float32)
    a[1, n // 3:] *= 1e-6                      # 120 dB down: below max - 8 -> clamped
    a[2, n // 2:] = 0.0                        # digital silence: tiles entirely on the floor
    a[3] *= np.linspace(1.0, 1e-5, n, dtype=np.float32)   # a slow fade: tiles that straddle the floor
    dev = torch.from_numpy(a).to(cuda)
    prep = qa._prepared_filters(cuda, 80)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    outs = []
    for nbytes in (4 * 4, lib.qw_log_mel_call_workspace_bytes(4, n)):
        mel = torch.empty(4, 80, n // 160, device=cuda)
        ws = torch.empty(max(nbytes, 256), device=cuda, dtype=torch.uint8)
        st = lib.qw_log_mel_prepared(p(dev), p(prep), p(mel), p(ws), nbytes, 4, n, 80,
                                     ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(st, "qw_log_mel_prepared")
        outs.append(mel)
    assert torch.equal(outs[0], outs[1])
    assert torch.equal(outs[1], qa.log_mel_spectrogram(dev))
    ref = lo.log_mel_spectrogram(a)
    _close(outs[1].cpu().numpy(), ref)
    assert (ref[1:] == ref[1:].min(axis=(1, 2), keepdims=True)).mean() > 0.2   # the clamp really is active in these clips


def test_batch_512_full_length_equals_single_utterances(cuda):
    """BASELINE config 5's largest point (batch 512 x 480 000 samples: 983 MB of audio, byte offsets past 2^31): utterances 0, 255 and
    511 of the batched call must be BIT-identical to the same utterance run alone (the per-utterance maximum and every frame are
    computed independently of the batch)."""
    from qasr_ijcnlp_b200 import audio as qa
    g = torch.Generator(device=cuda).manual_seed(512)
    a = torch.randn(512, qa.N_SAMPLES, device=cuda, generator=g) * 0.05
    a[511, 100000:] = 0.0  # silent tail: the selective second pass has to leave those tiles at the clamp floor
    mel = qa.log_mel_spectrogram(a)
    assert mel.shape == (512, 80, 3000) and torch.isfinite(mel).all()
    for b in (0, 255, 511):
        assert torch.equal(mel[b], qa.log_mel_spectrogram(a[b].clone())), b
