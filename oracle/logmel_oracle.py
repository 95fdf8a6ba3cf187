"""fp64 numpy restatement of Whisper's log-mel front end (TEST INFRASTRUCTURE ONLY).

Follows ``/root/reference/whisper/whisper/audio.py``:
  * ``:65-88``   ``pad_or_trim`` (right zero-pad / trim to 480 000 samples)
  * ``:91-107``  ``mel_filters`` -- the asset ``assets/mel_filters.npz`` was produced by
                 ``librosa.filters.mel(sr=16000, n_fft=400, n_mels=80|128)`` (docstring at
                 ``:96-103``); ``mel_filterbank`` below restates that published algorithm (Slaney
                 mel scale, Slaney area normalisation) and is checked bit-for-bit-ish (<=1e-7)
                 against the asset in the build container (tests/golden/make_golden.py).
  * ``:110-157`` ``log_mel_spectrogram``: periodic Hann(400), ``torch.stft(n_fft=400, hop=160,
                 center=True, pad_mode="reflect")`` -> 3001 frames, drop the last (``:149``),
                 power, mel matmul, ``log10(clamp(.,1e-10))``, ``max(., max-8)``, ``(.+4)/4``.

The ``max`` at ``:155`` is taken over the whole tensor handed in; the reference always hands in ONE
utterance, so the batched form takes it per utterance.

Pinned against the real vendored implementation run in the build container: fixtures in
``tests/golden/logmel_*.npz`` (generator committed beside them).
"""
from __future__ import annotations

import numpy as np

SAMPLE_RATE = 16000
N_FFT = 400
HOP_LENGTH = 160
N_SAMPLES = 480000
N_FRAMES = 3000
N_FREQ = N_FFT // 2 + 1


def pad_or_trim(audio: np.ndarray, length: int = N_SAMPLES) -> np.ndarray:
    """audio.py:65-88 along the last axis."""
    n = audio.shape[-1]
    if n > length:
        audio = audio[..., :length]
    elif n < length:
        pad = [(0, 0)] * audio.ndim
        pad[-1] = (0, length - n)
        audio = np.pad(audio, pad)
    return audio


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(n_mels: int = 80, sr: int = SAMPLE_RATE, n_fft: int = N_FFT) -> np.ndarray:
    """(n_mels, n_fft//2+1) float32 Slaney-style triangular filterbank == librosa.filters.mel
    defaults (htk=False, norm="slaney", fmin=0, fmax=sr/2), as audio.py:96-103 says the asset was made."""
    n_freq = n_fft // 2 + 1
    fftfreqs = np.linspace(0.0, sr / 2.0, n_freq)
    mel_pts = np.linspace(_hz_to_mel(0.0), _hz_to_mel(sr / 2.0), n_mels + 2)
    hz_pts = _mel_to_hz(mel_pts)
    fdiff = np.diff(hz_pts)
    ramps = hz_pts[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, n_freq))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (hz_pts[2:n_mels + 2] - hz_pts[:n_mels])
    w *= enorm[:, None]
    return w.astype(np.float32)


def hann_periodic(n: int = N_FFT) -> np.ndarray:
    """torch.hann_window(n) (periodic=True default), audio.py:147."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def frame_sample_indices(n_samples: int = N_SAMPLES, n_frames: int = N_FRAMES) -> np.ndarray:
    """(n_frames, N_FFT) int64: ORIGINAL sample index read by (frame t, tap j) after
    center=True reflect padding by N_FFT//2 (torch.stft defaults, audio.py:148).  Integer map that
    the CUDA kernel must reproduce bit-exactly."""
    pad = N_FFT // 2
    pos = np.arange(n_frames)[:, None] * HOP_LENGTH + np.arange(N_FFT)[None, :] - pad
    pos = np.where(pos < 0, -pos, pos)
    pos = np.where(pos >= n_samples, 2 * (n_samples - 1) - pos, pos)
    return pos


def log_mel_spectrogram(audio: np.ndarray, n_mels: int = 80, filters: np.ndarray | None = None) -> np.ndarray:
    """audio: (n_samples,) or (B, n_samples) float -> (n_mels, n_frames) / (B, n_mels, n_frames) float64.

    n_frames = n_samples // 160 (the 3001st STFT frame is dropped, audio.py:149)."""
    a = np.asarray(audio, dtype=np.float64)
    single = a.ndim == 1
    if single:
        a = a[None]
    B, n = a.shape
    n_frames = n // HOP_LENGTH
    filt = (mel_filterbank(n_mels) if filters is None else filters).astype(np.float64)
    win = hann_periodic()
    idx = frame_sample_indices(n, n_frames)
    out = np.empty((B, n_mels, n_frames))
    for b in range(B):
        frames = a[b][idx] * win[None, :]  # (T,400)
        spec = np.fft.rfft(frames, n=N_FFT, axis=1)  # (T,201)
        power = spec.real**2 + spec.imag**2  # :149
        mel = filt @ power.T  # :151-152
        logm = np.log10(np.maximum(mel, 1e-10))  # :154
        logm = np.maximum(logm, logm.max() - 8.0)  # :155 (per utterance)
        out[b] = (logm + 4.0) / 4.0  # :156
    return out[0] if single else out
