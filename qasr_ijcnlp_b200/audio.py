"""Log-mel front end on the GPU (host mirror of /root/reference/whisper/whisper/audio.py:65-157).

Same public names and argument meaning as the vendored whisper module for this path:
``pad_or_trim``, ``mel_filters``, ``log_mel_spectrogram`` and the hyper-parameter constants; the computation is
one fused sm_100a kernel (+ a small in-place finish pass) behind ``qw_log_mel`` in ``include/qw.h``.  There is
no CPU path and no torch.stft fallback.

Differences that are deliberate and documented:
  * a 2-D input ``(B, n_samples)`` is a BATCH: the ``max`` of audio.py:155 is taken per utterance (the reference
    only ever passes one utterance);
  * the mel filterbank is generated (``librosa.filters.mel`` defaults, as audio.py:96-103 says the asset was made)
    rather than loaded from ``assets/mel_filters.npz``; it matches the asset to <= 1e-7 (tests/golden);
  * file-path input (ffmpeg, audio.py:25-62) is out of scope: tensors / arrays only.
"""
from __future__ import annotations

import ctypes
from functools import lru_cache
from typing import Optional, Union

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib

# audio.py:12-22
SAMPLE_RATE = 16000
N_FFT = 400
HOP_LENGTH = 160
CHUNK_LENGTH = 30
N_SAMPLES = CHUNK_LENGTH * SAMPLE_RATE  # 480000
N_FRAMES = N_SAMPLES // HOP_LENGTH  # 3000
N_SAMPLES_PER_TOKEN = HOP_LENGTH * 2
FRAMES_PER_SECOND = SAMPLE_RATE // HOP_LENGTH
TOKENS_PER_SECOND = SAMPLE_RATE // N_SAMPLES_PER_TOKEN


def pad_or_trim(array, length: int = N_SAMPLES, *, axis: int = -1):
    """Right zero-pad or trim `axis` to `length` samples (audio.py:65-88); tensors stay on their device."""
    if torch.is_tensor(array):
        if array.shape[axis] > length:
            array = array.index_select(dim=axis, index=torch.arange(length, device=array.device))
        if array.shape[axis] < length:
            pad = [(0, 0)] * array.ndim
            pad[axis] = (0, length - array.shape[axis])
            array = F.pad(array, [p for sizes in pad[::-1] for p in sizes])
        return array
    array = np.asarray(array)
    if array.shape[axis] > length:
        array = array.take(indices=range(length), axis=axis)
    if array.shape[axis] < length:
        pad = [(0, 0)] * array.ndim
        pad[axis] = (0, length - array.shape[axis])
        array = np.pad(array, pad)
    return array


def _slaney_hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    lin = 3.0 * f / 200.0
    log_region = 15.0 + 27.0 * np.log(np.maximum(f, 1e-300) / 1000.0) / np.log(6.4)
    return np.where(f >= 1000.0, log_region, lin)


def _slaney_mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    return np.where(m >= 15.0, 1000.0 * np.exp(np.log(6.4) * (m - 15.0) / 27.0), 200.0 * m / 3.0)


@lru_cache(maxsize=None)
def _mel_filterbank_np(n_mels: int) -> np.ndarray:
    """Slaney-scale, area-normalised triangular filters over the 201 rFFT bins of a 400-point frame at 16 kHz."""
    centres_hz = np.linspace(0.0, SAMPLE_RATE / 2.0, N_FFT // 2 + 1)
    edges_hz = _slaney_mel_to_hz(np.linspace(_slaney_hz_to_mel(0.0), _slaney_hz_to_mel(SAMPLE_RATE / 2.0), n_mels + 2))
    widths = np.diff(edges_hz)
    rising = (centres_hz[None, :] - edges_hz[:-2, None]) / widths[:-1, None]
    falling = (edges_hz[2:, None] - centres_hz[None, :]) / widths[1:, None]
    tri = np.clip(np.minimum(rising, falling), 0.0, None)
    tri *= (2.0 / (edges_hz[2:] - edges_hz[:-2]))[:, None]
    return tri.astype(np.float32)


_FILTER_CACHE = {}
_PREP_CACHE = {}


def _prepared_filters(device, n_mels: int) -> torch.Tensor:
    """Device buffer holding the library's one-time analysis of the filterbank (qw_log_mel_prepare), cached per device."""
    key = (str(device), n_mels)
    if key not in _PREP_CACHE:
        lib = _lib.load()
        filt = mel_filters(device, n_mels)
        n = lib.qw_log_mel_prep_bytes(n_mels)
        prep = torch.empty(n, device=device, dtype=torch.uint8)
        with torch.cuda.device(device):
            st = lib.qw_log_mel_prepare(ctypes.c_void_p(filt.data_ptr()), n_mels, ctypes.c_void_p(prep.data_ptr()), n,
                                        ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(st, "qw_log_mel_prepare")
        torch.cuda.current_stream(device).synchronize()  # later calls may come from other streams
        _PREP_CACHE[key] = prep
    return _PREP_CACHE[key]


def mel_filters(device, n_mels: int) -> torch.Tensor:
    """(n_mels, 201) float32 filterbank on `device` (audio.py:91-107); cached per (device, n_mels)."""
    if n_mels not in (80, 128):
        raise AssertionError(f"Unsupported n_mels: {n_mels}")  # audio.py:103
    key = (str(device), n_mels)
    if key not in _FILTER_CACHE:
        _FILTER_CACHE[key] = torch.from_numpy(_mel_filterbank_np(n_mels)).to(device).contiguous()
    return _FILTER_CACHE[key]


def log_mel_spectrogram(audio: Union[np.ndarray, torch.Tensor], n_mels: int = 80, padding: int = 0,
                        device: Optional[Union[str, torch.device]] = None, *, pad_to: Optional[int] = None,
                        lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
    """audio (n_samples,) or (B, n_samples) -> (n_mels, n_frames) or (B, n_mels, n_frames), n_frames = n_samples // 160.

    Signature of audio.py:110-115 (minus the file-path form); any n_samples > 200 works like the reference (the last, partial
    STFT frame is dropped, audio.py:149).  The tensor must end up on a CUDA device.

    Batched data path (keyword-only, not in the reference): ``pad_to=N`` fuses ``pad_or_trim(audio, N)`` into the
    kernel -- the result equals ``log_mel_spectrogram(pad_or_trim(audio, N))`` but only the stored samples cross PCIe / HBM and
    tiles that lie in the zero padding skip the FFT (a 1 s Speech-Commands clip is 16 000 samples, not 480 000:
    train_quantum_whisper.py:62-77 pads every clip on the CPU).  ``lengths`` (B,) int32 gives the valid samples of every row of a
    ragged batch (see ``collate_clips``)."""
    if isinstance(audio, str):
        raise TypeError("file-path input (ffmpeg) is outside this package's scope: pass samples")
    if not torch.is_tensor(audio):
        audio = torch.from_numpy(np.asarray(audio))
    if device is not None:
        audio = audio.to(device)
    if padding > 0:
        audio = F.pad(audio, (0, padding))
    if not audio.is_cuda:
        raise RuntimeError("log_mel_spectrogram (B200 build) has no CPU path: pass a CUDA tensor or device='cuda'")
    if audio.dim() not in (1, 2):
        raise ValueError(f"expected (n_samples,) or (batch, n_samples), got {tuple(audio.shape)}")
    single = audio.dim() == 1
    a = audio.reshape(1, -1) if single else audio
    a = a.to(torch.float32).contiguous()
    B, n_in = a.shape
    n = int(pad_to) if pad_to is not None else n_in
    if n <= N_FFT // 2 or n < HOP_LENGTH:
        raise ValueError(f"n_samples={n} must be > {N_FFT // 2} (reflect padding of torch.stft) and >= {HOP_LENGTH}")
    if lengths is not None:
        lengths = lengths.to(device=a.device, dtype=torch.int32).contiguous()
        if lengths.shape != (B,):
            raise ValueError(f"lengths must have shape ({B},), got {tuple(lengths.shape)}")
    lib = _lib.load()
    T = n // HOP_LENGTH
    prep = _prepared_filters(a.device, n_mels)
    mel = torch.empty(B, n_mels, T, device=a.device, dtype=torch.float32)
    ws_bytes = lib.qw_log_mel_call_workspace_bytes(B, n)
    ws = torch.empty(ws_bytes, device=a.device, dtype=torch.uint8)
    P = ctypes.c_void_p
    with torch.cuda.device(a.device):
        st = lib.qw_log_mel_padded(P(a.data_ptr()), P(lengths.data_ptr()) if lengths is not None else None, P(prep.data_ptr()),
                                   P(mel.data_ptr()), P(ws.data_ptr()), ws_bytes, B, n_in, n, n_mels,
                                   P(torch.cuda.current_stream().cuda_stream))
    _lib.check(st, "qw_log_mel_padded")
    return mel[0] if single else mel


def collate_clips(clips, device, pin: bool = True):
    """Batched replacement of the reference's per-item CPU preprocessing (Dataset.__getitem__, train_quantum_whisper.py:52-77,
    librispeech_asr.py:53-84: pad_or_trim to 480 000 samples + log-mel per clip): stack variable-length clips into one
    (B, max_len) pinned host buffer, ship only that, and let ``log_mel_spectrogram(..., pad_to=N_SAMPLES, lengths=...)`` pad on the
    device.  Returns (audio (B, max_len) on `device`, lengths (B,) int32 on `device`)."""
    arrs = [torch.as_tensor(np.asarray(c) if not torch.is_tensor(c) else c, dtype=torch.float32).reshape(-1) for c in clips]
    if not arrs:
        raise ValueError("empty batch")
    lens = torch.tensor([min(int(t.numel()), N_SAMPLES) for t in arrs], dtype=torch.int32)
    width = max(int(lens.max().item()), 1)
    width = (width + 3) // 4 * 4  # 16-byte rows
    host = torch.zeros(len(arrs), width, dtype=torch.float32)
    if pin and torch.cuda.is_available():
        host = host.pin_memory()
    for i, t in enumerate(arrs):
        host[i, :lens[i]] = t[:lens[i]]
    return host.to(device, non_blocking=True), lens.to(device, non_blocking=True)
