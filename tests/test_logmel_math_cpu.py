"""The log-mel kernel's per-frame arithmetic (csrc/qw_logmel_math.cuh -- the same source the CUDA kernel compiles)
run on the host through tests/native/logmel_host_check.cpp: FFT factorisation, twiddle tables, slot maps and the
shared-memory tap map against numpy.  No GPU, no product code path: test infrastructure only."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import logmel_oracle as lo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lm(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("lm") / "lmcheck.so")
    subprocess.run(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-o", out,
                    os.path.join(ROOT, "tests", "native", "logmel_host_check.cpp")], check=True)
    return ctypes.CDLL(out)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_small_dfts(lm):
    rs = np.random.RandomState(0)
    r, i = rs.randn(25).astype(np.float32), rs.randn(25).astype(np.float32)
    ref = np.fft.fft(r.astype(np.float64) + 1j * i)
    lm.lm_dft25(_p(r), _p(i))
    assert np.abs(r + 1j * i - ref).max() <= 4e-6
    x = rs.randn(16).astype(np.float32)
    re, im = np.zeros(9, np.float32), np.zeros(9, np.float32)
    lm.lm_rfft16(_p(x), _p(re), _p(im))
    assert np.abs(re + 1j * im - np.fft.rfft(x.astype(np.float64))).max() <= 4e-6


def test_frame_power_spectrum(lm):
    """16 x 25 real FFT (pass A: one real FFT16 per n2 + W400 twiddles; pass B: one DFT25 per k1 = 0..8, mirrored bins for
    k2 >= 13) against numpy, and the bin map: every one of the 201 bins is written exactly once."""
    rs = np.random.RandomState(3)
    fr = np.concatenate([rs.randn(6, 400), np.ones((1, 400)), np.eye(400)[[0, 1, 399]]]).astype(np.float32)
    ref = np.abs(np.fft.rfft(fr.astype(np.float64) * lo.hann_periodic(), axis=1)) ** 2
    p = np.zeros((len(fr), 201), np.float32)
    w = np.zeros((len(fr), 201), np.int32)
    lm.lm_host_power(_p(fr), _p(p), _p(w), len(fr))
    assert (w == 1).all()
    # fp32 FFT: error relative to the frame's largest bin
    assert (np.abs(p - ref).max(axis=1) / np.maximum(ref.max(axis=1), 1e-30)).max() <= 1e-6


def test_work_layout_is_conflict_free():
    """Shared-memory work array of the kernel: value (plane pl, n2) of frame fr sits at float (pl * 25 + n2) * 33 + fr.
    Pass A stores with one n2 per lane (25 lanes, fixed fr and pl), pass B loads with one frame per lane (fixed pl, n2):
    both hit distinct banks."""
    for pl in range(17):
        for fr in (0, 7, 31):
            assert len({((pl * 25 + n2) * 33 + fr) % 32 for n2 in range(25)}) == 25
        for n2 in (0, 13, 24):
            assert len({((pl * 25 + n2) * 33 + fr) % 32 for fr in range(32)}) == 32
