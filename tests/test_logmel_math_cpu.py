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
    for n, fn in ((8, lm.lm_dft8), (25, lm.lm_dft25)):
        r, i = rs.randn(n).astype(np.float32), rs.randn(n).astype(np.float32)
        ref = np.fft.fft(r.astype(np.float64) + 1j * i)
        fn(_p(r), _p(i))
        assert np.abs(r + 1j * i - ref).max() <= 4e-6


@pytest.mark.parametrize("G", [1, 4, 8])
def test_frame_power_spectrum(lm, G):
    rs = np.random.RandomState(G)
    fr = np.concatenate([rs.randn(6, 400), np.ones((1, 400)), np.eye(400)[[0, 1, 399]]]).astype(np.float32)
    ref = np.abs(np.fft.rfft(fr.astype(np.float64) * lo.hann_periodic(), axis=1)) ** 2
    p = np.zeros((len(fr), 201), np.float32)
    lm.lm_host_power(_p(fr), _p(p), len(fr), G)
    # fp32 FFT: error relative to the frame's largest bin
    assert (np.abs(p - ref).max(axis=1) / np.maximum(ref.max(axis=1), 1e-30)).max() <= 1e-6


def test_tap_map_is_conflict_free_and_exact(lm):
    """staging index of tile sample s is s + (s >> 5); frame f (lane f) tap j reads sample 160 f + j, i.e. index
    165 f + j + (j >> 5): the two formulas agree and the 32 lanes hit 32 distinct banks for every tap."""
    for j in (0, 1, 31, 32, 159, 160, 161, 319, 320, 398, 399):
        idx = [lm.lm_tap_index(f, j) for f in range(32)]
        assert idx == [lm.lm_skew(160 * f + j) for f in range(32)]
        assert len({i % 32 for i in idx}) == 32
    assert lm.lm_skew(5359) < 5528
