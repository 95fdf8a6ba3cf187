"""The C-ABI library loads on a box with no GPU and exports exactly what include/qw.h declares.
No compute entry point is exercised here (that is the -m gpu suite); only argument validation, which
returns before any CUDA call."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "qw.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
    return sorted(set(re.findall(r"\b(qw_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from qasr_ijcnlp_b200 import _lib
    return _lib.load()


def test_header_and_binding_declare_the_same_symbols():
    from qasr_ijcnlp_b200 import _lib
    assert _header_functions() == sorted(_lib.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    from qasr_ijcnlp_b200 import _lib
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in _header_functions():
        assert hasattr(raw, name), f"libqw_b200.so does not export {name}"
    assert lib.qw_abi_version() == 1


def test_no_torch_types_in_the_abi():
    src = open(os.path.join(ROOT, "include", "qw.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)  # declarations only: comments may name the callers
    assert "torch" not in src.lower() and "tensor" not in src.lower()
    assert "at::" not in src and "c10::" not in src and "#include <torch" not in src


def test_kernel_names(lib):
    names = []
    k = 0
    while True:
        n = lib.qw_kernel_name(k)
        if not n:
            break
        names.append(n.decode())
        k += 1
    assert "qconv_fwd_kernel" in names and "qconv_bwd_pre_kernel" in names and len(names) == len(set(names))


def test_argument_validation_without_a_gpu(lib):
    # null pointers: -1 and a message, before any CUDA call
    st = lib.qw_conv1d_forward(None, None, None, None, None, None, None, None, 1, 1, 8, 3, 1, 1, 4, 1, 1, 0, None)
    assert st == -1 and b"null" in lib.qw_last_error()
    one = ctypes.c_void_p(256)  # never dereferenced: shape validation fails first
    st = lib.qw_conv1d_forward(one, one, one, one, one, one, one, None, 1, 2, 8, 3, 1, 1, 4, 7, 1, 0, None)
    assert st == -1 and b"n_qubits" in lib.qw_last_error()  # q must be <= C*K (quantum_whisper.py:55 clamp is the caller's)
    st = lib.qw_conv1d_forward(one, one, one, one, one, one, one, None, 1, 2, 1, 5, 1, 1, 4, 2, 1, 0, None)
    assert st == -1 and b"kernel_size" in lib.qw_last_error()
    st = lib.qw_conv1d_forward(one, one, one, one, one, one, one, None, 0, 2, 8, 3, 1, 1, 4, 2, 1, 0, None)
    assert st == -1
    st = lib.qw_circuit_forward(one, one, one, 16, 13, 1, 0, None)
    assert st == -2 and b"n_qubits" in lib.qw_last_error()
    st = lib.qw_log_mel(None, None, None, None, 0, 1, 480000, 80, None)
    assert st == -1
    # fused training forward of the stem: null pointers, then shapes outside its regime (-2: run the two layers separately)
    st = lib.qw_stem_train_forward(*([None] * 15), 16, 80, 3000, 384, 384, 1, 1, None)
    assert st == -1 and b"null" in lib.qw_last_error()
    st = lib.qw_stem_train_forward(*([one] * 15), 16, 128, 3000, 384, 384, 1, 1, None)      # C > 96
    assert st == -2 and b"fused regime" in lib.qw_last_error()
    st = lib.qw_stem_train_forward(*([one] * 15), 16, 80, 3004, 384, 384, 1, 1, None)       # L % 8 != 0
    assert st == -2
    st = lib.qw_stem_train_forward(*([one] * 15), 16, 80, 3000, 384, 384, 1, 7, None)       # unknown activation
    assert st == -2 and b"activation" in lib.qw_last_error()
    # chained backward: null pointers, then a circuit size outside the chained regime
    args = lambda q: ([one, one, 384] + [one] * 13 + [1 << 30] + [16, 80, 3000, 3, 1, 1, 384, q, 1, 0] + [0, None, None, 0, 1, ctypes.c_float(1.0), None])
    st = lib.qw_conv1d_backward_chained(None, None, 384, *args(4)[3:])
    assert st == -1 and b"null" in lib.qw_last_error()
    st = lib.qw_conv1d_backward_chained(*args(3))
    assert st == -2 and b"chained regime" in lib.qw_last_error()
    assert lib.qw_stem_train_forward_preferred(16, 3000) == 1 and lib.qw_stem_train_forward_preferred(0, 3000) == 0


def test_workspace_sizes_are_pure_host_arithmetic(lib):
    a = lib.qw_conv1d_workspace_bytes(16, 384, 3000, 3, 2, 1, 384, 4, 1, 4)
    b = lib.qw_conv1d_workspace_bytes(32, 384, 3000, 3, 2, 1, 384, 4, 1, 4)
    assert 0 < a < b
    assert a % 256 == 0
    assert lib.qw_conv1d_workspace_bytes(0, 384, 3000, 3, 2, 1, 384, 4, 1, 4) == 0
    assert lib.qw_circuit_workspace_bytes(1 << 20, 4, 1, 4) > 0
    assert lib.qw_log_mel_workspace_bytes(16, 480000, 80) > 0
    # per-call workspace of the prepared entry points: B utterance maxima + 9 (min, max) pairs per 32-frame tile
    w = lib.qw_log_mel_call_workspace_bytes(16, 480000)
    assert w % 256 == 0 and w >= 16 * 4 + 16 * 94 * 9 * 8
    assert lib.qw_log_mel_call_workspace_bytes(0, 480000) == 0
    assert lib.qw_log_mel_workspace_bytes(16, 480000, 80) >= w + lib.qw_log_mel_prep_bytes(80)


def test_product_path_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under qasr_ijcnlp_b200/ may import it, and there is no CPU fallback."""
    pkg = os.path.join(ROOT, "qasr_ijcnlp_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
    for fn in os.listdir(os.path.join(pkg, "csrc")):
        src = open(os.path.join(pkg, "csrc", fn)).read()
        assert not re.search(r"^\s*#\s*include[^\n]*oracle", src, flags=re.M), fn


def test_cpu_input_fails_loudly():
    import torch

    from qasr_ijcnlp_b200 import QuantumConv1d, quantum_circuit
    from qasr_ijcnlp_b200 import audio as qa

    m = QuantumConv1d(8, 16, 3, padding=1, n_qubits=4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.randn(1, 8, 10))
    with pytest.raises(RuntimeError, match="no CPU path"):
        quantum_circuit(torch.randn(5, 4), torch.randn(4, 3))
    with pytest.raises(RuntimeError, match="no CPU path"):
        qa.log_mel_spectrogram(torch.randn(2, 16000))


def test_new_entry_points_validate_arguments_without_a_gpu(lib):
    """qw_stem_forward / qw_conv1d_backward_dp / the dp size queries: bad arguments come back as negative status before any
    CUDA call; the size queries are pure host arithmetic."""
    st = lib.qw_stem_forward(*([None] * 13), None, 0, 1, 80, 3000, 384, 384, 1, None)
    assert st == -1 and b"null" in lib.qw_last_error()
    assert lib.qw_stem_workspace_bytes(16, 3000) >= 2 * 16 * 3000 * 4 * 4 and lib.qw_stem_workspace_bytes(0, 3000) == 0
    dims = (16, 384, 3000, 3, 2, 1, 384, 4, 1)
    n2, n8 = lib.qw_conv1d_dp_buffer_bytes(*dims, 2), lib.qw_conv1d_dp_buffer_bytes(*dims, 8)
    assert n2 > 0 and n8 == 4 * n2                                   # [2 slots][world][columns] 8-byte words
    assert n2 == 2 * 2 * (1920 + 64 + 4608) * 8                       # columns: gy rows + adjoint rows + pre_conv^T rows
    assert lib.qw_conv1d_dp_buffer_bytes(*dims, 9) == 0               # world <= 8
    assert lib.qw_conv1d_dp_buffer_bytes(16, 384, 3000, 3, 2, 1, 384, 6, 1, 2) == 0   # fast path is n_qubits == 4
    assert lib.qw_conv1d_dp_flag_bytes(*dims, 2) == ((1920 + 64 + 4608) // 32 + 1) * 4
    P = ctypes.c_void_p
    tbl = (P * 2)(P(0), P(0))
    st = lib.qw_conv1d_backward_dp(*([None] * 12), None, 0, *dims, 0, tbl, tbl, 3, 2, ctypes.c_float(0.5), None)
    assert st == -1 and b"rank/world" in lib.qw_last_error()
    assert lib.qw_grads_allreduce_p2p_buffer_bytes(9440, 8) == 2 * 8 * 9440 * 8
    assert lib.qw_timeline_set(None, 0) == 0
