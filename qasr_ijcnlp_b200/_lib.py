"""ctypes binding of libqw_b200.so (the C ABI declared in include/qw.h).

There is NO CPU fallback: if the shared library is missing it is built in-tree with nvcc, and if that is
impossible the import of the op raises.  Every function returns an int status that `check()` turns into a
Python exception carrying `qw_last_error()`.
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libqw_b200.so")

_lock = threading.Lock()
_lib = None

_P = ctypes.c_void_p
_I = ctypes.c_int
_LL = ctypes.c_longlong
_SZ = ctypes.c_size_t

_CONV_DIMS = [_I] * 10  # B C L K S P O q n_layers embedding

_SIGNATURES = {
    "qw_abi_version": (_I, []),
    "qw_build_stamp": (ctypes.c_char_p, []),
    "qw_last_error": (ctypes.c_char_p, []),
    "qw_launch_count": (_LL, []),
    "qw_profile_enable": (None, [_I]),
    "qw_profile_read": (_I, [_I, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_LL), _I]),
    "qw_kernel_name": (ctypes.c_char_p, [_I]),
    "qw_kernel_symbol": (ctypes.c_char_p, [_I]),
    "qw_timeline_set": (_I, [_P, _I]),
    "qw_set_fast_path": (None, [_I]),
    "qw_set_option": (_I, [ctypes.c_char_p, _I]),
    "qw_get_option": (_I, [ctypes.c_char_p]),
    "qw_conv1d_forward": (_I, [_P] * 8 + _CONV_DIMS + [_P]),
    "qw_conv1d_forward_f64": (_I, [_P] * 8 + _CONV_DIMS + [_P]),
    "qw_conv1d_workspace_bytes": (_SZ, [_I] * 10),
    "qw_conv1d_backward": (_I, [_P] * 12 + [_P, _SZ] + _CONV_DIMS + [_P]),
    "qw_conv1d_backward_f64": (_I, [_P] * 12 + [_P, _SZ] + _CONV_DIMS + [_P]),
    "qw_conv1d_forward_act": (_I, [_P] * 8 + _CONV_DIMS + [_I, _P]),
    "qw_conv1d_backward_act": (_I, [_P] * 13 + [_P, _SZ] + _CONV_DIMS + [_I, _P]),
    "qw_conv1d_dp_buffer_bytes": (_SZ, [_I] * 10),
    "qw_conv1d_dp_flag_bytes": (_SZ, [_I] * 10),
    "qw_conv1d_backward_dp": (_I, [_P] * 12 + [_P, _SZ] + _CONV_DIMS + [ctypes.POINTER(_P), ctypes.POINTER(_P), _I, _I, ctypes.c_float, _P]),
    "qw_stem_workspace_bytes": (_SZ, [_I, _I]),
    "qw_stem_forward": (_I, [_P] * 13 + [_P, _SZ] + [_I] * 6 + [_P]),
    "qw_stem_train_forward": (_I, [_P] * 15 + [_I] * 7 + [_P]),
    "qw_stem_train_forward_preferred": (_I, [_I, _I]),
    "qw_conv1d_backward_chained": (_I, [_P, _P, _I] + [_P] * 13 + [_SZ] + _CONV_DIMS + [_I] + [ctypes.POINTER(_P), ctypes.POINTER(_P), _I, _I, ctypes.c_float, _P]),
    "qw_circuit_workspace_bytes": (_SZ, [_LL, _I, _I, _I]),
    "qw_circuit_forward": (_I, [_P, _P, _P, _LL, _I, _I, _I, _P]),
    "qw_circuit_forward_f64": (_I, [_P, _P, _P, _LL, _I, _I, _I, _P]),
    "qw_circuit_backward": (_I, [_P] * 5 + [_P, _SZ, _LL, _I, _I, _I, _P]),
    "qw_circuit_backward_f64": (_I, [_P] * 5 + [_P, _SZ, _LL, _I, _I, _I, _P]),
    "qw_circuit_collapsed_workspace_bytes": (_SZ, [_LL, _I, _I]),
    "qw_circuit_forward_collapsed": (_I, [_P, _P, _P, _LL, _I, _P]),
    "qw_circuit_forward_collapsed_f64": (_I, [_P, _P, _P, _LL, _I, _P]),
    "qw_circuit_backward_collapsed": (_I, [_P] * 5 + [_P, _SZ, _LL, _I, _P]),
    "qw_circuit_backward_collapsed_f64": (_I, [_P] * 5 + [_P, _SZ, _LL, _I, _P]),
    "qw_log_mel_workspace_bytes": (_SZ, [_I, _I, _I]),
    "qw_log_mel": (_I, [_P, _P, _P, _P, _SZ, _I, _I, _I, _P]),
    "qw_log_mel_prep_bytes": (_SZ, [_I]),
    "qw_log_mel_call_workspace_bytes": (_SZ, [_I, _I]),
    "qw_log_mel_prepare": (_I, [_P, _I, _P, _SZ, _P]),
    "qw_log_mel_prepared": (_I, [_P, _P, _P, _P, _SZ, _I, _I, _I, _P]),
    "qw_log_mel_padded": (_I, [_P, _P, _P, _P, _P, _SZ, _I, _I, _I, _I, _P]),
    "qw_grads_allreduce_p2p_buffer_bytes": (_SZ, [_LL, _I]),
    "qw_grads_allreduce_p2p_flag_bytes": (_SZ, [_I]),
    "qw_grads_allreduce_p2p": (_I, [_P, _LL, ctypes.POINTER(_P), ctypes.POINTER(_P), _I, _I, ctypes.c_float, _P]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


class QwError(RuntimeError):
    pass


def load(build_if_missing: bool = True):
    """Load (building first if needed) libqw_b200.so.  Raises if it cannot be had."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            if not build_if_missing:
                raise QwError(f"{LIB_PATH} is missing; run `python -m qasr_ijcnlp_b200.build`")
            from . import build as _build

            _build.build()
        lib = ctypes.CDLL(LIB_PATH)
        lib = _refresh_if_stale(lib, build_if_missing)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        if lib.qw_abi_version() != 1:
            raise QwError("libqw_b200.so ABI version mismatch; rebuild with `python -m qasr_ijcnlp_b200.build --force`")
        _lib = lib
        return lib


def _refresh_if_stale(lib, may_build: bool):
    """The .so is git-ignored but travels with the tree: make sure it was built from THESE sources.  A library
    whose embedded source hash differs from the tree is rebuilt in-tree; if that is impossible the load fails loudly."""
    from . import build as _build

    want = _build._source_hash()
    try:
        fn = lib.qw_build_stamp
        fn.restype = ctypes.c_char_p
        have = fn().decode()
    except AttributeError:
        have = "missing"
    if have == want or os.environ.get("QW_ALLOW_STALE_LIB") == "1":
        return lib
    if not may_build:
        raise QwError(f"{LIB_PATH} was built from other sources (stamp {have}, tree {want}); run `python -m qasr_ijcnlp_b200.build`")
    # dlclose is not reliable through ctypes: build under a fresh name is not needed -- the process has not called into the old
    # mapping yet, so rebuild in place and map the new file
    handle = lib._handle
    del lib
    try:
        import _ctypes

        _ctypes.dlclose(handle)
    except Exception:
        pass
    _build.build(force=False)
    lib = ctypes.CDLL(LIB_PATH)
    fn = lib.qw_build_stamp
    fn.restype = ctypes.c_char_p
    if fn().decode() != want:
        raise QwError(f"{LIB_PATH} is stale (stamp {fn().decode()}, tree {want}) and could not be rebuilt")
    return lib


def check(status: int, what: str) -> None:
    if status == 0:
        return
    msg = load().qw_last_error().decode("utf-8", "replace")
    kind = "bad argument" if status < 0 else f"cudaError {status}"
    raise QwError(f"{what} failed ({kind}, status {status}): {msg}")


def launch_count() -> int:
    return int(load().qw_launch_count())


def profile_enable(on: bool) -> None:
    load().qw_profile_enable(1 if on else 0)


def profile_read(reset: bool = True) -> dict:
    """{kernel name: (total_ms, launches)} for every kernel timed since the last reset."""
    lib = load()
    out = {}
    kid = 0
    while True:
        name = lib.qw_kernel_name(kid)
        if not name:
            break
        ms, n = ctypes.c_double(0.0), _LL(0)
        lib.qw_profile_read(kid, ctypes.byref(ms), ctypes.byref(n), 1 if reset else 0)
        if n.value:
            out[name.decode()] = (ms.value, int(n.value))
        kid += 1
    return out


def kernel_symbols() -> dict:
    """{kernel name: symbol last launched under it} -- the names an ncu launch list shows."""
    lib = load()
    out = {}
    kid = 0
    while True:
        name = lib.qw_kernel_name(kid)
        if not name:
            break
        sym = lib.qw_kernel_symbol(kid)
        if sym:
            out[name.decode()] = sym.decode()
        kid += 1
    return out
