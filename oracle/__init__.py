"""CPU oracle for the Quantum-Whisper hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``qasr_ijcnlp_b200/`` imports this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it, and only as the checker (or as the CPU baseline being timed),
never as the product path.

Parity status: **parity unpinned at the PennyLane boundary** -- PennyLane is not
installed and cannot be installed here, and the reference ships no test or golden
vector for ``QuantumConv1d``.  The restatement is pinned
instead against (i) the known-answer vectors of SURVEY.md section 8c, (ii) a second,
independently written dense-unitary oracle (``qconv_oracle.dense_unitary_expvals``),
(iii) the algebraic identities of section 8c, and -- for the log-mel front end --
against the vendored ``whisper.log_mel_spectrogram`` run in the build container
(fixtures under ``tests/golden/``, made by ``tests/golden/make_golden.py``).
"""


def pennylane_reference():
    """if PennyLane ever becomes importable where the reference tree is present, the literal
    reference module IS the oracle.  Returns ``/root/reference/quantum_whisper.QuantumConv1d`` (the unmodified class) or
    ``None``; callers (tests/test_oracle_cpu.py, bench.py --impl reference) pin / time against it when it is there and say
    "parity unpinned" when it is not.  Never raises."""
    import importlib.util
    import os
    import sys

    ref = "/root/reference/quantum_whisper.py"
    if not os.path.isfile(ref):
        return None
    try:
        import pennylane  # noqa: F401
    except Exception:
        return None
    try:
        sys.path.insert(0, "/root/reference/whisper")
        spec = importlib.util.spec_from_file_location("_reference_quantum_whisper", ref)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod.QuantumConv1d
    except Exception:
        return None
    finally:
        if sys.path and sys.path[0] == "/root/reference/whisper":
            sys.path.pop(0)
