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
