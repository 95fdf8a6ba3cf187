"""qasr_ijcnlp_b200 -- B200-native hot path of Quantum Whisper (Debjit-Dhar/QASR_IJCNLP).

Only what the path needs: the CUDA kernels + C ABI (`csrc/`, `libqw_b200.so`), the drop-in `QuantumConv1d`
module, the log-mel front end, and the encoder classes that call them.  See DESIGN.md.
"""
from .quantum_conv1d import (  # noqa: F401
    QuantumConv1d, quantum_conv1d, quantum_circuit, out_length, fused_stem_forward, fused_stem_eligible,
    stem_train_forward, stem_train_eligible,
)
from .encoder import (  # noqa: F401
    ModelDimensions, AudioEncoder, QuantumAudioEncoder, QuantumWhisper, QuantumWhisperClassifier, QuantumWhisperASR,
    CharASRHead, CHAR_VOCAB, get_whisper_tiny_dims, freeze_non_quantum_layers,
)

__all__ = [
    "QuantumConv1d", "quantum_conv1d", "quantum_circuit", "out_length", "fused_stem_forward", "fused_stem_eligible", "stem_train_forward", "stem_train_eligible", "ModelDimensions", "AudioEncoder",
    "QuantumAudioEncoder", "QuantumWhisper", "QuantumWhisperClassifier", "QuantumWhisperASR", "CharASRHead",
    "CHAR_VOCAB", "get_whisper_tiny_dims", "freeze_non_quantum_layers",
]
