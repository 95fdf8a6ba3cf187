"""Loading the reference's checkpoints into the B200 modules.

The reference saves either a bare ``state_dict`` (``train_quantum_whisper.py:283-295,643``,
``train_quantum_whisper_asr.py:260-264,409``) or the dict written by ``utils.save_model`` (``utils.py:408-438``:
``{"model_state_dict": ..., "model_info": ..., "training_history": ...}``).  Its ``QuantumWhisper`` also carries the
Whisper text decoder (``decoder.*`` keys), which is downstream of the hot path and not built here, so a strict load on our
side means: every parameter / buffer of OUR module must be present with the right shape; reference-only keys are reported.
Parameter names and shapes of the quantum stem are identical by construction (``pre_conv.*``, ``post_conv.*``,
``quantum_weights``), so no renaming is involved.
"""
from __future__ import annotations

from typing import Dict, List, Mapping, Tuple

import torch


def extract_state_dict(checkpoint) -> Mapping[str, torch.Tensor]:
    """Accept a bare state_dict or the ``utils.save_model`` dict (``utils.py:422-431``)."""
    if isinstance(checkpoint, Mapping) and "model_state_dict" in checkpoint:
        return checkpoint["model_state_dict"]
    return checkpoint


def load_reference_state_dict(module: torch.nn.Module, checkpoint, strict: bool = True) -> Tuple[List[str], List[str]]:
    """Copy every tensor our module owns from a reference checkpoint.  Returns (missing, reference_only) key lists.
    With ``strict`` a missing or mis-shaped key raises, exactly like ``load_state_dict(strict=True)`` would for ours."""
    sd = extract_state_dict(checkpoint)
    own: Dict[str, torch.Tensor] = module.state_dict()
    missing = [k for k in own if k not in sd]
    reference_only = [k for k in sd if k not in own]
    bad = [f"{k}: checkpoint {tuple(sd[k].shape)} vs module {tuple(own[k].shape)}" for k in own
           if k in sd and tuple(sd[k].shape) != tuple(own[k].shape)]
    if strict and (missing or bad):
        raise RuntimeError("reference checkpoint does not fit this module: missing " + ", ".join(missing) +
                           ("; shape mismatch " + "; ".join(bad) if bad else ""))
    module.load_state_dict({k: v for k, v in sd.items() if k in own and k not in [b.split(":")[0] for b in bad]}, strict=False)
    return missing, reference_only


def load_reference_checkpoint(module: torch.nn.Module, path: str, device="cpu", strict: bool = True, allow_pickle: bool = False):
    """``utils.load_model`` counterpart (``utils.py:440-473``): returns (module, training_history, model_info).

    The reference format (``utils.save_model``) is a state_dict plus dicts of primitives and strings, so the safe
    ``weights_only=True`` loader is enough and is the default; ``allow_pickle=True`` opts into full unpickling for a legacy file
    from a TRUSTED source (unpickling executes arbitrary code)."""
    ckpt = torch.load(path, map_location=device, weights_only=not allow_pickle)
    load_reference_state_dict(module, ckpt, strict=strict)
    hist = ckpt.get("training_history", {}) if isinstance(ckpt, Mapping) else {}
    info = ckpt.get("model_info", {}) if isinstance(ckpt, Mapping) else {}
    return module, hist, info
