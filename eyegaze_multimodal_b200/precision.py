"""Compute-precision policy of the drop-in modules.

``fp32`` : activations and GEMM operands in fp32 (FFMA GEMM, fp32 attention) -- the parity mode whose logits
           match the reference CPU model to <= 1e-4.
``bf16`` : activations in bf16, tensor-core GEMMs with fp32 accumulation -- the throughput mode (<= 2e-2 relative).
``auto`` (default): bf16 inside ``torch.autocast('cuda', ...)`` (the reference's training loops enable autocast,
           train_multimodal_fuzzy_fusion.py:436), fp32 otherwise.
"""
import contextlib
import os

import torch

from . import _lib as L

_state = {"mode": os.environ.get("EGB_PRECISION", "auto")}


def set_precision(mode: str) -> None:
    if mode not in ("auto", "fp32", "bf16"):
        raise ValueError("precision must be 'auto', 'fp32' or 'bf16'")
    _state["mode"] = mode


def get_precision() -> str:
    m = _state["mode"]
    if m == "auto":
        return "bf16" if torch.is_autocast_enabled() else "fp32"
    return m


def compute_code() -> int:
    return L.BF16 if get_precision() == "bf16" else L.F32


@contextlib.contextmanager
def precision(mode: str):
    old = _state["mode"]
    set_precision(mode)
    try:
        yield
    finally:
        _state["mode"] = old
