"""Activation-precision switch of the oracle -- TEST INFRASTRUCTURE ONLY.

The restatements are fp32 by default (that is how they are pinned against the reference's modules, which were executed
on the CPU in fp32 to produce tests/golden/*.npz).  The reference's CUDA path, like the engine, keeps every tensor that
crosses a module boundary in bf16 (the checkpoint is loaded with torch_dtype=bfloat16, vlm_fo1/model/builder.py:44):
Linear / conv / norm / attention outputs, residual sums and activations are each rounded to bf16 when they are stored.
``R`` marks those storage points in oracle/{vit,davit,llm}.py; with ``act_bf16(True)`` it rounds, so that a GPU-vs-oracle
comparison sees only accumulation-order and fusion differences (one bf16 ulp = 2^-8 relative per storage point) instead
of the fp32-vs-bf16 drift compounded over depth.  Arithmetic inside an operator stays fp32 in both modes (the CUDA
kernels of both the reference and the engine accumulate in fp32)."""
from __future__ import annotations

import contextlib

import torch

_ACT_BF16 = False


def R(x: torch.Tensor) -> torch.Tensor:
    """Storage point of an activation: bf16 round trip when the switch is on, identity otherwise."""
    return x.to(torch.bfloat16).to(torch.float32) if _ACT_BF16 else x


@contextlib.contextmanager
def act_bf16(on: bool = True):
    global _ACT_BF16
    prev = _ACT_BF16
    _ACT_BF16 = bool(on)
    try:
        yield
    finally:
        _ACT_BF16 = prev
