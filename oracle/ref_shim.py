"""Import the *reference's own* modules in the build container -- TEST INFRASTRUCTURE ONLY.

Used only by ``oracle/gen_golden.py`` (fixture generation, run where ``/root/reference``
exists).  Nothing that runs on the GPU box imports this file.  The reference pins
transformers 4.50.1 / timm 1.0.9; the container has transformers 5.5 and no timm, so three
shims are installed before import:

1. ``transformers.cache_utils.SlidingWindowCache`` (removed in 5.x) -> subclass of StaticCache;
2. ``ROPE_INIT_FUNCTIONS["default"]`` -> the standard inverse-frequency initialiser;
3. stub ``timm`` modules (``DropPath`` -> Identity, ``trunc_normal_`` -> torch's).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("FO1_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "vlm_fo1"))


def _install_timm_stub() -> None:
    if "timm" in sys.modules:
        return
    import torch
    import torch.nn as nn

    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")
    layers2 = types.ModuleType("timm.layers")
    regnet = types.ModuleType("timm.models.regnet")

    class DropPath(nn.Identity):
        def __init__(self, *a, **k):
            super().__init__()

    for m in (layers, layers2):
        m.DropPath = DropPath
        m.trunc_normal_ = torch.nn.init.trunc_normal_
        m.LayerNorm = nn.LayerNorm
        m.LayerNorm2d = nn.LayerNorm
    regnet.RegStage = type("RegStage", (nn.Module,), {})
    timm.models = models
    models.layers = layers
    models.regnet = regnet
    timm.layers = layers2
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers,
                        "timm.layers": layers2, "timm.models.regnet": regnet})


def _install_transformers_shims() -> None:
    import torch
    import transformers.cache_utils as cu
    if not hasattr(cu, "SlidingWindowCache"):
        cu.SlidingWindowCache = type("SlidingWindowCache", (cu.StaticCache,), {})
    from transformers.modeling_rope_utils import ROPE_INIT_FUNCTIONS
    if "default" not in ROPE_INIT_FUNCTIONS:
        def _default(config, device=None, seq_len=None, **kw):
            base = config.rope_theta
            dim = getattr(config, "head_dim", None) or config.hidden_size // config.num_attention_heads
            inv = 1.0 / (base ** (torch.arange(0, dim, 2, dtype=torch.int64).float().to(device) / dim))
            return inv, 1.0
        ROPE_INIT_FUNCTIONS["default"] = _default


def load_hfre_only():
    """HFREModule + SimpleFP need only torch/torchvision: pre-seed empty package stubs so the
    reference's heavy ``vlm_fo1/model/__init__`` chain is never executed."""
    if not reference_available():
        raise RuntimeError("reference tree not present")
    for name in ("vlm_fo1", "vlm_fo1.model", "vlm_fo1.model.multimodal_visual_prompt_encoder"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(REFERENCE_ROOT, *name.split("."))]
            sys.modules[name] = m
    hf = importlib.import_module(
        "vlm_fo1.model.multimodal_visual_prompt_encoder.hybrid_finegrained_region_encoder")
    fpn = importlib.import_module("vlm_fo1.model.multimodal_visual_prompt_encoder.simple_fpn")
    return hf, fpn


def load_reference_package():
    """Full ``import vlm_fo1.model`` with the three shims (ViT, DaViT, projectors, mm_utils)."""
    if not reference_available():
        raise RuntimeError("reference tree not present")
    for k in [k for k in sys.modules if k == "vlm_fo1" or k.startswith("vlm_fo1.")]:
        mod = sys.modules[k]
        if getattr(mod, "__file__", None) is None:  # drop the light stubs from load_hfre_only()
            del sys.modules[k]
    _install_transformers_shims()   # imports transformers first (its lazy init probes for a real timm)
    _install_timm_stub()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return importlib.import_module("vlm_fo1.model")
