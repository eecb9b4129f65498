"""ncu target: the CTA-pair GEMM at a small-K, output-bound shape (DaViT stage 0 qkv: 294912 x 768 x 256)."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from importlib import import_module
import fo1_b200  # noqa
ops = import_module("vlm-fo1_b200.ops")
a = (torch.randn(294912, 256, device="cuda") * 0.5).bfloat16(); w = (torch.randn(768, 256, device="cuda") * 0.05).bfloat16()
bias = torch.randn(768, device="cuda").bfloat16()
out = torch.empty(294912, 768, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.gemm(a, w, bias=bias, out=out)
torch.cuda.synchronize()
