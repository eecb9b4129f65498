"""ncu target: the CTA-pair GEMM at the ViT qkv shape of a C3 step (run under `ncu --set full -k regex:pair_kernel -s 1 -c 1`)."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from importlib import import_module
import fo1_b200  # noqa
ops = import_module("vlm-fo1_b200.ops")
a = (torch.randn(131072, 1280, device="cuda") * 0.5).bfloat16(); w = (torch.randn(3840, 1280, device="cuda") * 0.05).bfloat16()
out = torch.empty(131072, 3840, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.gemm(a, w, out=out)
torch.cuda.synchronize()
