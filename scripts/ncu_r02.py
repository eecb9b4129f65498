"""ncu target: one launch of every roofline kernel at its C3 shape (run under `ncu --set full -k regex:...`)."""
import os, sys
import numpy as np
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from importlib import import_module
import fo1_b200  # noqa
ops = import_module("vlm-fo1_b200.ops"); HF = import_module("vlm-fo1_b200.hfre"); SY = import_module("vlm-fo1_b200.synthetic")


def attn(lens, qh, kvh, hd, causal):
    T = sum(lens)
    qkv = torch.randn(T, (qh + 2 * kvh) * hd, device="cuda").to(torch.bfloat16)
    cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device="cuda")
    ops.attention_varlen(qkv[:, : qh * hd], qkv[:, qh * hd:(qh + kvh) * hd], qkv[:, (qh + kvh) * hd:], cu, max(lens), qh, kvh, hd, hd ** -0.5, causal=causal)


for rep in range(2):                       # first pass warms the code / tensor maps, the capture takes the second (-s)
    attn([4096] * 8, 16, 16, 80, False)
    attn([64] * (64 * 8), 16, 16, 80, False)
    attn([1195] * 8, 16, 2, 128, True)
    a = (torch.randn(32768, 1280, device="cuda") * 0.5).to(torch.bfloat16); w = (torch.randn(3840, 1280, device="cuda") * 0.05).to(torch.bfloat16)
    ops.gemm(a, w)
    ops.channel_attention(torch.randn(8, 50176, 768, device="cuda").to(torch.bfloat16), 8)
    S, N, B = 896, 64, 8
    g = torch.Generator().manual_seed(1)
    aux = [torch.randn(S // (4 << i), S // (4 << i), c, generator=g).to(torch.bfloat16).cuda() for i, c in enumerate((256, 512, 1024, 2048))]
    pyr = [torch.randn(int(64 * f), int(64 * f), 512, generator=g).to(torch.bfloat16).cuda() for f in (4, 2, 1, 0.5)]
    boxes = [SY.synthetic_boxes(b, S, N).cuda() for b in range(B)]
    HF.hfre_forward([aux] * B, [pyr] * B, boxes, boxes, HF.HfreConfig(region_dim=5888, vt_mode="fpn"), [(64, 64)] * B)
    torch.cuda.synchronize()
