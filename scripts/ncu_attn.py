"""ncu driver: the ViT full-attention shape (8 images x 4096 tokens, 16 heads x 80) and the LLM causal prefill shape."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from importlib import import_module
import fo1_b200  # noqa
ops = import_module("vlm-fo1_b200.ops")
def run(B, S, qh, kvh, hd, causal):
    T = B * S
    q = torch.randn(T, qh * hd, device="cuda").to(torch.bfloat16); k = torch.randn(T, kvh * hd, device="cuda").to(torch.bfloat16); v = torch.randn(T, kvh * hd, device="cuda").to(torch.bfloat16)
    cu = torch.arange(0, T + 1, S, dtype=torch.int32, device="cuda")
    for _ in range(2):
        ops.attention_varlen(q, k, v, cu, S, qh, kvh, hd, hd ** -0.5, causal=causal)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); ops.attention_varlen(q, k, v, cu, S, qh, kvh, hd, hd ** -0.5, causal=causal); e1.record(); torch.cuda.synchronize()
    fl = 4.0 * B * S * S * qh * hd * (0.5 if causal else 1.0)
    print({"B": B, "S": S, "hd": hd, "causal": causal, "ms": e0.elapsed_time(e1), "tflops": fl / e0.elapsed_time(e1) / 1e9}, flush=True)
run(8, 4096, 16, 16, 80, False)
run(8, 1195, 16, 2, 128, True)
