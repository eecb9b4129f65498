"""Run each pinned tile_n=192 configuration in its own process with a timeout (diagnostic)."""
import subprocess, sys, json
CASES = [("plain", 32, 1), ("plain", 130, 1), ("plain", 32, 3), ("plain", 130, 3), ("gated", 32, 1), ("gated", 32, 2)]
CHILD = r'''
import sys, torch
sys.path.insert(0, ".")
from importlib import import_module
import fo1_b200
ops = import_module("vlm-fo1_b200.ops")
kind, M, ks = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
g = torch.Generator(device="cuda").manual_seed(1)
if kind == "plain":
    N, K = 1000, 2112
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(torch.bfloat16); w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    for dt in (torch.float32, torch.bfloat16):
        out = ops.gemm(a, w, out_dtype=dt, tile_n=192, split_k=ks); torch.cuda.synchronize()
        ref = a.float() @ w.float().t()
        print(kind, M, ks, dt, "err", ((out.float() - ref).abs().max() / ref.abs().max()).item(), flush=True)
else:
    K, I = 2048, 1000
    x = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    wg = (torch.randn(I, K, device="cuda", generator=g) * 0.03).to(torch.bfloat16); wu = (torch.randn(I, K, device="cuda", generator=g) * 0.03).to(torch.bfloat16)
    w = ops.interleave_gate_up(wg, wu)
    out = ops.gemm(x, w, act="silu", gated=True, out_dtype=torch.float32, tile_n=192, split_k=ks); torch.cuda.synchronize()
    ref = torch.nn.functional.silu(x.float() @ wg.float().t()) * (x.float() @ wu.float().t())
    print(kind, M, ks, "err", ((out[:, :I] - ref).abs().max() / ref.abs().max()).item(), flush=True)
'''
for kind, M, ks in CASES:
    try:
        r = subprocess.run([sys.executable, "-c", CHILD, kind, str(M), str(ks)], capture_output=True, text=True, timeout=40)
        print(kind, M, ks, "rc", r.returncode, r.stdout.strip().replace("\n", " | ")[-300:], r.stderr.strip()[-200:], flush=True)
    except subprocess.TimeoutExpired:
        print(kind, M, ks, "TIMEOUT", flush=True)
