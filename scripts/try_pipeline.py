"""Run the full-size pipeline once or twice on the GPU box and print stage timings (development aid)."""
import json
import os
import sys
import time

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from importlib import import_module

import fo1_b200  # noqa

E = import_module("vlm-fo1_b200.engine"); CK = import_module("vlm-fo1_b200.checkpoint")
P = import_module("vlm-fo1_b200.pipeline"); SY = import_module("vlm-fo1_b200.synthetic")

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
S = int(sys.argv[2]) if len(sys.argv) > 2 else 896
N = int(sys.argv[3]) if len(sys.argv) > 3 else 64
T = int(sys.argv[4]) if len(sys.argv) > 4 else 64


def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e


t0 = time.time()
cfg = E.EngineConfig()
sds = CK.random_state_dicts(cfg, "cuda", 0)
eng = CK.load_engine(cfg, sds)
del sds
torch.cuda.synchronize()
print(f"weights ready in {time.time() - t0:.1f}s, mem {torch.cuda.memory_allocated() / 2**30:.1f} GiB", flush=True)
samples = SY.synthetic_batch(0, B, S, N)
for s in samples:
    s.pixel_values = s.pixel_values.cuda(); s.image_aux = s.image_aux.cuda(); s.boxes = s.boxes.cuda()
pipe = P.Fo1Pipeline(eng)
for it in range(3):
    torch.cuda.synchronize()
    a = ev()
    feats, taps = eng.vit_forward([s.pixel_values for s in samples], [s.grid_hw for s in samples]); b = ev()
    st = eng.davit_forward([s.image_aux for s in samples]); c = ev()
    gh, gw = samples[0].grid_hw
    pyr = eng.fpn_forward(taps[-1].view(B, gh, gw, 1280)); d = ev()
    out = pipe.generate(samples, T, stop_ids=[], early_exit_interval=0); e = ev()
    torch.cuda.synchronize()
    print(json.dumps({"iter": it, "B": B, "vit_ms": a.elapsed_time(b), "davit_ms": b.elapsed_time(c), "fpn_ms": c.elapsed_time(d),
                      "generate_all_ms": d.elapsed_time(e), "launches": fo1_b200.lib().fo1_launch_count(),
                      "mem_GiB": torch.cuda.max_memory_allocated() / 2**30}), flush=True)
print("tokens[0][:8]", out["tokens"][0][:8].tolist(), "lens", out["lens"][:4].tolist())
