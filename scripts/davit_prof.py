"""python scripts/davit_prof.py [B] [px]: DaViT + SimpleFPN alone at the bench shapes (for an ncu launch list); prints the warm device time."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from importlib import import_module
import fo1_b200  # noqa
E = import_module("vlm-fo1_b200.engine"); CK = import_module("vlm-fo1_b200.checkpoint"); W = import_module("vlm-fo1_b200.weights")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
px = int(sys.argv[2]) if len(sys.argv) > 2 else 768
cfg = E.EngineConfig(); cfg.use_vit = False; cfg.use_llm = False; cfg.proj_aux_layers = 0; cfg.fpn_out = 0
g = torch.Generator(device="cuda").manual_seed(5)
eng = E.Engine(cfg)
eng.set_weights(W.prepare_davit(CK.random_davit(cfg.davit, g, "cuda"), cfg.davit, eng.device))
eng.finalize()
imgs = [torch.randn(3, px, px, device="cuda") for _ in range(B)]
def run():
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); eng.davit_forward(imgs); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
if os.environ.get("DAVIT_PROF_ONCE"):
    run()
else:
    run(); run()
    print("davit ms", min(run() for _ in range(5)))
