"""python scripts/determinism_probe.py [B] [size] [reps]: run the pipeline repeatedly on the same inputs and report which stage's output
changes from run to run (every kernel on the path is meant to be bit-reproducible)."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from importlib import import_module
import fo1_b200  # noqa
E = import_module("vlm-fo1_b200.engine"); CK = import_module("vlm-fo1_b200.checkpoint"); P = import_module("vlm-fo1_b200.pipeline")
SY = import_module("vlm-fo1_b200.synthetic")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 5
S = int(sys.argv[2]) if len(sys.argv) > 2 else 644
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
cfg = E.EngineConfig()
dev = torch.device("cuda", 0)
eng = CK.load_engine(cfg, CK.random_state_dicts(cfg, dev, 0), dev)
pipe = P.Fo1Pipeline(eng)
host = SY.synthetic_batch(0, B, S, 16)
res = [s.to(dev) if hasattr(s, "to") else s for s in host]


def flat(st):
    out = {"image_features": st["image_features"], "region_f32": torch.cat(st["region_f32"]), "region_tokens": torch.cat(st["region_tokens"])}
    for i, t in enumerate(st["taps"]): out[f"tap{i}"] = t
    for l in range(4):
        out[f"davit{l}"] = torch.stack(st["davit"][l]); out[f"fpn{l}"] = torch.stack(st["fpn"][l])
    return {k: v.clone() for k, v in out.items()}


ref = flat(pipe.encode_stages(host))
tok_ref = pipe.generate(host, 24, stop_ids=[], early_exit_interval=0)["tokens"].clone()
bad = {}
for r in range(reps):
    cur = flat(pipe.encode_stages(host))
    for k in ref:
        if not torch.equal(ref[k], cur[k]):
            bad[k] = bad.get(k, 0) + 1
    tok = pipe.generate(host, 24, stop_ids=[], early_exit_interval=0)["tokens"]
    if not torch.equal(tok, tok_ref):
        bad["tokens"] = bad.get("tokens", 0) + 1
    for tag in ("FO1_NO_MEGA",):
        os.environ[tag] = "1"
        t2 = pipe.generate(host, 24, stop_ids=[], early_exit_interval=0)["tokens"]
        del os.environ[tag]
        if r == 0: tok_nm = t2.clone()
        elif not torch.equal(t2, tok_nm): bad["tokens_no_mega"] = bad.get("tokens_no_mega", 0) + 1
torch.cuda.synchronize()
print("stages that changed between identical runs (count of", reps, "):", bad if bad else "none")
