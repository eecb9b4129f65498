"""compute-sanitizer --tool initcheck target: a reduced-depth pipeline (same kernels, fewer layers) on ragged images, batch then one by one."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from importlib import import_module
import fo1_b200  # noqa
E = import_module("vlm-fo1_b200.engine"); CK = import_module("vlm-fo1_b200.checkpoint"); P = import_module("vlm-fo1_b200.pipeline")
SY = import_module("vlm-fo1_b200.synthetic")
cfg = E.EngineConfig()
cfg.vit = dict(cfg.vit, depth=4, fullatt_block_indexes=[1, 3])
cfg.davit = dict(cfg.davit, depths=[1, 1, 1, 1])
cfg.llm = dict(cfg.llm, num_hidden_layers=2)
dev = torch.device("cuda", 0)
eng = CK.load_engine(cfg, CK.random_state_dicts(cfg, dev, 0), dev)
pipe = P.Fo1Pipeline(eng)
host = [SY.synthetic_sample(0, 448, 5), SY.synthetic_sample(1, 336, 3), SY.synthetic_sample(2, 560, 9)]
tb = pipe.generate(host, 6, stop_ids=[], early_exit_interval=0)["tokens"].cpu()
for i, s in enumerate(host):
    t1 = pipe.generate([s], 6, stop_ids=[], early_exit_interval=0)["tokens"].cpu()
    print(i, bool(torch.equal(t1[0], tb[i])), t1[0].tolist(), tb[i].tolist(), flush=True)
