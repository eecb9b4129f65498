"""-m gpu: the whole path (ViT -> DaViT -> FPN -> HFRE -> projector -> splice -> prefill -> decode) on the engine vs
the CPU oracle pipeline, real layer widths at reduced depth, same seed-0 random checkpoint, 2 samples of different
sizes in one batch (ragged grids, ragged box counts, one sample without boxes)."""
import numpy as np
import pytest
import torch

from tests.golden_io import nerr

pytestmark = pytest.mark.gpu


def test_pipeline_matches_oracle_end_to_end():
    from importlib import import_module
    import fo1_b200  # noqa: F401
    E = import_module("vlm-fo1_b200.engine"); CK = import_module("vlm-fo1_b200.checkpoint")
    P = import_module("vlm-fo1_b200.pipeline"); SY = import_module("vlm-fo1_b200.synthetic")
    from oracle import pipeline as OP
    cfg = E.EngineConfig()
    cfg.vit = dict(cfg.vit, depth=2, fullatt_block_indexes=[0, 1])
    cfg.davit = dict(cfg.davit, depths=[1, 1, 1, 1])
    cfg.llm = dict(cfg.llm, num_hidden_layers=2)
    sds = CK.random_state_dicts(cfg, "cuda", 0)
    eng = CK.load_engine(cfg, sds)
    sds_cpu = {k: {n: t.float().cpu() for n, t in v.items()} for k, v in sds.items()}
    samples = [SY.synthetic_sample(0, 224, 5), SY.synthetic_sample(1, 280, 9), SY.synthetic_sample(2, 224, 0)]
    samples[2].input_ids = [t for t in samples[2].input_ids]          # no region placeholders: the dummy box is unused by the prompt
    pipe = P.Fo1Pipeline(eng)
    dev_samples = [P.SampleInputs(s.input_ids, s.pixel_values.cuda(), s.grid_hw, s.image_aux.cuda(), s.boxes.cuda()) for s in samples]
    feats, img_off, region_tokens, region_f32 = pipe.encode(dev_samples)
    out = pipe.generate(dev_samples, 4, stop_ids=[], early_exit_interval=0, want_prefill_logits=True)
    torch.cuda.synchronize()
    for b, s in enumerate(samples):
        ref = OP.run_sample(sds_cpu, cfg.vit, cfg.davit, cfg.llm, input_ids=s.input_ids, pixel_values=s.pixel_values, grid_hw=s.grid_hw,
                            image_aux=s.image_aux, boxes=s.boxes, region_dim=cfg.region_dim, max_new_tokens=4)
        # region features: fp32 pooling of bf16 tower maps that themselves carry bf16 rounding noise from 2 blocks
        assert nerr(region_f32[b].cpu(), ref["region_features"]) < 2e-2, b
        assert nerr(region_tokens[b].float().cpu(), ref["region_tokens"]) < 3e-2, b
        assert nerr(feats[img_off[b]:img_off[b + 1]].float().cpu(), ref["image_features"]) < 2e-2, b
        assert out["prompt_lens"][b] == ref["prompt_len"]
        assert nerr(out["prefill_logits"][b].cpu(), ref["prompt_last_logits"]) < 3e-2, b
