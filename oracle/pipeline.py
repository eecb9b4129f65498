"""CPU reference arm of the whole hot path for ONE sample -- TEST INFRASTRUCTURE ONLY.

Composes the per-stage oracles in the order of OmChatQwen25VLForCausalLM.forward
(vlm_fo1/model/language_model/omchat_qwen2_5_vl.py:466-532 -> :135-463 -> :44-128): ViT (+taps) -> DaViT ->
SimpleFPN on the last tap -> HFRE -> region projector -> splice + M-RoPE -> decoder prefill + greedy decode.
Activations are rounded to bf16 exactly where the reference stores bf16 tensors between its towers
(tower outputs, region features :106, projector output), everything else is fp32.
Used by bench.py's cpu_baseline / ``--impl reference`` legs (timed on the host cores) and by smoke()."""
from __future__ import annotations

import time
from typing import Dict, Sequence

import torch

try:  # imported as the package ``oracle`` (tests, bench) ...
    from . import davit as OD, hfre as OH, llm as OL, vit as OV
except ImportError:  # ... or from inside the directory (gen_golden.py)
    import davit as OD, hfre as OH, llm as OL, vit as OV  # type: ignore


def _bf(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


def run_sample(sds: Dict[str, Dict[str, torch.Tensor]], vit_cfg: dict, davit_cfg: dict, llm_cfg: dict, *, input_ids: Sequence[int],
               pixel_values: torch.Tensor, grid_hw, image_aux: torch.Tensor, boxes: torch.Tensor, region_dim: int,
               max_new_tokens: int, stop_ids: Sequence[int] = (), vt_mode: str = "fpn", image_token_id: int = 151655,
               vision_start_token_id: int = 151652, stages_only: bool = False, forced: Sequence[int] = None,
               keep_stages: bool = False, round_towers: bool = True) -> dict:
    """``forced``: teacher forcing -- feed these tokens instead of the argmax (the per-step logits are still returned).
    ``keep_stages``: also return the taps, DaViT stage maps and FPN levels.  ``round_towers=False`` keeps even the tensors the
    reference stores in bf16 between its towers in fp32 (the all-fp32 yardstick of the parity tests)."""
    _bf = (lambda x: x.to(torch.bfloat16).to(torch.float32)) if round_towers else (lambda x: x)
    t = {}
    gh, gw = grid_hw
    t0 = time.perf_counter()
    t["vit_detail"] = {}
    merged, taps = OV.vit_forward(sds["vit"], vit_cfg, pixel_values, gh, gw, timing=t["vit_detail"])
    merged = _bf(merged); taps = [_bf(x) for x in taps]
    t["vit_s"] = time.perf_counter() - t0; t0 = time.perf_counter()
    full_cfg = dict(davit_cfg)
    full_cfg.setdefault("patch_prenorm", [False, True, True, True]); full_cfg.setdefault("patch_stride", [4, 2, 2, 2])
    full_cfg.setdefault("patch_padding", [3, 1, 1, 1])
    t["davit_detail"] = {}
    aux = [_bf(x) for x in OD.davit_forward(sds["davit"], full_cfg, image_aux, timing=t["davit_detail"])]
    t["davit_s"] = time.perf_counter() - t0; t0 = time.perf_counter()
    if vt_mode == "fpn":
        vt = [_bf(x) for x in OD.fpn_forward(sds["fpn"], taps[-1])]
    else:
        vt = taps
    t["fpn_s"] = time.perf_counter() - t0; t0 = time.perf_counter()
    Ha, Wa = image_aux.shape[-2:]
    b = boxes.float() if boxes.numel() else torch.tensor([[0.0, 10.0, 0.0, 10.0]])
    p = vit_cfg["patch_size"]
    vt_boxes = b * torch.tensor([gw * p / Wa, gh * p / Ha, gw * p / Wa, gh * p / Ha])
    region = OH.hfre_forward([a.permute(2, 0, 1) for a in aux], b, [v.permute(2, 0, 1) for v in vt], vt_boxes, vt_mode=vt_mode,
                             region_dim=region_dim, vt_grid_hw=(gh, gw))
    t["hfre_s"] = time.perf_counter() - t0; t0 = time.perf_counter()
    region_tokens = _bf(OD.projector_forward(sds["proj_aux"], _bf(region)))
    t["proj_s"] = time.perf_counter() - t0
    out = dict(region_features=region, region_tokens=region_tokens, image_features=merged, timings=t)
    if keep_stages:
        out.update(taps=taps, davit=aux, fpn=vt if vt_mode == "fpn" else None)
    if stages_only:
        return out
    t0 = time.perf_counter()
    new_ids, kind, index = OL.splice_plan(input_ids, [grid_hw], region_tokens.shape[0], image_token_id, vit_cfg["spatial_merge_size"])
    pos, delta = OL.rope_index(new_ids, [grid_hw], image_token_id, vision_start_token_id, vit_cfg["spatial_merge_size"])
    emb = sds["llm"]["embed_tokens.weight"].float()
    rows = []
    for k, ix in zip(kind, index):
        rows.append(emb[ix] if k == 0 else (merged[ix] if k == 1 else region_tokens[ix]))
    embeds = torch.stack(rows)
    # prefill and decode timed separately (the decode cost per token is what a bounded sample extrapolates from)
    dec = OL.Decoder(sds["llm"], llm_cfg)
    tp = time.perf_counter()
    h = dec.forward(embeds.float(), pos)
    t["llm_prefill_layer_s"] = list(dec.layer_seconds)
    th = time.perf_counter()
    prompt_logits = dec.logits(h[-1:])
    t["llm_head_s"] = time.perf_counter() - th
    t["llm_prefill_s"] = time.perf_counter() - tp
    lg = prompt_logits[-1]
    toks, lgs = [], []
    L = embeds.shape[0]
    td = time.perf_counter()
    for s_ in range(max_new_tokens):
        lgs.append(lg)
        tok = int(lg.argmax()) if forced is None else int(forced[s_])
        toks.append(tok)
        if (forced is None and tok in stop_ids) or s_ == max_new_tokens - 1:
            break
        h = dec.forward(emb[tok][None, :], torch.full((3, 1), L + s_ + delta, dtype=torch.long))
        t.setdefault("llm_decode_layer_s", []).append(sum(dec.layer_seconds) / max(len(dec.layer_seconds), 1))
        lg = dec.logits(h)[0]
    t["llm_decode_s_per_token"] = (time.perf_counter() - td) / max(len(toks) - 1, 1)
    step_logits = torch.stack(lgs)
    out.update(tokens=toks, step_logits=step_logits, prompt_last_logits=prompt_logits[-1], position_ids=pos, rope_delta=delta,
               prompt_len=len(new_ids))
    return out
