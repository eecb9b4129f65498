"""Host-side mirror of the reference's multimodal forward for a BATCH of samples, over the engine:

    encode_images   (omchat_qwen2_5_vl.py:44-72)   -> Engine.vit_forward      (one packed launch sequence)
    encode_regions  (omchat_qwen2_5_vl.py:75-128)  -> Engine.davit_forward + fpn_forward + hfre_forward + region_project
    splice + M-RoPE (omchat_qwen2_5_vl.py:291-463) -> splice_plan (host ints) + Engine.build_embeds
    generate        (HF greedy loop)               -> Engine.generate (prefill + device-resident decode)

One image per sample (what every caller in the reference does).  torch is the memory / stream plumbing."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import hfre as HF
from .engine import Engine, splice_plan, splice_plan_batch


@dataclass
class SampleInputs:
    input_ids: Sequence[int]            # prompt ids with -200 (image) / -300 (region) placeholders (mm_utils.py:83-135)
    pixel_values: Optional[torch.Tensor]  # fp32 [gh*gw, 1176]  (Qwen2VLImageProcessor); None with image_u8
    grid_hw: Optional[Tuple[int, int]]
    image_aux: Optional[torch.Tensor]   # fp32 [3, H, W]      (CLIPImageProcessor, 'dynamic' mode); None with image_u8
    boxes: torch.Tensor                 # fp32 [N, 4] xyxy in aux-tensor pixels
    image_u8: Optional[torch.Tensor] = None   # uint8 [H, W, 3] RGB: pre-processed on the device by Fo1Pipeline.preprocess()


class Fo1Pipeline:
    def __init__(self, engine: Engine, vt_mode: str = "fpn", image_token_id: int = 151655, vision_start_token_id: int = 151652,
                 video_token_id: int = 151656, roi_size: int = 7, apply_pos_embed: bool = True):
        self.eng = engine
        self.vt_mode = vt_mode
        self.ids = dict(image_token_id=image_token_id, vision_start_token_id=vision_start_token_id, video_token_id=video_token_id)
        self.hcfg = HF.HfreConfig(region_dim=engine.cfg.region_dim, vt_mode=vt_mode, roi_size=roi_size, apply_pos_embed=apply_pos_embed)
        self.ws = HF.HfreWorkspace()
        self.profile_stages = False          # when set, CUDA events bracket every stage (read with stage_ms())
        self._marks = []
        self._keep_stages = None
        self._pre = None
        self.check_plan = True

    def _mark(self, name: str) -> None:
        if self.profile_stages:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self._marks.append((name, e))

    def stage_ms(self) -> Dict[str, float]:
        """Milliseconds between consecutive stage marks of the last call (synchronises)."""
        torch.cuda.synchronize()
        out: Dict[str, float] = {}
        for (n0, e0), (n1, e1) in zip(self._marks[:-1], self._marks[1:]):
            out[n1] = out.get(n1, 0.0) + e0.elapsed_time(e1)
        return out

    # ---- device-side pre-processing -----------------------------------------------
    def preprocess(self, samples: Sequence[SampleInputs]) -> List[SampleInputs]:
        """Samples that carry a uint8 image instead of the processors' tensors get them from the GPU (both towers;
        the aux tower in 'dynamic' mode, as the released checkpoint is configured)."""
        if all(s.image_u8 is None for s in samples):
            return list(samples)
        if self._pre is None:
            from .preprocess import DevicePreprocessor
            v = self.eng.cfg.vit
            self._pre = DevicePreprocessor(self.eng.device, v["patch_size"], v["spatial_merge_size"], v["temporal_patch_size"])
        out = []
        for s in samples:
            if s.image_u8 is None:
                out.append(s); continue
            img = s.image_u8.to(self.eng.device, non_blocking=True)
            px, grid = self._pre.primary(img)
            out.append(SampleInputs(s.input_ids, px, grid, self._pre.aux(img, 0), s.boxes))
        return out

    # ---- vision side -------------------------------------------------------------------------------------
    def encode(self, samples: Sequence[SampleInputs]):
        """-> (img_feats [sum merged tokens, hidden] bf16, per-sample row offsets, region tokens list of [N_b, hidden] bf16,
        region features fp32 list)."""
        eng, dev = self.eng, self.eng.device
        self._marks = []
        self._mark("start")
        samples = self.preprocess(samples)
        if any(s.image_u8 is not None for s in samples) or self._pre is not None:
            self._mark("preprocess")
        grids = [s.grid_hw for s in samples]
        feats, taps = eng.vit_forward([s.pixel_values for s in samples], grids)
        feats = eng.image_project(feats)                                             # mm_projector (:58-66); identity for the released checkpoint
        self._mark("vit")
        B = len(samples)
        H0, W0 = samples[0].image_aux.shape[-2:]
        same_aux = all(s.image_aux.shape[-2:] == (H0, W0) for s in samples)
        same_grid = all(g == grids[0] for g in grids)
        aux_b = vt_b = None                      # batch-contiguous level maps (the packed HFRE path)
        if same_aux:
            st = eng.davit_forward([s.image_aux for s in samples])
            aux_b = st
            aux = None
        else:
            aux = []
            for s in samples:
                st = eng.davit_forward([s.image_aux])
                aux.append([st[l][0] for l in range(4)])
        self._mark("davit")
        tok_off = np.cumsum([0] + [gh * gw for gh, gw in grids])
        hid = eng.cfg.vit["hidden_size"]
        vt = None
        if self.vt_mode == "fpn":
            if same_grid:
                gh, gw = grids[0]
                vt_b = eng.fpn_forward(taps[-1].view(B, gh, gw, hid))           # SimpleFPN on the LAST tap (:82-83)
            else:
                vt = []
                for b, (gh, gw) in enumerate(grids):
                    pyr = eng.fpn_forward(taps[-1][tok_off[b]:tok_off[b + 1]].view(1, gh, gw, hid))
                    vt.append([pyr[l][0] for l in range(4)])
        else:
            vt = [[taps[t][tok_off[b]:tok_off[b + 1]].view(grids[b][0], grids[b][1], hid) for t in range(len(taps))] for b in range(B)]
        self._mark("fpn")
        packed = aux_b is not None and vt_b is not None
        if not packed or self._keep_stages is not None:
            if aux is None:
                aux = [[aux_b[l][b] for l in range(4)] for b in range(B)]
            if vt is None:
                vt = [[vt_b[l][b] for l in range(4)] for b in range(B)]
        # boxes: the aux set in image pixels, the vt set rescaled to the ViT grid (:94-99); one packed array per set
        p = eng.cfg.vit["patch_size"]
        host_boxes = all(not s.boxes.is_cuda for s in samples)
        bl, sl = [], []
        for b, s in enumerate(samples):
            bx = s.boxes
            if bx.numel() == 0:
                bx = torch.tensor([[0.0, 10.0, 0.0, 10.0]], device=bx.device)        # the reference's dummy box (:90-91)
            Ha, Wa = s.image_aux.shape[-2:]
            gh, gw = grids[b]
            bl.append(bx)
            sl.append(np.tile(np.asarray([gw * p / Wa, gh * p / Ha, gw * p / Wa, gh * p / Ha], dtype=np.float32), (bx.shape[0], 1)))
        counts = [int(x.shape[0]) for x in bl]
        scale_rows = torch.from_numpy(np.concatenate(sl, 0))
        if host_boxes:     # fp32 multiply on the host: the same IEEE product the reference forms on the device; two copies instead of 4 B
            ba_h = torch.cat([x.to(torch.float32) for x in bl], 0)
            both = torch.stack([ba_h, ba_h * scale_rows], 0).pin_memory()
            both_d = both.to(dev, non_blocking=True)
            boxes_aux_p, boxes_vt_p = both_d[0], both_d[1]
        else:
            boxes_aux_p = torch.cat([x.to(dev, torch.float32) for x in bl], 0)
            boxes_vt_p = boxes_aux_p * scale_rows.to(dev, non_blocking=True)
        if packed:
            region_f32_p, region_bf16_p = HF.hfre_forward_packed(aux_b, vt_b, boxes_aux_p, boxes_vt_p, counts, self.hcfg, grids[0],
                                                                 want_bf16=True, workspace=self.ws)
        else:
            ba = list(torch.split(boxes_aux_p, counts, 0)); bv = list(torch.split(boxes_vt_p, counts, 0))
            rf, rb = HF.hfre_forward(aux, vt, ba, bv, self.hcfg, grids, want_bf16=True, workspace=self.ws)
            region_f32_p, region_bf16_p = torch.cat(rf, 0), torch.cat(rb, 0)
        self._mark("hfre")
        region_f32 = list(torch.split(region_f32_p, counts, 0))
        tokens = eng.region_project(region_bf16_p)
        self._mark("projector")
        region_tokens = list(torch.split(tokens, counts, 0))
        unit = eng.cfg.vit["spatial_merge_size"] ** 2
        img_off = np.cumsum([0] + [gh * gw // unit for gh, gw in grids])
        if self._keep_stages is not None:
            self._keep_stages.update(image_features=feats, taps=taps, davit=aux, fpn=vt if self.vt_mode == "fpn" else None,
                                     region_f32=region_f32, region_tokens=region_tokens)
        return feats, img_off, region_tokens, region_f32

    def encode_stages(self, samples: Sequence[SampleInputs]) -> dict:
        """encode() that also hands back every intermediate map for the stage-drift test: 'taps' (list over the tap layers of
        [sum gh*gw, hidden]), 'davit' / 'fpn' as [level][image] maps, 'image_features', 'region_f32', 'region_tokens'."""
        self._keep_stages = {}
        try:
            self.encode(samples)
            out = self._keep_stages
        finally:
            self._keep_stages = None
        B = len(samples)
        out["davit"] = [[out["davit"][b][l] for b in range(B)] for l in range(4)]
        if out["fpn"] is not None:
            out["fpn"] = [[out["fpn"][b][l] for b in range(B)] for l in range(4)]
        return out

    # ---- language side -----------------------------------------------------------------------------------
    def generate(self, samples: Sequence[SampleInputs], max_new_tokens: int, stop_ids: Sequence[int], pad_id: int = 151643,
                 early_exit_interval: int = 8, want_prefill_logits: bool = False):
        eng, dev = self.eng, self.eng.device
        samples = self.preprocess(samples)           # uint8 images -> the towers' input tensors, on the device
        feats, img_off, region_tokens, _ = self.encode(samples)
        # splice + M-RoPE bookkeeping of the whole batch in one launch on the device (fo1_splice_plan_batch); the per-sample host
        # function fo1_splice_plan stays the bit-exact reference of it (tests/test_gpu_prompt.py)
        merge = eng.cfg.vit["spatial_merge_size"]
        plan = splice_plan_batch([s.input_ids for s in samples], [[s.grid_hw] for s in samples], [r.shape[0] for r in region_tokens], dev,
                                 merge=merge, **self.ids)
        lens = plan["lens"]
        embeds = eng.build_embeds(plan["kind"], plan["index"], feats, torch.cat(region_tokens, 0))
        pos = plan["position_ids"]
        if self.check_plan:                                      # opt-in: surface a malformed prompt as an exception (one D2H sync)
            st = plan["status"].cpu().tolist()
            if any(st):
                raise ValueError(f"fo1_splice_plan_batch: per-sample status {st} (placeholders and features disagree)")
        deltas = plan["rope_delta"]
        self._mark("splice")
        out = eng.generate(embeds, pos, lens, deltas, max_new_tokens, stop_ids, pad_id, want_prefill_logits=want_prefill_logits,
                           early_exit_interval=early_exit_interval)
        self._mark("llm_prefill_decode")
        out["prompt_lens"] = lens
        return out
