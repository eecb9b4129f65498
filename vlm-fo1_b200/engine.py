"""Model handle over the C ABI: owns the fo1_model*, keeps the prepared weight tensors alive, and exposes
the stage entry points with torch tensors as memory handles."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import check, lib

_DT = {torch.bfloat16: _lib.FO1_BF16, torch.float32: _lib.FO1_F32}


class ModelConfigC(C.Structure):
    _fields_ = [
        ("vit_depth", C.c_int32), ("vit_hidden", C.c_int32), ("vit_heads", C.c_int32), ("vit_inter", C.c_int32),
        ("vit_inter_pad", C.c_int32), ("vit_out_hidden", C.c_int32), ("vit_patch", C.c_int32), ("vit_merge", C.c_int32),
        ("vit_temporal", C.c_int32), ("vit_in_ch", C.c_int32), ("vit_window", C.c_int32),
        ("vit_n_fullatt", C.c_int32), ("vit_fullatt", C.c_int32 * 8),
        ("davit_dims", C.c_int32 * 4), ("davit_depths", C.c_int32 * 4), ("davit_heads", C.c_int32 * 4),
        ("davit_groups", C.c_int32 * 4), ("davit_window", C.c_int32),
        ("fpn_out", C.c_int32), ("region_dim", C.c_int32), ("proj_aux_layers", C.c_int32), ("proj_img_layers", C.c_int32),
        ("llm_layers", C.c_int32), ("llm_hidden", C.c_int32), ("llm_heads", C.c_int32), ("llm_kv_heads", C.c_int32),
        ("llm_head_dim", C.c_int32), ("llm_inter", C.c_int32), ("llm_vocab", C.c_int32),
        ("mrope_section", C.c_int32 * 3), ("rope_theta", C.c_float), ("rms_eps", C.c_float), ("tie_embeddings", C.c_int32),
    ]


@dataclass
class EngineConfig:
    """Python-side mirror of fo1_model_config; defaults are the released 3B checkpoint's architecture
    (ViT 32x1280, DaViT-large, SimpleFPN 512, Qwen2.5-3B)."""
    vit: dict = field(default_factory=lambda: dict(depth=32, hidden_size=1280, num_heads=16, intermediate_size=3420,
                                                   out_hidden_size=2048, patch_size=14, spatial_merge_size=2,
                                                   temporal_patch_size=2, in_channels=3, window_size=112,
                                                   fullatt_block_indexes=[7, 15, 23, 31]))
    davit: dict = field(default_factory=lambda: dict(depths=[1, 1, 9, 1], dim_embed=[256, 512, 1024, 2048],
                                                     num_heads=[8, 16, 32, 64], num_groups=[8, 16, 32, 64], window_size=12))
    fpn_out: int = 512
    region_dim: int = 5888
    proj_aux_layers: int = 2
    proj_img_layers: int = 0
    llm: dict = field(default_factory=lambda: dict(num_hidden_layers=36, hidden_size=2048, num_attention_heads=16,
                                                   num_key_value_heads=2, intermediate_size=11008, vocab_size=151936,
                                                   rope_theta=1000000.0, rms_norm_eps=1e-6, mrope_section=[16, 24, 24],
                                                   tie_word_embeddings=True))
    use_vit: bool = True
    use_davit: bool = True
    use_llm: bool = True

    def to_c(self) -> ModelConfigC:
        from .weights import vit_inter_pad
        c = ModelConfigC()
        if self.use_vit:
            v = self.vit
            c.vit_depth, c.vit_hidden, c.vit_heads = v["depth"], v["hidden_size"], v["num_heads"]
            c.vit_inter, c.vit_inter_pad = v["intermediate_size"], vit_inter_pad(v["intermediate_size"])
            c.vit_out_hidden, c.vit_patch, c.vit_merge = v["out_hidden_size"], v["patch_size"], v["spatial_merge_size"]
            c.vit_temporal, c.vit_in_ch, c.vit_window = v["temporal_patch_size"], v.get("in_channels", 3), v["window_size"]
            fa = list(v["fullatt_block_indexes"])
            c.vit_n_fullatt = len(fa)
            for i, x in enumerate(fa):
                c.vit_fullatt[i] = x
            c.fpn_out = self.fpn_out
        if self.use_davit:
            d = self.davit
            for i in range(4):
                c.davit_dims[i], c.davit_depths[i] = d["dim_embed"][i], d["depths"][i]
                c.davit_heads[i], c.davit_groups[i] = d["num_heads"][i], d["num_groups"][i]
            c.davit_window = d["window_size"]
        c.region_dim, c.proj_aux_layers, c.proj_img_layers = self.region_dim, self.proj_aux_layers, self.proj_img_layers
        l = self.llm
        c.llm_hidden = l["hidden_size"]
        if self.use_llm:
            c.llm_layers, c.llm_heads, c.llm_kv_heads = l["num_hidden_layers"], l["num_attention_heads"], l["num_key_value_heads"]
            c.llm_head_dim = l["hidden_size"] // l["num_attention_heads"]
            c.llm_inter, c.llm_vocab = l["intermediate_size"], l["vocab_size"]
            for i in range(3):
                c.mrope_section[i] = l["mrope_section"][i]
            c.rope_theta, c.rms_eps = l["rope_theta"], l["rms_norm_eps"]
            c.tie_embeddings = 1 if l.get("tie_word_embeddings", False) else 0
        return c


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Engine:
    def __init__(self, cfg: EngineConfig, device: Optional[torch.device] = None):
        self.cfg = cfg
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self._c = cfg.to_c()
        self._h = C.c_void_p()
        L = lib()
        L.fo1_model_create.restype = C.c_int
        L.fo1_model_create.argtypes = [C.POINTER(ModelConfigC), C.POINTER(C.c_void_p)]
        L.fo1_model_destroy.restype = None
        L.fo1_model_destroy.argtypes = [C.c_void_p]
        L.fo1_model_set_weight.restype = C.c_int
        L.fo1_model_set_weight.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int64)]
        L.fo1_model_finalize.restype = C.c_int
        L.fo1_model_finalize.argtypes = [C.c_void_p]
        for fn in ("fo1_vit_forward", "fo1_davit_forward", "fo1_fpn_forward", "fo1_region_project", "fo1_image_project"):
            getattr(L, fn).restype = C.c_int
        check(L.fo1_model_create(C.byref(self._c), C.byref(self._h)), "fo1_model_create")
        self.weights: Dict[str, torch.Tensor] = {}

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                lib().fo1_model_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def last_decode_path(self) -> int:
        """1: the last generate() ran the persistent decode kernel, 0: the per-kernel CUDA graph, -1: none yet."""
        return int(lib().fo1_last_decode_path(self._h))

    # ---- weights ----
    def set_weights(self, tensors: Dict[str, torch.Tensor]) -> None:
        L = lib()
        for name, t in tensors.items():
            assert t.is_cuda and t.is_contiguous(), name
            self.weights[name] = t
            shape = (C.c_int64 * t.dim())(*t.shape)
            check(L.fo1_model_set_weight(self._h, name.encode(), C.c_void_p(t.data_ptr()), _DT[t.dtype], t.dim(), shape),
                  f"fo1_model_set_weight({name})")

    def finalize(self) -> None:
        check(lib().fo1_model_finalize(self._h), "fo1_model_finalize")

    # ---- stages ----
    def vit_forward(self, pixel_values: Sequence[torch.Tensor], grids: Sequence[Tuple[int, int]]):
        """-> (img_feats [sum gh*gw/4, out_hidden] bf16, taps: list over tap layers of [sum gh*gw, hidden] bf16).
        Image b's slice of a tap is its channels-last map [gh, gw, hidden]."""
        v = self.cfg.vit
        B = len(pixel_values)
        px = [p.to(self.device, torch.float32).contiguous() for p in pixel_values]
        T = sum(gh * gw for gh, gw in grids)
        unit = v["spatial_merge_size"] ** 2
        feats = torch.empty((T // unit, v["out_hidden_size"]), dtype=torch.bfloat16, device=self.device)
        taps = [torch.empty((T, v["hidden_size"]), dtype=torch.bfloat16, device=self.device) for _ in v["fullatt_block_indexes"]]
        pp = (C.c_void_p * B)(*[p.data_ptr() for p in px])
        g = (C.c_int32 * (2 * B))(*[x for gw_ in grids for x in gw_])
        tp = (C.c_void_p * len(taps))(*[t.data_ptr() for t in taps])
        check(lib().fo1_vit_forward(self._h, pp, g, B, C.c_void_p(feats.data_ptr()), tp, _stream()), "fo1_vit_forward")
        return feats, taps

    def davit_forward(self, images: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        """images: B x fp32 [3, H, W] (same size) -> 4 stage maps bf16 [B, H_s, W_s, C_s]."""
        B = len(images)
        im = [i.to(self.device, torch.float32).contiguous() for i in images]
        H, W = im[0].shape[-2:]
        assert all(i.shape[-2:] == (H, W) for i in im), "fo1_davit_forward batches images of one size"
        outs = []
        h, w = (H + 6 - 7) // 4 + 1, (W + 6 - 7) // 4 + 1
        for s in range(4):
            if s:
                h, w = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
            outs.append(torch.empty((B, h, w, self.cfg.davit["dim_embed"][s]), dtype=torch.bfloat16, device=self.device))
        ip = (C.c_void_p * B)(*[i.data_ptr() for i in im])
        op = (C.c_void_p * 4)(*[o.data_ptr() for o in outs])
        check(lib().fo1_davit_forward(self._h, ip, H, W, B, op, _stream()), "fo1_davit_forward")
        return outs

    def fpn_forward(self, tap: torch.Tensor) -> List[torch.Tensor]:
        """tap bf16 [B, gh, gw, hidden] -> 4 levels bf16 [B, gh*f, gw*f, fpn_out]."""
        B, gh, gw, _ = tap.shape
        tap = tap.contiguous()
        dims = [(gh * 4, gw * 4), (gh * 2, gw * 2), (gh, gw), (gh // 2, gw // 2)]
        outs = [torch.empty((B, h, w, self.cfg.fpn_out), dtype=torch.bfloat16, device=self.device) for h, w in dims]
        op = (C.c_void_p * 4)(*[o.data_ptr() for o in outs])
        check(lib().fo1_fpn_forward(self._h, C.c_void_p(tap.data_ptr()), gh, gw, B, op, _stream()), "fo1_fpn_forward")
        return outs

    def image_project(self, feats: torch.Tensor) -> torch.Tensor:
        """mm_projector on the merged image tokens (encode_images, omchat_qwen2_5_vl.py:58-66); identity when proj_img_layers == 0."""
        if self.cfg.proj_img_layers == 0:
            return feats
        feats = feats.contiguous()
        out = torch.empty((feats.shape[0], self.cfg.llm["hidden_size"]), dtype=torch.bfloat16, device=self.device)
        check(lib().fo1_image_project(self._h, C.c_void_p(feats.data_ptr()), feats.shape[0], C.c_void_p(out.data_ptr()), _stream()),
              "fo1_image_project")
        return out

    def region_project(self, feats: torch.Tensor) -> torch.Tensor:
        feats = feats.contiguous()
        out = torch.empty((feats.shape[0], self.cfg.llm["hidden_size"]), dtype=torch.bfloat16, device=self.device)
        check(lib().fo1_region_project(self._h, C.c_void_p(feats.data_ptr()), feats.shape[0], C.c_void_p(out.data_ptr()), _stream()),
              "fo1_region_project")
        return out


class SpliceCfgC(C.Structure):
    _fields_ = [("image_token_id", C.c_int32), ("video_token_id", C.c_int32), ("vision_start_token_id", C.c_int32),
                ("merge", C.c_int32), ("image_placeholder", C.c_int32), ("region_placeholder", C.c_int32)]


class GenerateDescC(C.Structure):
    _fields_ = [("n_seqs", C.c_int32), ("seq_lens", C.POINTER(C.c_int32)), ("inputs_embeds", C.c_void_p),
                ("position_ids", C.c_void_p), ("rope_deltas", C.POINTER(C.c_int32)), ("max_new_tokens", C.c_int32),
                ("stop_ids", C.POINTER(C.c_int32)), ("n_stop_ids", C.c_int32), ("pad_id", C.c_int32),
                ("out_tokens", C.c_void_p), ("out_lens", C.c_void_p), ("prefill_logits", C.c_void_p),
                ("all_logits", C.c_void_p), ("early_exit_interval", C.c_int32), ("steps_run", C.c_int32), ("rope_deltas_device", C.c_void_p)]


def splice_plan(input_ids: Sequence[int], image_grids: Sequence[Tuple[int, int]], n_regions: int,
                image_token_id: int = 151655, video_token_id: int = 151656, vision_start_token_id: int = 151652,
                merge: int = 2, image_placeholder: int = -200, region_placeholder: int = -300):
    """Host-side integer work for one sample (fo1_splice_plan): -> dict(new_ids, kind, index, position_ids [3, L], rope_delta)."""
    import numpy as np
    L = lib()
    L.fo1_splice_plan.restype = C.c_int
    cfg = SpliceCfgC(image_token_id, video_token_id, vision_start_token_id, merge, image_placeholder, region_placeholder)
    ids = (C.c_int64 * len(input_ids))(*[int(t) for t in input_ids])
    grids = (C.c_int32 * max(2 * len(image_grids), 1))(*[int(x) for g in image_grids for x in g])
    cap = len(input_ids) + sum(gh * gw // (merge * merge) for gh, gw in image_grids) + 8
    new_ids = np.empty(cap, dtype=np.int64); kind = np.empty(cap, dtype=np.int32); index = np.empty(cap, dtype=np.int32)
    pos = np.empty((3, cap), dtype=np.int32)
    delta = C.c_int32(); n = C.c_int32()
    check(L.fo1_splice_plan(ids, len(input_ids), grids, len(image_grids), n_regions, C.byref(cfg),
                            new_ids.ctypes.data_as(C.POINTER(C.c_int64)), kind.ctypes.data_as(C.POINTER(C.c_int32)),
                            index.ctypes.data_as(C.POINTER(C.c_int32)), pos.ctypes.data_as(C.POINTER(C.c_int32)),
                            C.byref(delta), C.byref(n), cap), "fo1_splice_plan")
    k = n.value
    return dict(new_ids=new_ids[:k], kind=kind[:k], index=index[:k], position_ids=pos[:, :k].copy(), rope_delta=delta.value)


def _engine_build_embeds(self, kind: torch.Tensor, index: torch.Tensor, img_feats: Optional[torch.Tensor],
                         region_feats: Optional[torch.Tensor]) -> torch.Tensor:
    """inputs_embeds rows from the embedding table / image features / region features (device int32 kind, index)."""
    n = kind.numel()
    out = torch.empty((n, self.cfg.llm["hidden_size"]), dtype=torch.bfloat16, device=self.device)
    L = lib()
    L.fo1_llm_build_embeds.restype = C.c_int
    check(L.fo1_llm_build_embeds(self._h, C.c_void_p(kind.data_ptr()), C.c_void_p(index.data_ptr()), n,
                                 C.c_void_p(img_feats.data_ptr() if img_feats is not None else 0),
                                 C.c_void_p(region_feats.data_ptr() if region_feats is not None else 0),
                                 C.c_void_p(out.data_ptr()), _stream()), "fo1_llm_build_embeds")
    return out


def _engine_generate(self, inputs_embeds: torch.Tensor, position_ids: torch.Tensor, seq_lens: Sequence[int],
                     rope_deltas: Sequence[int], max_new_tokens: int, stop_ids: Sequence[int], pad_id: int,
                     want_prefill_logits: bool = False, want_all_logits: bool = False, early_exit_interval: int = 8):
    """Prefill + greedy decode of a packed batch.  position_ids int32 [3, sum L].  Returns a dict of device tensors."""
    B = len(seq_lens)
    V = self.cfg.llm["vocab_size"]
    T = int(sum(seq_lens))
    assert inputs_embeds.shape == (T, self.cfg.llm["hidden_size"]) and inputs_embeds.dtype == torch.bfloat16
    inputs_embeds = inputs_embeds.contiguous()
    position_ids = position_ids.to(self.device, torch.int32).contiguous()
    assert position_ids.shape == (3, T)
    toks = torch.empty((B, max(max_new_tokens, 1)), dtype=torch.int32, device=self.device)
    lens = torch.zeros((B,), dtype=torch.int32, device=self.device)
    pl = torch.empty((B, V), dtype=torch.float32, device=self.device) if want_prefill_logits else None
    al = torch.empty((T, V), dtype=torch.float32, device=self.device) if want_all_logits else None
    d = GenerateDescC()
    d.n_seqs = B
    sl = (C.c_int32 * B)(*[int(x) for x in seq_lens])
    dev_deltas = rope_deltas if isinstance(rope_deltas, torch.Tensor) else None       # device int32 [B] from splice_plan_batch
    rd = (C.c_int32 * B)(*([0] * B if dev_deltas is not None else [int(x) for x in rope_deltas]))
    if dev_deltas is not None:
        assert dev_deltas.is_cuda and dev_deltas.dtype == torch.int32 and dev_deltas.numel() == B
        d.rope_deltas_device = dev_deltas.data_ptr()
    st = (C.c_int32 * max(len(stop_ids), 1))(*[int(x) for x in stop_ids])
    d.seq_lens, d.rope_deltas, d.stop_ids, d.n_stop_ids = sl, rd, st, len(stop_ids)
    d.inputs_embeds, d.position_ids = inputs_embeds.data_ptr(), position_ids.data_ptr()
    d.max_new_tokens, d.pad_id = max_new_tokens, pad_id
    d.out_tokens, d.out_lens = toks.data_ptr(), lens.data_ptr()
    d.prefill_logits = pl.data_ptr() if pl is not None else None
    d.all_logits = al.data_ptr() if al is not None else None
    d.early_exit_interval = early_exit_interval
    L = lib()
    L.fo1_llm_generate.restype = C.c_int
    check(L.fo1_llm_generate(self._h, C.byref(d), _stream()), "fo1_llm_generate")
    return dict(tokens=toks, lens=lens, prefill_logits=pl, all_logits=al, steps_run=d.steps_run)


Engine.build_embeds = _engine_build_embeds
Engine.generate = _engine_generate


def splice_plan_batch(prompts: Sequence[Sequence[int]], image_grids: Sequence[Sequence[Tuple[int, int]]], n_regions: Sequence[int], device,
                      image_token_id: int = 151655, video_token_id: int = 151656, vision_start_token_id: int = 151652, merge: int = 2,
                      image_placeholder: int = -200, region_placeholder: int = -300):
    """The splice + M-RoPE bookkeeping of a whole batch in ONE launch on the device (fo1_splice_plan_batch).  Per sample b:
    ``prompts[b]`` ids with placeholders, ``image_grids[b]`` its images' (gh, gw), ``n_regions[b]`` region rows.  Image rows /
    region rows of the batch are assumed packed back to back in sample order (what Fo1Pipeline.encode produces).
    -> dict(kind, index, new_ids int32 [T]; position_ids int32 [3, T]; rope_delta int32 [B]) on the device, lens (host list)."""
    import numpy as np
    B = len(prompts)
    unit = merge * merge
    id_off = np.zeros(B + 1, np.int32); img_off = np.zeros(B + 1, np.int32); out_off = np.zeros(B + 1, np.int32)
    img_row = np.zeros(B, np.int32); reg_row = np.zeros(B, np.int32)
    rows_i = rows_r = 0
    for b in range(B):
        toks = sum(gh * gw // unit for gh, gw in image_grids[b])
        n_img_ph = sum(1 for t in prompts[b] if t == image_placeholder)
        id_off[b + 1] = id_off[b] + len(prompts[b]); img_off[b + 1] = img_off[b] + len(image_grids[b])
        # every image placeholder expands to its image's rows; more placeholders than images is reported per sample by the kernel
        exp = sum(gh * gw // unit for gh, gw in list(image_grids[b])[:n_img_ph])
        out_off[b + 1] = out_off[b] + len(prompts[b]) - min(n_img_ph, len(image_grids[b])) + exp
        img_row[b] = rows_i; reg_row[b] = rows_r
        rows_i += toks; rows_r += int(n_regions[b])
    T = int(out_off[B])
    host = np.concatenate([np.asarray([t for p in prompts for t in p], np.int32), id_off,
                           np.asarray([x for g in image_grids for hw in g for x in hw], np.int32).reshape(-1), img_off,
                           np.asarray(n_regions, np.int32), out_off, img_row, reg_row])
    d = torch.from_numpy(host).to(device, non_blocking=True)      # one H2D for every table
    cuts = np.cumsum([0, int(id_off[B]), B + 1, 2 * int(img_off[B]), B + 1, B, B + 1, B, B])
    v = [d[cuts[i]:cuts[i + 1]] for i in range(8)]
    out = torch.empty(6 * T + 2 * B, dtype=torch.int32, device=device)
    new_ids, kind, index, pos = out[:T], out[T:2 * T], out[2 * T:3 * T], out[3 * T:6 * T].view(3, T)
    delta, status = out[6 * T:6 * T + B], out[6 * T + B:]
    cfg = SpliceCfgC(image_token_id, video_token_id, vision_start_token_id, merge, image_placeholder, region_placeholder)
    L = lib()
    L.fo1_splice_plan_batch.restype = C.c_int
    L.fo1_splice_plan_batch.argtypes = [C.c_void_p] * 8 + [C.c_int32, C.c_int64, C.POINTER(SpliceCfgC)] + [C.c_void_p] * 7
    P = lambda t: C.c_void_p(t.data_ptr())
    check(L.fo1_splice_plan_batch(P(v[0]), P(v[1]), P(v[2]), P(v[3]), P(v[4]), P(v[5]), P(v[6]), P(v[7]), B, T, C.byref(cfg),
                                  P(new_ids), P(kind), P(index), P(pos), P(delta), P(status), _stream()), "fo1_splice_plan_batch")
    return dict(new_ids=new_ids, kind=kind, index=index, position_ids=pos, rope_delta=delta, status=status,
                lens=[int(out_off[b + 1] - out_off[b]) for b in range(B)])


def parse_predictions(tokens: torch.Tensor, lens: torch.Tensor, ground_ids: Tuple[int, int], objects_ids: Tuple[int, int],
                      region_ids: Sequence[int], newline_token_ids: Sequence[int] = (), vocab: int = 0, max_records: int = 256):
    """fo1_parse_predictions over a batch of decoded ids (device int32 [B, T], lens [B]) -> per sample list of
    (label first token, label end token, N or -1)."""
    import numpy as np
    B, T = tokens.shape
    dev = tokens.device
    rid = torch.tensor(list(region_ids), dtype=torch.int32, device=dev)
    bm = None
    if vocab > 0 and len(newline_token_ids):
        words = np.zeros((vocab + 31) // 32, np.uint32)
        for t in newline_token_ids:
            words[t >> 5] |= np.uint32(1 << (t & 31))
        bm = torch.from_numpy(words.view(np.int32)).to(dev)
    rec = torch.empty((B, max_records, 3), dtype=torch.int32, device=dev)
    cnt = torch.zeros((B,), dtype=torch.int32, device=dev)
    L = lib()
    L.fo1_parse_predictions.restype = C.c_int
    L.fo1_parse_predictions.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                        C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    tokens = tokens.contiguous(); lens = lens.to(torch.int32).contiguous()
    check(L.fo1_parse_predictions(C.c_void_p(tokens.data_ptr()), tokens.stride(0), C.c_void_p(lens.data_ptr()), B, ground_ids[0], ground_ids[1],
                                  objects_ids[0], objects_ids[1], C.c_void_p(rid.data_ptr()), rid.numel(),
                                  C.c_void_p(bm.data_ptr()) if bm is not None else None, vocab, C.c_void_p(rec.data_ptr()), max_records,
                                  C.c_void_p(cnt.data_ptr()), _stream()), "fo1_parse_predictions")
    rec_h, cnt_h = rec.cpu().numpy(), cnt.cpu().numpy()
    return [[tuple(int(x) for x in rec_h[b, i]) for i in range(int(cnt_h[b]))] for b in range(B)]
