"""CPU oracle for the LLM half: embedding splice, M-RoPE index, Qwen2.5 decoder prefill + greedy decode
-- TEST INFRASTRUCTURE ONLY.

fp32 restatement of, in the reference:
  * prepare_inputs_labels_for_qwen2_5_vl_multimodal     vlm_fo1/model/language_model/omchat_qwen2_5_vl.py:291-373, 434-458
  * get_rope_index                                       vlm_fo1/model/multimodal_encoder/qwen2_5_vl/modeling_qwen2_5_vl.py:1546-1721
  * Qwen2RMSNorm / rotary / M-RoPE sections / eager attention / Qwen2MLP / decoder layer / model loop / lm_head
                                                         modeling_qwen2_5_vl.py:126-140, 603-624, 643-685, 738-802, 627-640, 1064-1095, 1126-1242, 1876
  * decode position rule                                 modeling_qwen2_5_vl.py:1848-1860
The whole OmChatQwen25VLForCausalLM cannot be constructed under the container's transformers 5.5
(SURVEY.md section 8c); tests/golden/llm_small.npz pins this file against the reference's own
Qwen2_5_VLDecoderLayer / Qwen2RMSNorm / Qwen2_5_VLRotaryEmbedding modules and its get_rope_index function.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

try:
    from .precision import R
except ImportError:  # run from inside the directory (gen_golden.py)
    from precision import R  # type: ignore

IMAGE_PLACEHOLDER = -200   # vlm_fo1/constants.py IMAGE_TOKEN_INDEX
REGION_PLACEHOLDER = -300  # DEFAULT_REGION_INDEX


def splice_plan(input_ids: Sequence[int], image_grids: Sequence[Tuple[int, int]], n_regions: int, image_token_id: int = 151655,
                merge: int = 2):
    """-> (new_ids, src_kind, src_index): every -200 expands to its image's gh*gw/merge^2 rows (ids <- image_token_id),
    every -300 keeps one slot (id stays -300)   (omchat_qwen2_5_vl.py:318-368)."""
    new_ids: List[int] = []
    kind: List[int] = []
    index: List[int] = []
    img = reg = row = 0
    for t in input_ids:
        t = int(t)
        if t == IMAGE_PLACEHOLDER:
            gh, gw = image_grids[img]
            n = gh * gw // (merge * merge)
            new_ids += [image_token_id] * n; kind += [1] * n; index += list(range(row, row + n))
            row += n; img += 1
        elif t == REGION_PLACEHOLDER:
            assert reg < n_regions
            new_ids.append(REGION_PLACEHOLDER); kind.append(2); index.append(reg); reg += 1
        else:
            new_ids.append(t); kind.append(0); index.append(t)
    return new_ids, kind, index


def rope_index(ids: Sequence[int], image_grids: Sequence[Tuple[int, int]], image_token_id: int = 151655,
               vision_start_token_id: int = 151652, merge: int = 2):
    """get_rope_index for one un-padded sample with images only -> (position_ids [3, L] int64, rope_delta)."""
    ids = [int(t) for t in ids]
    L = len(ids)
    n_img = sum(1 for i in range(L - 1) if ids[i] == vision_start_token_id and ids[i + 1] == image_token_id)
    chunks: List[torch.Tensor] = []
    st = 0
    for k in range(n_img):
        ed = ids.index(image_token_id, st)
        gh, gw = image_grids[k]
        lh, lw = gh // merge, gw // merge
        text_len = ed - st
        base = int(chunks[-1].max()) + 1 if chunks else 0
        chunks.append(torch.arange(text_len).view(1, -1).expand(3, -1) + base)
        t_idx = torch.zeros(lh * lw, dtype=torch.long)
        h_idx = torch.arange(lh).view(-1, 1).expand(-1, lw).flatten()
        w_idx = torch.arange(lw).view(1, -1).expand(lh, -1).flatten()
        chunks.append(torch.stack([t_idx, h_idx, w_idx]) + text_len + base)
        st = ed + lh * lw
    if st < L:
        base = int(chunks[-1].max()) + 1 if chunks else 0
        chunks.append(torch.arange(L - st).view(1, -1).expand(3, -1) + base)
    pos = torch.cat(chunks, dim=1).reshape(3, -1)
    return pos, int(pos.max()) + 1 - L


def rms_norm(x, w, eps):
    return R(w * R(x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)))   # :137-140 casts before the gain


def _rot_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def mrope_cos_sin(pos3: torch.Tensor, head_dim: int, theta: float, sections: Sequence[int]):
    """pos3 [3, L] -> cos, sin [L, head_dim] with the t/h/w sections interleaved (:603-624, :675-681)."""
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.float32) / head_dim))
    freqs = pos3.float()[:, :, None] * inv[None, None, :]          # [3, L, hd/2]
    emb = torch.cat([freqs, freqs], dim=-1)                          # [3, L, hd]
    sec = list(sections) * 2
    cos = torch.cat([m[i % 3] for i, m in enumerate(emb.cos().split(sec, dim=-1))], dim=-1)
    sin = torch.cat([m[i % 3] for i, m in enumerate(emb.sin().split(sec, dim=-1))], dim=-1)
    return cos, sin


class Decoder:
    """Functional Qwen2.5 decoder over one sequence with a growing K/V list (fp32)."""

    def __init__(self, sd: Dict[str, torch.Tensor], cfg: dict):
        self.w = {k: v.float() for k, v in sd.items()}
        self.cfg = cfg
        self.H = cfg["hidden_size"]; self.nh = cfg["num_attention_heads"]; self.nkv = cfg["num_key_value_heads"]
        self.hd = self.H // self.nh
        self.kv: List[Tuple[torch.Tensor, torch.Tensor]] = [(None, None)] * cfg["num_hidden_layers"]
        self.layer_seconds: List[float] = []   # per-layer wall time of the last forward() (bench extrapolation)

    def forward(self, x: torch.Tensor, pos3: torch.Tensor) -> torch.Tensor:
        """x [n, H] new rows at positions pos3 [3, n]; attends causally to everything cached -> hidden [n, H]."""
        c, w = self.cfg, self.w
        cos, sin = mrope_cos_sin(pos3, self.hd, c["rope_theta"], c["mrope_section"])
        n = x.shape[0]
        import time as _time
        self.layer_seconds = []
        for i in range(c["num_hidden_layers"]):
            _tl = _time.perf_counter()
            p = f"layers.{i}."
            y = rms_norm(x, w[p + "input_layernorm.weight"], c["rms_norm_eps"])
            q = R(y @ w[p + "self_attn.q_proj.weight"].t() + w[p + "self_attn.q_proj.bias"]).reshape(n, self.nh, self.hd)
            k = R(y @ w[p + "self_attn.k_proj.weight"].t() + w[p + "self_attn.k_proj.bias"]).reshape(n, self.nkv, self.hd)
            v = R(y @ w[p + "self_attn.v_proj.weight"].t() + w[p + "self_attn.v_proj.bias"]).reshape(n, self.nkv, self.hd)
            q = R(q * cos[:, None, :] + _rot_half(q) * sin[:, None, :])
            k = R(k * cos[:, None, :] + _rot_half(k) * sin[:, None, :])
            pk, pv = self.kv[i]
            k_all = k if pk is None else torch.cat([pk, k], 0)
            v_all = v if pv is None else torch.cat([pv, v], 0)
            self.kv[i] = (k_all, v_all)
            past = k_all.shape[0] - n
            rep = self.nh // self.nkv
            s = torch.einsum("qhd,khd->hqk", q, k_all.repeat_interleave(rep, 1)) / math.sqrt(self.hd)
            mask = torch.arange(k_all.shape[0])[None, :] > (past + torch.arange(n))[:, None]
            s = s.masked_fill(mask[None], float("-inf"))
            a = R(torch.einsum("hqk,khd->qhd", s.softmax(-1), v_all.repeat_interleave(rep, 1))).reshape(n, self.H)
            x = R(x + a @ w[p + "self_attn.o_proj.weight"].t())
            y = rms_norm(x, w[p + "post_attention_layernorm.weight"], c["rms_norm_eps"])
            x = R(x + R(F.silu(y @ w[p + "mlp.gate_proj.weight"].t()) * (y @ w[p + "mlp.up_proj.weight"].t())) @ w[p + "mlp.down_proj.weight"].t())
            self.layer_seconds.append(_time.perf_counter() - _tl)
        return x

    def logits(self, hidden: torch.Tensor) -> torch.Tensor:
        head = self.w["lm_head.weight"] if "lm_head.weight" in self.w else self.w["embed_tokens.weight"]
        return rms_norm(hidden, self.w["norm.weight"], self.cfg["rms_norm_eps"]) @ head.t()


def generate(sd, cfg, inputs_embeds: torch.Tensor, pos3: torch.Tensor, rope_delta: int, max_new: int, stop_ids: Sequence[int],
             forced: Sequence[int] = None):
    """Greedy decode of one sequence.  Returns (tokens, logits per step [steps, V], top1-top2 margins).
    ``forced``: teacher forcing -- feed these tokens instead of the argmax (the logits are still returned)."""
    dec = Decoder(sd, cfg)
    L = inputs_embeds.shape[0]
    h = dec.forward(inputs_embeds.float(), pos3)
    all_prompt_logits = dec.logits(h)
    lg = all_prompt_logits[-1]
    toks, lgs = [], []
    emb = dec.w["embed_tokens.weight"]
    for s in range(max_new):
        lgs.append(lg)
        tok = int(lg.argmax()) if forced is None else int(forced[s])
        toks.append(tok)
        if (forced is None and tok in stop_ids) or s == max_new - 1:
            break
        p = L + s + rope_delta                        # cache_position[0] + rope_deltas (:1848-1860)
        h = dec.forward(emb[tok][None, :], torch.full((3, 1), p, dtype=torch.long))
        lg = dec.logits(h)[0]
    return toks, torch.stack(lgs), all_prompt_logits
