"""CPU oracle for the primary tower (Qwen2.5-VL ViT with FO1's multi-level taps) -- TEST INFRASTRUCTURE ONLY.

fp32 restatement of, in the reference:
  * ``custom_forward``                         vlm_fo1/model/multimodal_encoder/qwen2_5_vl_encoder.py:86-158
  * ``VisionFeaturesGather.extract_...``        qwen2_5_vl_encoder.py:37-80
  * patch embed / RMSNorm / merger / eager attention / block / rot_pos_emb / get_window_index
                                                qwen2_5_vl/modeling_qwen2_5_vl.py:88-111, 126-140, 146-159, 233-280, 333-357, 436-504
Weights are taken by their checkpoint names (state_dict of Qwen2_5_VisionTransformerPretrainedModel).
Pinned by tests/golden/vit_small.npz (outputs of the reference module itself).
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


def window_index(gh: int, gw: int, window_size: int = 112, merge: int = 2, patch: int = 14) -> Tuple[List[int], List[int]]:
    """get_window_index (:465-504) for one image with t = 1, after unique_consecutive (encoder :109).
    Returns (window_index over merged cells, cu_window_seqlens over tokens)."""
    lh, lw = gh // merge, gw // merge
    ws = window_size // merge // patch
    pad_h, pad_w = ws - lh % ws, ws - lw % ws
    nwh, nww = (lh + pad_h) // ws, (lw + pad_w) // ws
    idx: List[int] = []
    cu = [0]
    for wy in range(nwh):
        for wx in range(nww):
            cnt = 0
            for iy in range(ws):
                for ix in range(ws):
                    y, x = wy * ws + iy, wx * ws + ix
                    if y < lh and x < lw:
                        idx.append(y * lw + x)
                        cnt += 1
            nxt = cu[-1] + cnt * merge * merge
            if nxt != cu[-1]:
                cu.append(nxt)
    return idx, cu


def patch_positions(gh: int, gw: int, merge: int = 2) -> torch.Tensor:
    """(h, w) index of every token in the processor's merge order (rot_pos_emb :436-458) -> [gh*gw, 2]."""
    h = torch.arange(gh).unsqueeze(1).expand(-1, gw).reshape(gh // merge, merge, gw // merge, merge).permute(0, 2, 1, 3).flatten()
    w = torch.arange(gw).unsqueeze(0).expand(gh, -1).reshape(gh // merge, merge, gw // merge, merge).permute(0, 2, 1, 3).flatten()
    return torch.stack([h, w], dim=-1)


def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    return R(w * R(x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)))   # :137-140 casts before the gain


def _rot_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def vit_forward(sd: Dict[str, torch.Tensor], cfg: dict, pixel_values: torch.Tensor, gh: int, gw: int, timing: dict = None):
    """One image.  Returns (merged tokens [gh*gw/4, out_hidden] in raster order,
    taps: list of [gh, gw, hidden] channels-last maps, one per full-attention layer)."""
    H, heads, merge = cfg["hidden_size"], cfg["num_heads"], cfg["spatial_merge_size"]
    unit = merge * merge
    hd = H // heads
    T = gh * gw
    import time as _time
    _t0 = _time.perf_counter()
    w = {k: v.float() for k, v in sd.items()}
    x = R(R(pixel_values.float()) @ w["patch_embed.proj.weight"].reshape(H, -1).t())
    widx, cu_win = window_index(gh, gw, cfg["window_size"], merge, cfg["patch_size"])
    perm = torch.tensor(widx)
    x = x.reshape(T // unit, unit, H)[perm].reshape(T, H)
    # 2-D rope angles in window order
    inv = 1.0 / (10000.0 ** (torch.arange(0, hd // 2, 2, dtype=torch.float32) / (hd // 2)))
    pos = patch_positions(gh, gw, merge).reshape(T // unit, unit, 2)[perm].reshape(T, 2).float()
    ang = torch.cat([pos[:, :1] * inv, pos[:, 1:] * inv], dim=1)   # [T, hd/2]
    emb = torch.cat([ang, ang], dim=1)
    cos, sin = emb.cos()[:, None, :], emb.sin()[:, None, :]
    taps = []
    if timing is not None:
        timing["embed_s"] = _time.perf_counter() - _t0
        timing["blocks"] = []
    for L in range(cfg["depth"]):
        _tb = _time.perf_counter()
        p = f"blocks.{L}."
        full = L in cfg["fullatt_block_indexes"]
        cu = [0, T] if full else cu_win
        y = rms_norm(x, w[p + "norm1.weight"])
        qkv = R(y @ w[p + "attn.qkv.weight"].t() + w[p + "attn.qkv.bias"]).reshape(T, 3, heads, hd)
        q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]
        q = R(q * cos + _rot_half(q) * sin)
        k = R(k * cos + _rot_half(k) * sin)
        out = torch.empty(T, heads, hd)
        seg = cu[1] - cu[0]
        if all(b - a == seg for a, b in zip(cu[:-1], cu[1:])):   # equal-length segments: one batched product
            n = len(cu) - 1
            qq, kk, vv = (t.reshape(n, seg, heads, hd) for t in (q, k, v))
            s = torch.einsum("nqhd,nkhd->nhqk", qq, kk) / math.sqrt(hd)
            out = torch.einsum("nhqk,nkhd->nqhd", s.softmax(-1), vv).reshape(T, heads, hd)
        else:
            for a, b in zip(cu[:-1], cu[1:]):
                s = torch.einsum("qhd,khd->hqk", q[a:b], k[a:b]) / math.sqrt(hd)
                out[a:b] = torch.einsum("hqk,khd->qhd", s.softmax(-1), v[a:b])
        x = R(x + R(out).reshape(T, H) @ w[p + "attn.proj.weight"].t() + w[p + "attn.proj.bias"])
        y = rms_norm(x, w[p + "norm2.weight"])
        g = y @ w[p + "mlp.gate_proj.weight"].t() + w[p + "mlp.gate_proj.bias"]
        u = y @ w[p + "mlp.up_proj.weight"].t() + w[p + "mlp.up_proj.bias"]
        x = R(x + R(F.silu(g) * u) @ w[p + "mlp.down_proj.weight"].t() + w[p + "mlp.down_proj.bias"])
        if full:
            # un-window: merged cell j of the sequence is cell widx[j] of the image; its 4 tokens are the 2x2 patches
            cells = torch.empty(T // unit, unit, H)
            cells[perm] = x.reshape(T // unit, unit, H)
            m = cells.reshape(gh // merge, gw // merge, merge, merge, H).permute(0, 2, 1, 3, 4).reshape(gh, gw, H)
            taps.append(m.clone())
        if timing is not None:
            timing["blocks"].append((bool(full), _time.perf_counter() - _tb))
    _tm = _time.perf_counter()
    y = rms_norm(x, w["merger.ln_q.weight"]).reshape(T // unit, unit * H)
    y = R(R(F.gelu(y @ w["merger.mlp.0.weight"].t() + w["merger.mlp.0.bias"])) @ w["merger.mlp.2.weight"].t() + w["merger.mlp.2.bias"])
    merged = torch.empty_like(y)
    merged[perm] = y
    if timing is not None:
        timing["merger_s"] = _time.perf_counter() - _tm
    return merged, taps
