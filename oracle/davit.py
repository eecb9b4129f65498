"""CPU oracle for the aux tower (DaViT) and SimpleFPN -- TEST INFRASTRUCTURE ONLY.

fp32 restatement of, in the reference:
  * DaViT.forward_features / ConvEmbed / SpatialBlock / WindowAttention / ChannelBlock / ChannelAttention / Mlp
        vlm_fo1/model/multimodal_encoder/davit/modeling_davit.py:478-506, 102-148, 285-315, 225-282, 175-206, 151-172, 51-69
  * SimpleFP.forward and its channel LayerNorm
        vlm_fo1/model/multimodal_visual_prompt_encoder/simple_fpn.py:100-216, 58-78
Weights by checkpoint names (state_dict of DaViT / SimpleFP).  Pinned by tests/golden/davit_small.npz and fpn_small.npz.
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn.functional as F

try:
    from .precision import R
except ImportError:  # run from inside the directory (gen_golden.py)
    from precision import R  # type: ignore


def _ln(x, w, b, eps=1e-5):
    return R(F.layer_norm(x, (x.shape[-1],), w, b, eps))


def _dw(x, H, W, w, b):  # tokens [N, C] -> + depth-wise 3x3
    C = x.shape[1]
    y = F.conv2d(x.t().reshape(1, C, H, W), w, b, padding=1, groups=C)
    return R(x + y.reshape(C, H * W).t())


def _window_attn(y, H, W, heads, ws, p, w):
    C = y.shape[1]
    pad_b, pad_r = (ws - H % ws) % ws, (ws - W % ws) % ws
    m = F.pad(y.reshape(H, W, C), (0, 0, 0, pad_r, 0, pad_b))          # zero pad AFTER the norm (:248-251)
    Hp, Wp = H + pad_b, W + pad_r
    win = m.reshape(Hp // ws, ws, Wp // ws, ws, C).permute(0, 2, 1, 3, 4).reshape(-1, ws * ws, C)
    qkv = R(win @ w[p + "qkv.weight"].t() + w[p + "qkv.bias"]).reshape(win.shape[0], ws * ws, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * (C // heads) ** -0.5, qkv[1], qkv[2]
    o = R((q @ k.transpose(-2, -1)).softmax(-1) @ v).transpose(1, 2).reshape(win.shape[0], ws * ws, C)
    o = R(o @ w[p + "proj.weight"].t() + w[p + "proj.bias"])
    o = o.reshape(Hp // ws, Wp // ws, ws, ws, C).permute(0, 2, 1, 3, 4).reshape(Hp, Wp, C)[:H, :W]
    return o.reshape(H * W, C)


def _channel_attn(y, groups, p, w):
    N, C = y.shape
    qkv = R(y @ w[p + "qkv.weight"].t() + w[p + "qkv.bias"]).reshape(N, 3, groups, C // groups).permute(1, 2, 0, 3)
    q, k, v = qkv[0] * float(N) ** -0.5, qkv[1], qkv[2]                  # [g, N, c]
    a = (q.transpose(-1, -2) @ k).softmax(-1)                              # [g, c, c]
    o = (a @ v.transpose(-1, -2)).transpose(-1, -2)                        # [g, N, c]
    o = R(o.permute(1, 0, 2).reshape(N, C))
    return R(o @ w[p + "proj.weight"].t() + w[p + "proj.bias"])


def davit_forward(sd: Dict[str, torch.Tensor], cfg: dict, image: torch.Tensor, timing: dict = None) -> List[torch.Tensor]:
    """image [3, H, W] -> 4 stage maps, channels-last [H_s, W_s, C_s]."""
    w = {k: v.float() for k, v in sd.items()}
    ws = cfg["window_size"]
    x4 = image.float().unsqueeze(0)
    outs = []
    x = None
    H = W = 0
    import time as _time
    if timing is not None:
        timing["embed_s"] = [0.0] * 4
        timing["block_s"] = [[] for _ in range(4)]
    for s in range(4):
        _t0 = _time.perf_counter()
        p = f"convs.{s}."
        if cfg["patch_prenorm"][s]:
            x = _ln(x, w[p + "norm.weight"], w[p + "norm.bias"])
            x4 = x.t().reshape(1, -1, H, W)
        y = R(F.conv2d(R(x4), w[p + "proj.weight"], w[p + "proj.bias"], stride=cfg["patch_stride"][s], padding=cfg["patch_padding"][s]))
        _, C, H, W = y.shape
        x = y.reshape(C, H * W).t()
        if not cfg["patch_prenorm"][s]:
            x = _ln(x, w[p + "norm.weight"], w[p + "norm.bias"])
        if timing is not None:
            timing["embed_s"][s] = _time.perf_counter() - _t0
        for j in range(cfg["depths"][s]):
            _tb = _time.perf_counter()
            for kind in ("spatial_block", "channel_block"):
                q = f"blocks.{s}.{j}.{kind}."
                x = _dw(x, H, W, w[q + "conv1.fn.dw.weight"], w[q + "conv1.fn.dw.bias"])
                if kind == "spatial_block":
                    a = "window_attn."
                    yy = _ln(x, w[q + a + "norm.weight"], w[q + a + "norm.bias"])
                    x = R(x + _window_attn(yy, H, W, cfg["num_heads"][s], ws, q + a + "fn.", w))
                else:
                    a = "channel_attn."
                    yy = _ln(x, w[q + a + "norm.weight"], w[q + a + "norm.bias"])
                    x = R(x + _channel_attn(yy, cfg["num_groups"][s], q + a + "fn.", w))
                x = _dw(x, H, W, w[q + "conv2.fn.dw.weight"], w[q + "conv2.fn.dw.bias"])
                yy = _ln(x, w[q + "ffn.norm.weight"], w[q + "ffn.norm.bias"])
                hmid = R(F.gelu(yy @ w[q + "ffn.fn.net.fc1.weight"].t() + w[q + "ffn.fn.net.fc1.bias"]))
                x = R(x + hmid @ w[q + "ffn.fn.net.fc2.weight"].t() + w[q + "ffn.fn.net.fc2.bias"])
            if timing is not None:
                timing["block_s"][s].append(_time.perf_counter() - _tb)
        outs.append(x.reshape(H, W, C).clone())
        x4 = None
    return outs


def _chan_ln(x, w, b, eps=1e-6):  # [1, C, H, W], simple_fpn.py:73-78
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    return R(w[None, :, None, None] * ((x - u) / torch.sqrt(s + eps)) + b[None, :, None, None])


def fpn_forward(sd: Dict[str, torch.Tensor], tap_hwc: torch.Tensor) -> List[torch.Tensor]:
    """tap [gh, gw, D] -> 4 pyramid levels, channels-last [gh*f, gw*f, out] for f = 4, 2, 1, 1/2."""
    w = {k: v.float() for k, v in sd.items()}
    x = tap_hwc.float().permute(2, 0, 1).unsqueeze(0)
    outs = []
    for stage in (1, 2, 3, 4):
        p = f"simfp_{stage}."
        y = x
        if stage == 1:
            y = R(F.conv_transpose2d(y, w[p + "0.weight"], w[p + "0.bias"], stride=2))
            y = R(F.gelu(_chan_ln(y, w[p + "1.weight"], w[p + "1.bias"])))
            y = R(F.conv_transpose2d(y, w[p + "3.weight"], w[p + "3.bias"], stride=2))
            i = 4
        elif stage == 2:
            y = R(F.conv_transpose2d(y, w[p + "0.weight"], w[p + "0.bias"], stride=2))
            i = 1
        elif stage == 3:
            i = 0
        else:
            y = F.max_pool2d(y, 2, 2)
            i = 1
        y = _chan_ln(R(F.conv2d(y, w[p + f"{i}.weight"])), w[p + f"{i}.norm.weight"], w[p + f"{i}.norm.bias"])
        y = _chan_ln(R(F.conv2d(y, w[p + f"{i + 1}.weight"], padding=1)), w[p + f"{i + 1}.norm.weight"], w[p + f"{i + 1}.norm.bias"])
        outs.append(y[0].permute(1, 2, 0).contiguous())
    return outs


def projector_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """mlpNx_gelu / linear projector (multimodal_projector/builder.py:100-106): keys '0.weight','2.weight',... or 'weight'."""
    w = {k: v.float() for k, v in sd.items()}
    if "weight" in w:
        return x.float() @ w["weight"].t() + w["bias"]
    y = x.float()
    idx = sorted({int(k.split(".")[0]) for k in w})
    for n, i in enumerate(idx):
        if n:
            y = R(F.gelu(y))
        y = R(y @ w[f"{i}.weight"].t() + w[f"{i}.bias"])
    return y
