"""Checkpoint tensors -> the layouts libfo1's engine consumes (one-time host-side prep, not hot path).

Input names are the reference checkpoint's (the state_dicts of the modules that
vlm_fo1/model/builder.py:90-131 loads: ``model.vision_tower.image_tower.*``,
``model.vision_tower_aux.image_tower.*``, ``model.object_vp_extractor.simple_fpn.*``,
``model.mm_projector_aux.*``, ``model.layers.*`` ...), with the component prefix already stripped.
Layout changes (all exact, zero padding only):
  * gate_proj / up_proj rows interleaved in [32 gate | 32 up] blocks (fused SiLU(g)*u GEMM epilogue),
    intermediate size zero-padded to a multiple of 32; down_proj K zero-padded to match;
  * conv weights flattened to GEMM layout: stem [C,3,7,7] -> [C,147->152]; 3x3 convs -> [Cout,(ky,kx,ci)];
    ConvTranspose2d(k2,s2) [Cin,Cout,2,2] -> [(dy,dx,co), ci] with the bias tiled 4x;
  * depth-wise 3x3 [C,1,3,3] -> tap-major [9, C];
  * LLM q/k/v projections concatenated to one [q|k|v] GEMM.
"""
from __future__ import annotations

from typing import Dict

import torch

from .ops import interleave_gate_up


def _dev(t: torch.Tensor, device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=torch.bfloat16).contiguous()


def pad_cols(w: torch.Tensor, k: int) -> torch.Tensor:
    if w.shape[1] == k:
        return w
    out = torch.zeros((w.shape[0], k), dtype=w.dtype, device=w.device)
    out[:, : w.shape[1]] = w
    return out


def vit_inter_pad(inter: int) -> int:
    return (inter + 31) // 32 * 32


def prepare_vit(sd: Dict[str, torch.Tensor], cfg: dict, device) -> Dict[str, torch.Tensor]:
    H = cfg["hidden_size"]
    ip = vit_inter_pad(cfg["intermediate_size"])
    out = {"vit.patch_embed.w": _dev(sd["patch_embed.proj.weight"].reshape(H, -1), device)}
    for i in range(cfg["depth"]):
        s, d = f"blocks.{i}.", f"vit.blk.{i}."
        out[d + "norm1.w"] = _dev(sd[s + "norm1.weight"], device)
        out[d + "qkv.w"] = _dev(sd[s + "attn.qkv.weight"], device)
        out[d + "qkv.b"] = _dev(sd[s + "attn.qkv.bias"], device)
        out[d + "proj.w"] = _dev(sd[s + "attn.proj.weight"], device)
        out[d + "proj.b"] = _dev(sd[s + "attn.proj.bias"], device)
        out[d + "norm2.w"] = _dev(sd[s + "norm2.weight"], device)
        out[d + "gateup.w"] = interleave_gate_up(_dev(sd[s + "mlp.gate_proj.weight"], device), _dev(sd[s + "mlp.up_proj.weight"], device))
        out[d + "gateup.b"] = interleave_gate_up(_dev(sd[s + "mlp.gate_proj.bias"], device), _dev(sd[s + "mlp.up_proj.bias"], device))
        out[d + "down.w"] = pad_cols(_dev(sd[s + "mlp.down_proj.weight"], device), ip)
        out[d + "down.b"] = _dev(sd[s + "mlp.down_proj.bias"], device)
    out["vit.merger.ln_q.w"] = _dev(sd["merger.ln_q.weight"], device)
    out["vit.merger.fc1.w"] = _dev(sd["merger.mlp.0.weight"], device)
    out["vit.merger.fc1.b"] = _dev(sd["merger.mlp.0.bias"], device)
    out["vit.merger.fc2.w"] = _dev(sd["merger.mlp.2.weight"], device)
    out["vit.merger.fc2.b"] = _dev(sd["merger.mlp.2.bias"], device)
    return out


def prepare_davit(sd: Dict[str, torch.Tensor], cfg: dict, device) -> Dict[str, torch.Tensor]:
    out = {}
    for s in range(4):
        p, d = f"convs.{s}.", f"davit.s{s}."
        w = sd[p + "proj.weight"]
        if s == 0:
            out[d + "conv.w"] = pad_cols(_dev(w.reshape(w.shape[0], -1), device), 152)
        else:
            out[d + "conv.w"] = _dev(w.permute(0, 2, 3, 1).reshape(w.shape[0], -1), device)
        out[d + "conv.b"] = _dev(sd[p + "proj.bias"], device)
        out[d + "norm.w"] = _dev(sd[p + "norm.weight"], device)
        out[d + "norm.b"] = _dev(sd[p + "norm.bias"], device)
        for j in range(cfg["depths"][s]):
            for kind, tag, attn in (("spatial_block", "sp", "window_attn"), ("channel_block", "ch", "channel_attn")):
                q, e = f"blocks.{s}.{j}.{kind}.", f"{d}b{j}.{tag}."
                for c in ("conv1", "conv2"):
                    dw = sd[q + c + ".fn.dw.weight"]
                    out[e + c + ".w9"] = _dev(dw.reshape(dw.shape[0], 9).t(), device)
                    out[e + c + ".b"] = _dev(sd[q + c + ".fn.dw.bias"], device)
                out[e + "norm1.w"] = _dev(sd[q + attn + ".norm.weight"], device)
                out[e + "norm1.b"] = _dev(sd[q + attn + ".norm.bias"], device)
                out[e + "qkv.w"] = _dev(sd[q + attn + ".fn.qkv.weight"], device)
                out[e + "qkv.b"] = _dev(sd[q + attn + ".fn.qkv.bias"], device)
                out[e + "proj.w"] = _dev(sd[q + attn + ".fn.proj.weight"], device)
                out[e + "proj.b"] = _dev(sd[q + attn + ".fn.proj.bias"], device)
                out[e + "norm2.w"] = _dev(sd[q + "ffn.norm.weight"], device)
                out[e + "norm2.b"] = _dev(sd[q + "ffn.norm.bias"], device)
                out[e + "fc1.w"] = _dev(sd[q + "ffn.fn.net.fc1.weight"], device)
                out[e + "fc1.b"] = _dev(sd[q + "ffn.fn.net.fc1.bias"], device)
                out[e + "fc2.w"] = _dev(sd[q + "ffn.fn.net.fc2.weight"], device)
                out[e + "fc2.b"] = _dev(sd[q + "ffn.fn.net.fc2.bias"], device)
    return out


def _deconv(w: torch.Tensor, b: torch.Tensor, device):
    cin, cout = w.shape[0], w.shape[1]
    return _dev(w.permute(2, 3, 1, 0).reshape(4 * cout, cin), device), _dev(b.repeat(4), device)


def prepare_fpn(sd: Dict[str, torch.Tensor], device) -> Dict[str, torch.Tensor]:
    """simple_fpn.py:139-181: stages simfp_1..4 = scales 4, 2, 1, 0.5."""
    out = {}
    out["fpn.l0.deconv1.w"], out["fpn.l0.deconv1.b"] = _deconv(sd["simfp_1.0.weight"], sd["simfp_1.0.bias"], device)
    out["fpn.l0.ln.w"] = _dev(sd["simfp_1.1.weight"], device)
    out["fpn.l0.ln.b"] = _dev(sd["simfp_1.1.bias"], device)
    out["fpn.l0.deconv2.w"], out["fpn.l0.deconv2.b"] = _deconv(sd["simfp_1.3.weight"], sd["simfp_1.3.bias"], device)
    out["fpn.l1.deconv1.w"], out["fpn.l1.deconv1.b"] = _deconv(sd["simfp_2.0.weight"], sd["simfp_2.0.bias"], device)
    for l, (stage, i) in enumerate(((1, 4), (2, 1), (3, 0), (4, 1))):
        p, d = f"simfp_{stage}.", f"fpn.l{l}."
        w1 = sd[p + f"{i}.weight"]
        out[d + "conv1.w"] = _dev(w1.reshape(w1.shape[0], -1), device)
        out[d + "ln1.w"] = _dev(sd[p + f"{i}.norm.weight"], device)
        out[d + "ln1.b"] = _dev(sd[p + f"{i}.norm.bias"], device)
        w2 = sd[p + f"{i + 1}.weight"]
        out[d + "conv2.w"] = _dev(w2.permute(0, 2, 3, 1).reshape(w2.shape[0], -1), device)
        out[d + "ln2.w"] = _dev(sd[p + f"{i + 1}.norm.weight"], device)
        out[d + "ln2.b"] = _dev(sd[p + f"{i + 1}.norm.bias"], device)
    return out


def prepare_projector(sd: Dict[str, torch.Tensor], prefix: str, device) -> Dict[str, torch.Tensor]:
    """nn.Linear ('weight') or nn.Sequential mlpNx_gelu ('0.weight', '2.weight', ...)."""
    out = {}
    if "weight" in sd:
        out[f"{prefix}.0.w"] = _dev(sd["weight"], device)
        out[f"{prefix}.0.b"] = _dev(sd["bias"], device)
        return out
    idx = sorted({int(k.split(".")[0]) for k in sd})
    for n, i in enumerate(idx):
        out[f"{prefix}.{n}.w"] = _dev(sd[f"{i}.weight"], device)
        out[f"{prefix}.{n}.b"] = _dev(sd[f"{i}.bias"], device)
    return out


def prepare_llm(sd: Dict[str, torch.Tensor], cfg: dict, device, decode_copies: bool = True) -> Dict[str, torch.Tensor]:
    """sd keys: 'embed_tokens.weight', 'layers.N.*', 'norm.weight', 'lm_head.weight' (optional when tied)."""
    out = {"llm.embed": _dev(sd["embed_tokens.weight"], device), "llm.norm.w": _dev(sd["norm.weight"], device)}
    if "lm_head.weight" in sd:
        out["llm.lm_head"] = _dev(sd["lm_head.weight"], device)
    elif cfg.get("tie_word_embeddings", False):
        out["llm.lm_head"] = out["llm.embed"]                     # tied head (configuration_qwen2_5_vl.py: tie_word_embeddings)
    else:
        raise KeyError("lm_head.weight is missing and tie_word_embeddings is false: refusing to decode with the embedding table")
    for i in range(cfg["num_hidden_layers"]):
        s, d = f"layers.{i}.", f"llm.l{i}."
        out[d + "ln1.w"] = _dev(sd[s + "input_layernorm.weight"], device)
        out[d + "qkv.w"] = torch.cat([_dev(sd[s + f"self_attn.{n}_proj.weight"], device) for n in "qkv"], 0).contiguous()
        out[d + "qkv.b"] = torch.cat([_dev(sd[s + f"self_attn.{n}_proj.bias"], device) for n in "qkv"], 0).contiguous()
        out[d + "o.w"] = _dev(sd[s + "self_attn.o_proj.weight"], device)
        out[d + "ln2.w"] = _dev(sd[s + "post_attention_layernorm.weight"], device)
        out[d + "gateup.w"] = interleave_gate_up(_dev(sd[s + "mlp.gate_proj.weight"], device), _dev(sd[s + "mlp.up_proj.weight"], device))
        out[d + "down.w"] = pad_cols(_dev(sd[s + "mlp.down_proj.weight"], device), vit_inter_pad(cfg["intermediate_size"]))
    if decode_copies and cfg["intermediate_size"] % 64 == 0 and cfg["hidden_size"] % 64 == 0:
        # persistent decode kernel (csrc/decode_mega.cu): RMSNorm folded into the consumer, y = rsqrt(mean x^2 + eps) * (W diag(g)) x;
        # gate / up rows interleaved 8 + 8 so that a 16-row tile yields 8 finished SwiGLU columns
        for i in range(cfg["num_hidden_layers"]):
            s, d = f"layers.{i}.", f"llm.l{i}."
            g1 = _dev(sd[s + "input_layernorm.weight"], device).float()
            g2 = _dev(sd[s + "post_attention_layernorm.weight"], device).float()
            out[d + "qkv_dec.w"] = (out[d + "qkv.w"].float() * g1[None, :]).to(torch.bfloat16).contiguous()
            gate = (_dev(sd[s + "mlp.gate_proj.weight"], device).float() * g2[None, :]).to(torch.bfloat16)
            up = (_dev(sd[s + "mlp.up_proj.weight"], device).float() * g2[None, :]).to(torch.bfloat16)
            out[d + "gu_dec.w"] = interleave_gate_up(gate, up, block=8)
        out["llm.head_dec"] = (out["llm.lm_head"].float() * out["llm.norm.w"].float()[None, :]).to(torch.bfloat16).contiguous()
    return out
