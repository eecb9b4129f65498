"""Random-init checkpoints (no network: real weights are unavailable offline) and engine loading.

``random_state_dicts`` follows the reference initialisers -- N(0, 0.02^2) for Linear / Conv / Embedding weights,
zero biases, unit norm gains (modeling_qwen2_5_vl.py:392-401, modeling_davit.py:435-450) -- and stores bf16,
keyed by the reference checkpoint's tensor names (component prefixes stripped)."""
from __future__ import annotations

from typing import Dict

import torch

from . import weights as W
from .engine import Engine, EngineConfig


def _rn(shape, g, device, std=0.02):
    return (torch.randn(shape, generator=g, device=device, dtype=torch.float32) * std).to(torch.bfloat16)


def _ones(n, device):
    return torch.ones(n, dtype=torch.bfloat16, device=device)


def _zeros(n, device):
    return torch.zeros(n, dtype=torch.bfloat16, device=device)


def random_vit(cfg: dict, g, device) -> Dict[str, torch.Tensor]:
    H, I, O = cfg["hidden_size"], cfg["intermediate_size"], cfg["out_hidden_size"]
    pk = cfg.get("in_channels", 3) * cfg["temporal_patch_size"] * cfg["patch_size"] ** 2
    sd = {"patch_embed.proj.weight": _rn((H, pk), g, device)}
    for i in range(cfg["depth"]):
        p = f"blocks.{i}."
        sd[p + "norm1.weight"] = _ones(H, device); sd[p + "norm2.weight"] = _ones(H, device)
        sd[p + "attn.qkv.weight"] = _rn((3 * H, H), g, device); sd[p + "attn.qkv.bias"] = _zeros(3 * H, device)
        sd[p + "attn.proj.weight"] = _rn((H, H), g, device); sd[p + "attn.proj.bias"] = _zeros(H, device)
        for n in ("gate_proj", "up_proj"):
            sd[p + f"mlp.{n}.weight"] = _rn((I, H), g, device); sd[p + f"mlp.{n}.bias"] = _zeros(I, device)
        sd[p + "mlp.down_proj.weight"] = _rn((H, I), g, device); sd[p + "mlp.down_proj.bias"] = _zeros(H, device)
    M4 = H * cfg["spatial_merge_size"] ** 2
    sd["merger.ln_q.weight"] = _ones(H, device)
    sd["merger.mlp.0.weight"] = _rn((M4, M4), g, device); sd["merger.mlp.0.bias"] = _zeros(M4, device)
    sd["merger.mlp.2.weight"] = _rn((O, M4), g, device); sd["merger.mlp.2.bias"] = _zeros(O, device)
    return sd


def random_davit(cfg: dict, g, device) -> Dict[str, torch.Tensor]:
    sd = {}
    dims = cfg["dim_embed"]
    for s in range(4):
        C, Cin = dims[s], (3 if s == 0 else dims[s - 1])
        k = 7 if s == 0 else 3
        p = f"convs.{s}."
        sd[p + "proj.weight"] = _rn((C, Cin, k, k), g, device); sd[p + "proj.bias"] = _zeros(C, device)
        nd = C if s == 0 else Cin
        sd[p + "norm.weight"] = _ones(nd, device); sd[p + "norm.bias"] = _zeros(nd, device)
        for j in range(cfg["depths"][s]):
            for kind, attn in (("spatial_block", "window_attn"), ("channel_block", "channel_attn")):
                q = f"blocks.{s}.{j}.{kind}."
                for c in ("conv1", "conv2"):
                    sd[q + c + ".fn.dw.weight"] = _rn((C, 1, 3, 3), g, device); sd[q + c + ".fn.dw.bias"] = _zeros(C, device)
                sd[q + attn + ".norm.weight"] = _ones(C, device); sd[q + attn + ".norm.bias"] = _zeros(C, device)
                sd[q + attn + ".fn.qkv.weight"] = _rn((3 * C, C), g, device); sd[q + attn + ".fn.qkv.bias"] = _zeros(3 * C, device)
                sd[q + attn + ".fn.proj.weight"] = _rn((C, C), g, device); sd[q + attn + ".fn.proj.bias"] = _zeros(C, device)
                sd[q + "ffn.norm.weight"] = _ones(C, device); sd[q + "ffn.norm.bias"] = _zeros(C, device)
                sd[q + "ffn.fn.net.fc1.weight"] = _rn((4 * C, C), g, device); sd[q + "ffn.fn.net.fc1.bias"] = _zeros(4 * C, device)
                sd[q + "ffn.fn.net.fc2.weight"] = _rn((C, 4 * C), g, device); sd[q + "ffn.fn.net.fc2.bias"] = _zeros(C, device)
    return sd


def random_fpn(dim: int, out: int, g, device) -> Dict[str, torch.Tensor]:
    sd = {}
    sd["simfp_1.0.weight"] = _rn((dim, dim // 2, 2, 2), g, device); sd["simfp_1.0.bias"] = _zeros(dim // 2, device)
    sd["simfp_1.1.weight"] = _ones(dim // 2, device); sd["simfp_1.1.bias"] = _zeros(dim // 2, device)
    sd["simfp_1.3.weight"] = _rn((dim // 2, dim // 4, 2, 2), g, device); sd["simfp_1.3.bias"] = _zeros(dim // 4, device)
    sd["simfp_2.0.weight"] = _rn((dim, dim // 2, 2, 2), g, device); sd["simfp_2.0.bias"] = _zeros(dim // 2, device)
    for stage, i, cin in ((1, 4, dim // 4), (2, 1, dim // 2), (3, 0, dim), (4, 1, dim)):
        p = f"simfp_{stage}."
        sd[p + f"{i}.weight"] = _rn((out, cin, 1, 1), g, device)
        sd[p + f"{i}.norm.weight"] = _ones(out, device); sd[p + f"{i}.norm.bias"] = _zeros(out, device)
        sd[p + f"{i + 1}.weight"] = _rn((out, out, 3, 3), g, device)
        sd[p + f"{i + 1}.norm.weight"] = _ones(out, device); sd[p + f"{i + 1}.norm.bias"] = _zeros(out, device)
    return sd


def random_projector(in_dim: int, hidden: int, layers: int, g, device) -> Dict[str, torch.Tensor]:
    sd = {}
    for k in range(layers):
        sd[f"{2 * k}.weight"] = _rn((hidden, in_dim if k == 0 else hidden), g, device)
        sd[f"{2 * k}.bias"] = _zeros(hidden, device)
    return sd


def random_llm(cfg: dict, g, device) -> Dict[str, torch.Tensor]:
    H, I, V = cfg["hidden_size"], cfg["intermediate_size"], cfg["vocab_size"]
    hd = H // cfg["num_attention_heads"]
    KD = cfg["num_key_value_heads"] * hd
    sd = {"embed_tokens.weight": _rn((V, H), g, device), "norm.weight": _ones(H, device)}
    if not cfg.get("tie_word_embeddings", False):
        sd["lm_head.weight"] = _rn((V, H), g, device)
    for i in range(cfg["num_hidden_layers"]):
        p = f"layers.{i}."
        sd[p + "input_layernorm.weight"] = _ones(H, device); sd[p + "post_attention_layernorm.weight"] = _ones(H, device)
        sd[p + "self_attn.q_proj.weight"] = _rn((H, H), g, device); sd[p + "self_attn.q_proj.bias"] = _zeros(H, device)
        sd[p + "self_attn.k_proj.weight"] = _rn((KD, H), g, device); sd[p + "self_attn.k_proj.bias"] = _zeros(KD, device)
        sd[p + "self_attn.v_proj.weight"] = _rn((KD, H), g, device); sd[p + "self_attn.v_proj.bias"] = _zeros(KD, device)
        sd[p + "self_attn.o_proj.weight"] = _rn((H, H), g, device)
        sd[p + "mlp.gate_proj.weight"] = _rn((I, H), g, device); sd[p + "mlp.up_proj.weight"] = _rn((I, H), g, device)
        sd[p + "mlp.down_proj.weight"] = _rn((H, I), g, device)
    return sd


def random_state_dicts(cfg: EngineConfig, device, seed: int = 0) -> Dict[str, Dict[str, torch.Tensor]]:
    g = torch.Generator(device=device).manual_seed(seed)
    out = {}
    if cfg.use_vit:
        out["vit"] = random_vit(cfg.vit, g, device)
        if cfg.fpn_out:
            out["fpn"] = random_fpn(cfg.vit["hidden_size"], cfg.fpn_out, g, device)
    if cfg.use_davit:
        out["davit"] = random_davit(cfg.davit, g, device)
    if cfg.proj_aux_layers:
        out["proj_aux"] = random_projector(cfg.region_dim, cfg.llm["hidden_size"], cfg.proj_aux_layers, g, device)
    if cfg.proj_img_layers and cfg.use_vit:
        out["proj_img"] = random_projector(cfg.vit["out_hidden_size"], cfg.llm["hidden_size"], cfg.proj_img_layers, g, device)
    if cfg.use_llm:
        out["llm"] = random_llm(cfg.llm, g, device)
    return out


def load_engine(cfg: EngineConfig, sds: Dict[str, Dict[str, torch.Tensor]], device=None) -> Engine:
    """state_dicts by component (reference tensor names) -> a finalized Engine."""
    eng = Engine(cfg, device)
    dev = eng.device
    if "vit" in sds:
        eng.set_weights(W.prepare_vit(sds["vit"], cfg.vit, dev))
    if "fpn" in sds:
        eng.set_weights(W.prepare_fpn(sds["fpn"], dev))
    if "davit" in sds:
        eng.set_weights(W.prepare_davit(sds["davit"], cfg.davit, dev))
    if "proj_aux" in sds:
        eng.set_weights(W.prepare_projector(sds["proj_aux"], "proj_aux", dev))
    if "proj_img" in sds:
        eng.set_weights(W.prepare_projector(sds["proj_img"], "proj_img", dev))
    if "llm" in sds:
        eng.set_weights(W.prepare_llm(sds["llm"], cfg.llm, dev))
    eng.finalize()
    return eng
