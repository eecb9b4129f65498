"""Fabricate a random-init checkpoint DIRECTORY in the layout ``vlm_fo1.model.builder.load_pretrained_model`` reads
(the released ``omlab/VLM-FO1_Qwen2.5-VL-3B-v01`` is not available offline): config.json with the mm_* flags,
safetensors shards with the reference's tensor names, a byte-level Qwen2 tokenizer whose special tokens sit at the real
ids (<|im_start|> 151644, <|im_end|> 151645, <|vision_start|> 151652 ...), generation_config.json."""
from __future__ import annotations

import json
import os
from typing import Dict

import torch

from .checkpoint import random_state_dicts
from .engine import EngineConfig

SPECIAL = ["<|endoftext|>", "<|im_start|>", "<|im_end|>", "<|object_ref_start|>", "<|object_ref_end|>", "<|box_start|>", "<|box_end|>",
           "<|quad_start|>", "<|quad_end|>", "<|vision_start|>", "<|vision_end|>", "<|vision_pad|>", "<|image_pad|>", "<|video_pad|>"]


def _bytes_to_unicode():
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("\xa1"), ord("\xac") + 1)) + list(range(ord("\xae"), ord("\xff") + 1))
    cs, n = bs[:], 0
    for b in range(256):
        if b not in bs:
            bs.append(b); cs.append(256 + n); n += 1
    return dict(zip(bs, [chr(c) for c in cs]))


def write_tokenizer(path: str, first_special: int = 151643) -> None:
    b2u = _bytes_to_unicode()
    vocab = {b2u[i]: i for i in range(256)}
    for n in range(256, first_special):
        vocab[f"<unused{n}>"] = n
    json.dump(vocab, open(os.path.join(path, "vocab.json"), "w"))
    open(os.path.join(path, "merges.txt"), "w").write("#version: 0.2\n")
    added = {str(first_special + i): {"content": t, "lstrip": False, "normalized": False, "rstrip": False, "single_word": False, "special": True}
             for i, t in enumerate(SPECIAL)}
    cfg = {"add_prefix_space": False, "added_tokens_decoder": added, "bos_token": None, "eos_token": "<|im_end|>", "pad_token": "<|endoftext|>",
           "unk_token": None, "tokenizer_class": "Qwen2Tokenizer", "model_max_length": 131072, "clean_up_tokenization_spaces": False,
           "errors": "replace", "split_special_tokens": False}
    json.dump(cfg, open(os.path.join(path, "tokenizer_config.json"), "w"))


def fabricate_checkpoint(path: str, cfg: EngineConfig = None, seed: int = 0, device="cpu", variant: str = "fpn") -> str:
    """Write the directory (its name must contain 'vlm-fo1' and 'qwen2.5-vl', builder.py:35,39).  Returns the path."""
    from safetensors.torch import save_file
    cfg = cfg or EngineConfig()
    os.makedirs(path, exist_ok=True)
    sds = random_state_dicts(cfg, device, seed)
    flat: Dict[str, torch.Tensor] = {}
    pref = {"vit": "model.vision_tower.image_tower.", "davit": "model.vision_tower_aux.image_tower.",
            "fpn": "model.object_vp_extractor.simple_fpn.", "proj_aux": "model.mm_projector_aux."}
    for comp, sd in sds.items():
        if comp == "llm":
            for k, v in sd.items():
                flat[k if k == "lm_head.weight" else "model." + k] = v
        else:
            for k, v in sd.items():
                flat[pref[comp] + k] = v
    shard, size, idx = {}, 0, 0
    for k, v in flat.items():
        shard[k] = v.detach().cpu().contiguous()
        size += v.numel() * v.element_size()
        if size > (4 << 30):
            save_file(shard, os.path.join(path, f"model-{idx:05d}.safetensors")); shard, size, idx = {}, 0, idx + 1
    if shard:
        save_file(shard, os.path.join(path, f"model-{idx:05d}.safetensors"))
    l, v = cfg.llm, cfg.vit
    config = {"model_type": "omchat_qwen2_5_vl", "architectures": ["OmChatQwen25VLForCausalLM"], "hidden_size": l["hidden_size"],
              "num_hidden_layers": l["num_hidden_layers"], "num_attention_heads": l["num_attention_heads"],
              "num_key_value_heads": l["num_key_value_heads"], "intermediate_size": l["intermediate_size"], "vocab_size": l["vocab_size"],
              "rope_theta": l["rope_theta"], "rms_norm_eps": l["rms_norm_eps"], "tie_word_embeddings": l["tie_word_embeddings"],
              "rope_scaling": {"type": "mrope", "mrope_section": l["mrope_section"]}, "max_position_embeddings": 128000,
              "image_token_id": 151655, "video_token_id": 151656, "vision_start_token_id": 151652, "vision_end_token_id": 151653,
              "bos_token_id": 151643, "eos_token_id": 151645, "pad_token_id": 151643,
              "vision_config": {"depth": v["depth"], "hidden_size": v["hidden_size"], "num_heads": v["num_heads"],
                                "intermediate_size": v["intermediate_size"], "out_hidden_size": v["out_hidden_size"], "patch_size": v["patch_size"],
                                "spatial_merge_size": v["spatial_merge_size"], "temporal_patch_size": v["temporal_patch_size"],
                                "window_size": v["window_size"], "fullatt_block_indexes": v["fullatt_block_indexes"], "in_chans": 3},
              "mm_vision_tower": "qwen2.5-vl-vit", "mm_vision_tower_aux": "davit-large" if cfg.davit["dim_embed"][0] == 256 else "davit-base",
              "mm_projector_type": "identity", "mm_projector_aux_type": f"mlp{cfg.proj_aux_layers}x_gelu" if cfg.proj_aux_layers > 1 else "linear",
              "mm_hidden_size": v["out_hidden_size"], "mm_region_hidden_size": cfg.region_dim, "mm_roi_output_size": 7,
              "mm_apply_position_embedding": True, "mm_pos_embedding_strategy": "bbox_based", "mm_use_vt_region_feature_only": False,
              "mm_use_vision_tower_region_feature": True, "mm_region_feature_combination": "concat", "mm_apply_region_layer_norm": False,
              "mm_use_simpleFPN_for_vt": variant == "fpn", "mm_use_region_index_token": True, "aux_image_size": 768,
              "aux_image_aspect_ratio": "dynamic", "davit_depths": cfg.davit["depths"]}
    json.dump(config, open(os.path.join(path, "config.json"), "w"), indent=1)
    json.dump({"eos_token_id": [151645, 151643], "pad_token_id": 151643, "do_sample": False}, open(os.path.join(path, "generation_config.json"), "w"))
    write_tokenizer(path)
    return path
