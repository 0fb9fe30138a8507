"""Checkpoint loading for the drop-in boundary.

The reference loads with `OmChatQwen2ForCausalLM.from_pretrained(model_path, torch_dtype=float16, device_map="auto")`
(omchat/model/builder.py:22-35) from HF safetensors shards, in one of two parameter-name layouts:
  * omchat layout  — `model.vision_tower.vision_tower.*`, `model.mm_projector.{0,2}.*`, `model.layers.*`, `lm_head.*`
  * HF-hub layout  — `vision_tower.*`, `multi_modal_projector.linear_{1,2}.*`, `language_model.model.*`,
    `language_model.lm_head.*` (what convert_omchat_to_hf.py:26-35,47-59 writes)
Both are accepted; `weights.from_state_dict` normalises the names. Tensors are read shard by shard with safetensors
(fp16 checkpoints are converted to bf16 when they are copied to the device, weights.py `_dev`), `*.inv_freq` buffers are
skipped like convert_omchat_to_hf.py:51-52 does. `config.json` is mapped onto OmChatQwen2Config: the omchat layout is
a flat Qwen2 config plus mm_* attributes (omchat_qwen2.py:16-19), the hub layout nests `text_config` /
`vision_config` (omchat/hf/configuration_omchat.py:99-198).
"""
from __future__ import annotations

import glob
import json
import os
from dataclasses import fields
from typing import Dict, Optional, Tuple

import torch

from ..config import InternVisionConfig, OmChatQwen2Config, OmChatQwen2MoeConfig

_TEXT_KEYS = ("vocab_size", "hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads",
              "num_key_value_heads", "rms_norm_eps", "rope_theta", "max_position_embeddings")
_MM_KEYS = ("mm_vision_tower", "mm_projector_type", "mm_hidden_size", "mm_vision_select_layer",
            "mm_vision_select_feature", "image_grid_pinpoints", "tokenizer_model_max_length", "tokenizer_padding_side",
            "tune_mm_mlp_adapter", "mm_use_im_start_end", "eos_token_id", "pad_token_id", "mm_pixel_shuffle_ratio",
            "kv_page_size")
_MOE_KEYS = ("num_experts", "num_experts_per_tok", "moe_intermediate_size", "shared_expert_intermediate_size", "norm_topk_prob",
             "decoder_sparse_step", "mlp_only_layers")


def config_from_dict(d: dict) -> OmChatQwen2Config:
    """config.json (either layout) -> OmChatQwen2Config; unknown keys are ignored, missing keys keep the defaults."""
    kw = {}
    text = d.get("text_config") or d
    for k in _TEXT_KEYS:
        if text.get(k) is not None:
            kw[k] = text[k]
    for k in _MM_KEYS:
        if d.get(k) is not None:
            kw[k] = d[k]
    if "vision_feature_layer" in d and "mm_vision_select_layer" not in kw:  # hub layout (hf/configuration_omchat.py)
        kw["mm_vision_select_layer"] = d["vision_feature_layer"]
    if isinstance(kw.get("eos_token_id"), (list, tuple)):
        kw["eos_token_id"] = kw["eos_token_id"][0]
    vd = d.get("vision_config")
    if isinstance(vd, dict):
        names = {f.name for f in fields(InternVisionConfig)}
        kw["vision_config"] = InternVisionConfig(**{k: v for k, v in vd.items() if k in names})
        kw.setdefault("mm_hidden_size", kw["vision_config"].hidden_size)
    if d.get("model_type") == "omchat_qwen2_moe" or text.get("model_type") in ("omchat_qwen2_moe", "qwen2_moe"):
        for k in _MOE_KEYS:  # omchat_qwen2_moe.py:14-17 (transformers Qwen2MoeConfig)
            if text.get(k) is not None:
                kw[k] = text[k]
        return OmChatQwen2MoeConfig(**kw)
    return OmChatQwen2Config(**kw)


def load_state_dict(path: str) -> Dict[str, torch.Tensor]:
    """All tensors of every *.safetensors shard under `path` (or of the single file `path`), on the host."""
    from safetensors import safe_open
    files = [path] if os.path.isfile(path) else sorted(glob.glob(os.path.join(path, "*.safetensors")))
    if not files:
        raise FileNotFoundError(f"no *.safetensors under {path}")
    sd: Dict[str, torch.Tensor] = {}
    for f in files:
        with safe_open(f, framework="pt", device="cpu") as sf:
            for k in sf.keys():
                if k.endswith(".inv_freq"):
                    continue
                sd[k] = sf.get_tensor(k)
    return sd


def load_checkpoint(path: str, config: Optional[OmChatQwen2Config] = None) -> Tuple[Dict[str, torch.Tensor], OmChatQwen2Config]:
    if config is None:
        cj = os.path.join(path, "config.json") if os.path.isdir(path) else os.path.join(os.path.dirname(path), "config.json")
        if os.path.exists(cj):
            with open(cj) as fh:
                config = config_from_dict(json.load(fh))
        else:
            config = OmChatQwen2Config()
    return load_state_dict(path), config


def save_checkpoint(sd: Dict[str, torch.Tensor], config: OmChatQwen2Config, path: str, hub_layout: bool = False,
                    max_shard_bytes: int = 4 << 30) -> None:
    """Writes `sd` (omchat names) as safetensors shards + config.json; hub_layout=True renames to the HF-hub layout the
    way convert_omchat_to_hf.py does. Used by the round-trip tests and to export random-init benchmark weights."""
    from safetensors.torch import save_file
    from .weights import KEYS_TO_MODIFY_MAPPING
    os.makedirs(path, exist_ok=True)
    out = {}
    for k, v in sd.items():
        if hub_layout:
            for a, b in KEYS_TO_MODIFY_MAPPING.items():
                if k.startswith(a):
                    k = b + k[len(a):]
                    break
        out[k] = v.detach().cpu().contiguous()
    shards, cur, size = [], {}, 0
    for k, v in out.items():
        n = v.numel() * v.element_size()
        if cur and size + n > max_shard_bytes:
            shards.append(cur)
            cur, size = {}, 0
        cur[k] = v
        size += n
    if cur:
        shards.append(cur)
    for i, s in enumerate(shards):
        save_file(s, os.path.join(path, f"model-{i + 1:05d}-of-{len(shards):05d}.safetensors"))
    d = config.to_dict()
    if hub_layout:  # omchat/hf/configuration_omchat.py:99-198: model_type "omchat", nested text / vision configs
        text = {k: d[k] for k in _TEXT_KEYS}
        d = {k: v for k, v in d.items() if k not in _TEXT_KEYS}
        d["text_config"] = text
        d["model_type"] = "omchat"
        d["vision_feature_layer"] = d["mm_vision_select_layer"]
        d["image_token_index"] = -200
    with open(os.path.join(path, "config.json"), "w") as fh:
        json.dump(d, fh, indent=1)
