"""The reference's Hugging Face `Auto*` surface on the sm_100a path.

Importing this module does what importing `omchat.model` / loading the hub repo with `trust_remote_code` does in the
reference:

  * `AutoConfig.register("omchat_qwen2", OmChatQwen2Config)` and
    `AutoModelForCausalLM.register(OmChatQwen2Config, OmChatQwen2ForCausalLM)`
    (omchat/model/language_model/omchat_qwen2.py:113-114) — the cli.py / builder.py:22-35 loading path;
  * the hub twin of hf_example.py:7-18: `AutoModel.from_pretrained(dir)` -> `OmChatForConditionalGeneration`
    (omchat/hf/modeling_omchat.py:677-689) and `AutoProcessor.from_pretrained(dir)` -> `OmChatProcessor`
    (omchat/hf/processing_omchat.py:142-246), keyed by `OmChatConfig` (model_type "omchat",
    omchat/hf/configuration_omchat.py:99-198).

A checkpoint directory whose config.json carries an `auto_map` (the published hub layout) resolves to these local
classes as well: transformers prefers explicitly registered local code over remote code, and
`install_remote_code(dir)` writes the three `*_omchat.py` shims a hub repo needs for hosts that have not imported this
module (`trust_remote_code=True` then imports omchat_b200 through them).

The HF config classes only carry attributes; `to_native()` turns them into the dataclass config the kernels' host code
uses (omchat_b200/config.py). The model classes are the native ones with `config_class` pointing at the HF config and a
`from_pretrained` that accepts what `_BaseAutoModelClass.from_pretrained` passes.
"""
from __future__ import annotations

import json
import os
from typing import Optional

import torch
from transformers import AutoConfig, AutoModel, AutoModelForCausalLM, AutoProcessor, PretrainedConfig

from .config import InternVisionConfig as NativeVisionConfig
from .config import OmChatQwen2Config as NativeConfig
from .config import OmChatQwen2MoeConfig as NativeMoeConfig
from .model import moe as native_moe
from .model import omchat as native
from .model.checkpoint import config_from_dict
from .processing import OmChatImageProcessor, OmChatProcessor

_NATIVE_DEFAULTS = NativeConfig().to_dict()


class OmChatQwen2Config(PretrainedConfig):
    """`OmChatQwen2Config(Qwen2Config)` of omchat_qwen2.py:16-19: a flat Qwen2 config plus the mm_* attributes the glue
    reads with getattr (omchat_arch.py:25-28,100,161,176)."""
    model_type = "omchat_qwen2"
    rotary_type = "normal_rotary"
    multi_scale_im = None

    def __init__(self, **kwargs):
        for k, v in _NATIVE_DEFAULTS.items():
            if k in ("model_type", "vision_config"):
                continue
            setattr(self, k, kwargs.pop(k, v))
        vc = kwargs.pop("vision_config", None)
        self.vision_config = vc.__dict__.copy() if isinstance(vc, NativeVisionConfig) else vc
        super().__init__(**kwargs)

    def to_native(self) -> NativeConfig:
        return config_from_dict(self.to_dict())


class OmChatQwen2MoeConfig(OmChatQwen2Config):
    """`OmChatQwen2MoeConfig(Qwen2MoeConfig)` of omchat_qwen2_moe.py:14-17: the flat config + the mixture-of-experts fields."""
    model_type = "omchat_qwen2_moe"

    def __init__(self, **kwargs):
        base, moe = NativeConfig().to_dict(), NativeMoeConfig().to_dict()
        for k, v in moe.items():
            if k not in ("model_type", "vision_config") and (k not in base or base[k] != v):
                kwargs.setdefault(k, v)  # the MoE defaults (Qwen1.5-MoE-A2.7B sizes) where they differ from the dense model's
        extra = {k: kwargs.pop(k) for k in list(kwargs) if k in moe and k not in base}
        super().__init__(**kwargs)
        for k, v in extra.items():
            setattr(self, k, v)


class OmChatConfig(PretrainedConfig):
    """Hub-layout config (hf/configuration_omchat.py:99-198): nested `vision_config` / `text_config`, `image_token_index`,
    `vision_feature_layer`, `image_grid_pinpoints`."""
    model_type = "omchat"

    def __init__(self, vision_config=None, text_config=None, ignore_index=-100, image_token_index=-200,
                 projector_hidden_act="gelu", vision_feature_select_strategy="default", vision_feature_layer=-1,
                 image_grid_pinpoints=None, tie_word_embeddings=False, **kwargs):
        if vision_feature_select_strategy not in ("default", "full"):
            raise ValueError(f"vision_feature_select_strategy should be one of 'default', 'full'. Got: {vision_feature_select_strategy}")
        self.ignore_index = ignore_index
        self.image_token_index = image_token_index
        self.projector_hidden_act = projector_hidden_act
        self.vision_feature_select_strategy = vision_feature_select_strategy
        self.vision_feature_layer = vision_feature_layer
        self.image_grid_pinpoints = image_grid_pinpoints if image_grid_pinpoints is not None else \
            [[448, 896], [896, 448], [896, 896], [1344, 448], [448, 1344]]
        self.vision_config = dict(vision_config) if vision_config is not None else NativeVisionConfig().__dict__.copy()
        self.text_config = dict(text_config) if text_config is not None else \
            {k: _NATIVE_DEFAULTS[k] for k in ("vocab_size", "hidden_size", "intermediate_size", "num_hidden_layers",
                                              "num_attention_heads", "num_key_value_heads", "rms_norm_eps", "rope_theta",
                                              "max_position_embeddings")}
        super().__init__(tie_word_embeddings=tie_word_embeddings, **kwargs)

    def to_native(self) -> NativeConfig:
        return config_from_dict(self.to_dict())


def _native_config(config) -> Optional[NativeConfig]:
    if config is None or isinstance(config, NativeConfig):
        return config
    if hasattr(config, "to_native"):
        return config.to_native()
    if isinstance(config, PretrainedConfig):
        return config_from_dict(config.to_dict())
    raise TypeError(f"unsupported config object {type(config)}")


_MODEL_KW = ("seed", "tp_rank", "tp_size", "tp_group")


class _AutoLoadable:
    """from_pretrained as `_BaseAutoModelClass.from_pretrained` calls it: positional model args, config=<PretrainedConfig>,
    hub / dtype / trust_remote_code keywords that have no meaning here (weights are bf16 on the device by construction)."""

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, *model_args, config=None, device=None, **kwargs):
        kw = {k: kwargs[k] for k in _MODEL_KW if k in kwargs}
        dev = device or (kwargs.get("device_map") if isinstance(kwargs.get("device_map"), (str, torch.device)) and
                         str(kwargs.get("device_map")) not in ("auto", "balanced", "sequential") else None) or "cuda"
        model = super().from_pretrained(str(pretrained_model_name_or_path), config=_native_config(config), device=dev, **kw)
        model.hf_config = config
        return model

    # no-ops the reference's callers chain after from_pretrained (hf_example.py:7): the model already lives on its GPU
    def cuda(self, *a, **k):
        return self

    def half(self):
        return self


class OmChatQwen2ForCausalLM(_AutoLoadable, native.OmChatQwen2ForCausalLM):
    config_class = OmChatQwen2Config


class OmChatQwen2MoeForCausalLM(_AutoLoadable, native_moe.OmChatQwen2MoeForCausalLM):
    config_class = OmChatQwen2MoeConfig


class OmChatForConditionalGeneration(_AutoLoadable, native.OmChatForConditionalGeneration):
    config_class = OmChatConfig


def _processor_from_pretrained(cls, pretrained_model_name_or_path, **kwargs):
    """`AutoProcessor.from_pretrained(dir)` -> OmChatProcessor: tokenizer through AutoTokenizer, image processor from
    preprocessor_config.json (image_grid_pinpoints, crop size, mean / std) when present."""
    from transformers import AutoTokenizer
    path = str(pretrained_model_name_or_path)
    tok = kwargs.pop("tokenizer", None) or AutoTokenizer.from_pretrained(path)
    ip_kw = {}
    pj = os.path.join(path, "preprocessor_config.json")
    if os.path.exists(pj):
        with open(pj) as fh:
            d = json.load(fh)
        if d.get("image_grid_pinpoints") is not None:
            ip_kw["image_grid_pinpoints"] = d["image_grid_pinpoints"]
        size = d.get("crop_size") or d.get("size")
        if isinstance(size, dict):
            ip_kw["size"] = int(size.get("height") or size.get("shortest_edge"))
        if d.get("image_mean") is not None:
            ip_kw["image_mean"] = tuple(d["image_mean"])
        if d.get("image_std") is not None:
            ip_kw["image_std"] = tuple(d["image_std"])
    dev = kwargs.pop("device", None) or ("cuda" if torch.cuda.is_available() else "cpu")
    return cls(image_processor=OmChatImageProcessor(device=dev, **ip_kw), tokenizer=tok)


OmChatProcessor.from_pretrained = classmethod(_processor_from_pretrained)


def register_auto_classes() -> None:
    AutoConfig.register("omchat_qwen2", OmChatQwen2Config, exist_ok=True)
    AutoModelForCausalLM.register(OmChatQwen2Config, OmChatQwen2ForCausalLM, exist_ok=True)
    # omchat_qwen2_moe.py:116-117
    AutoConfig.register("omchat_qwen2_moe", OmChatQwen2MoeConfig, exist_ok=True)
    AutoModelForCausalLM.register(OmChatQwen2MoeConfig, OmChatQwen2MoeForCausalLM, exist_ok=True)
    AutoConfig.register("omchat", OmChatConfig, exist_ok=True)
    AutoModel.register(OmChatConfig, OmChatForConditionalGeneration, exist_ok=True)
    AutoProcessor.register(OmChatConfig, OmChatProcessor, exist_ok=True)


register_auto_classes()

_SHIMS = {
    "configuration_omchat.py": "from omchat_b200.hf import OmChatConfig  # noqa: F401\n",
    "modeling_omchat.py": "from omchat_b200.hf import OmChatConfig, OmChatForConditionalGeneration  # noqa: F401\n",
    "processing_omchat.py": "from omchat_b200.hf import OmChatProcessor  # noqa: F401\n",
}
_AUTO_MAP = {"AutoConfig": "configuration_omchat.OmChatConfig", "AutoModel": "modeling_omchat.OmChatForConditionalGeneration",
             "AutoProcessor": "processing_omchat.OmChatProcessor"}


def install_remote_code(checkpoint_dir: str) -> None:
    """Make a hub-layout checkpoint directory load through this package with `trust_remote_code=True` (hf_example.py:7-8)
    even in a process that never imported omchat_b200.hf: writes the three module files config.json's `auto_map` names
    (each a one-line import of the class from here) and sets the `auto_map`."""
    for name, body in _SHIMS.items():
        with open(os.path.join(checkpoint_dir, name), "w") as fh:
            fh.write('"""Remote-code shim: the implementation lives in the installed omchat_b200 package."""\n' + body)
    cj = os.path.join(checkpoint_dir, "config.json")
    with open(cj) as fh:
        d = json.load(fh)
    d["auto_map"] = dict(_AUTO_MAP)
    d["model_type"] = "omchat"
    with open(cj, "w") as fh:
        json.dump(d, fh, indent=1)
