"""Configuration objects mirroring the reference's config surface.

`OmChatQwen2Config` keeps the attribute names the reference reads with getattr on the hot path (omchat_arch.py:25-28,
100,161,176; internVIT_encoder.py:16-17) and adds three knobs whose defaults reproduce reference behaviour:
`mm_pixel_shuffle_ratio` (1.0 = none), `kv_page_size`, `tp_size`.
"""
from __future__ import annotations

from dataclasses import dataclass, field, asdict
from typing import List, Optional

IMAGE_TOKEN_INDEX = -200  # omchat/constants.py:8
IGNORE_INDEX = -100


@dataclass
class InternVisionConfig:
    """intern_vit_6b/configuration_intern_vit.py:63-83 defaults (InternViT-6B-448px-V1-5)."""
    num_channels: int = 3
    patch_size: int = 14
    image_size: int = 448
    qkv_bias: bool = False
    hidden_size: int = 3200
    num_attention_heads: int = 25
    intermediate_size: int = 12800
    qk_normalization: bool = True
    num_hidden_layers: int = 45
    hidden_act: str = "gelu"
    layer_norm_eps: float = 1e-6
    initializer_factor: float = 0.1
    norm_type: str = "rms_norm"  # 'layer_norm' for the InternViT-300M variant (intern_vit_300m/modeling_intern_vit.py:61-64)

    def __post_init__(self):
        if self.norm_type not in ("rms_norm", "layer_norm"):
            raise ValueError(f"unknown norm_type {self.norm_type!r}")

    @classmethod
    def intern_vit_300m(cls, **overrides) -> "InternVisionConfig":
        """intern_vit_300m/configuration_intern_vit.py:60-80 defaults: the lighter tower (LayerNorm, no QK-norm, 16 heads of 64)."""
        kw = dict(hidden_size=1024, num_attention_heads=16, intermediate_size=4096, qk_normalization=False, num_hidden_layers=24,
                  norm_type="layer_norm", initializer_factor=1.0)
        kw.update(overrides)
        return cls(**kw)

    @property
    def num_patches(self) -> int:
        return (self.image_size // self.patch_size) ** 2

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads


@dataclass
class OmChatQwen2Config:
    """omchat_qwen2.py:16-19 (Qwen2Config fields + the mm_* attributes); defaults = Qwen2-7B + InternViT-6B."""
    model_type: str = "omchat_qwen2"
    vocab_size: int = 152064
    hidden_size: int = 3584
    intermediate_size: int = 18944
    num_hidden_layers: int = 28
    num_attention_heads: int = 28
    num_key_value_heads: int = 4
    rms_norm_eps: float = 1e-6
    rope_theta: float = 1e6
    max_position_embeddings: int = 32768
    # multimodal attributes read by the reference
    mm_vision_tower: Optional[str] = "InternViT-6B-448px-V1-5"
    mm_projector_type: str = "mlp2x_gelu"
    mm_hidden_size: int = 3200
    mm_vision_select_layer: int = -1
    mm_vision_select_feature: str = "patch"
    delay_load: bool = False
    image_grid_pinpoints: List[List[int]] = field(default_factory=lambda: [
        [448, 896], [896, 448], [896, 896], [1344, 448], [448, 1344], [1344, 1344]])
    tokenizer_model_max_length: Optional[int] = None
    tokenizer_padding_side: str = "right"
    tune_mm_mlp_adapter: bool = False
    mm_use_im_start_end: bool = False
    eos_token_id: int = 151645
    pad_token_id: int = 151643
    # build-specific knobs (defaults reproduce the reference)
    mm_pixel_shuffle_ratio: float = 1.0
    kv_page_size: int = 64
    tp_size: int = 1
    vision_config: InternVisionConfig = field(default_factory=InternVisionConfig)

    def __post_init__(self):
        # build_vision_tower (multimodal_encoder/builder.py:11-14) picks the tower class from the NAME: a config that names the
        # 300M tower and leaves vision_config at the 6B defaults gets the 300M defaults
        name = (self.mm_vision_tower or "").lower()
        if "internvit-300m" in name and self.vision_config == InternVisionConfig():
            self.vision_config = InternVisionConfig.intern_vit_300m()
            if self.mm_hidden_size == 3200:
                self.mm_hidden_size = self.vision_config.hidden_size

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads

    @property
    def pixel_shuffle_down(self) -> int:
        d = round(1.0 / self.mm_pixel_shuffle_ratio)
        if d < 1 or abs(1.0 / d - self.mm_pixel_shuffle_ratio) > 1e-6:
            raise ValueError(f"mm_pixel_shuffle_ratio must be 1/k, got {self.mm_pixel_shuffle_ratio}")
        return d

    @property
    def image_tokens_per_crop(self) -> int:
        g = self.vision_config.image_size // self.vision_config.patch_size
        return (g // self.pixel_shuffle_down) ** 2

    def to_dict(self):
        return asdict(self)


@dataclass
class OmChatQwen2MoeConfig(OmChatQwen2Config):
    """omchat_qwen2_moe.py:14-17: transformers Qwen2MoeConfig fields + the mm_* attributes. Defaults = Qwen1.5-MoE-A2.7B's
    language model (Qwen2MoeConfig defaults) under the same vision tower."""
    model_type: str = "omchat_qwen2_moe"
    hidden_size: int = 2048
    intermediate_size: int = 5632
    num_hidden_layers: int = 24
    num_attention_heads: int = 16
    num_key_value_heads: int = 16
    num_experts: int = 60
    num_experts_per_tok: int = 4
    moe_intermediate_size: int = 1408
    shared_expert_intermediate_size: int = 5632
    norm_topk_prob: bool = False
    decoder_sparse_step: int = 1
    mlp_only_layers: List[int] = field(default_factory=list)

    def layer_is_sparse(self, li: int) -> bool:
        """Qwen2MoeDecoderLayer.__init__ (transformers modeling_qwen2_moe.py:381-386)."""
        return li not in self.mlp_only_layers and self.num_experts > 0 and (li + 1) % self.decoder_sparse_step == 0
