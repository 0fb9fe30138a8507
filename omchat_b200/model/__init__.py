"""`from omchat_b200.model import OmChatQwen2ForCausalLM, OmChatQwen2Config` — the import the reference's callers use
(omchat/model/__init__.py:2, omchat/model/builder.py:20)."""
from ..config import InternVisionConfig, OmChatQwen2Config, OmChatQwen2MoeConfig  # noqa: F401
from .omchat import OmChatForConditionalGeneration, OmChatQwen2ForCausalLM, OmChatQwen2Model  # noqa: F401
from .moe import OmChatQwen2MoeForCausalLM, OmChatQwen2MoeModel  # noqa: F401
from .vision import InternVIT300mVisionTower, InternVITVisionTower  # noqa: F401
