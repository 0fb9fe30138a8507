"""Deterministic tiny OmChat configuration + weights shared by the golden generator and the parity tests.

head_dim stays 128 (ViT 2 heads x 128, LLM 2 q heads / 1 kv head x 128) because the CUDA attention kernels are
specialised for the real models' head_dim; everything else is shrunk so the CPU oracle runs in seconds.
"""
from __future__ import annotations

import torch

TINY = dict(
    vit_hidden=256, vit_heads=2, vit_inter=512, vit_layers=2, image_size=224, patch_size=14,
    hidden=256, heads=2, kv_heads=1, inter=512, layers=2, vocab=1000, rope_theta=1e6,
)


def tiny_state_dict(seed: int = 0) -> dict:
    """Random weights under the reference's state-dict names. Norm / layer-scale weights are perturbed away from their
    init values (ones / 0.1) so that every parameter influences the outputs."""
    g = torch.Generator().manual_seed(seed)
    c = TINY
    sd = {}

    def rn(*shape, std=0.02, mean=0.0):
        return torch.randn(*shape, generator=g) * std + mean

    vt = "model.vision_tower.vision_tower."
    C, I = c["vit_hidden"], c["vit_inter"]
    npos = (c["image_size"] // c["patch_size"]) ** 2 + 1
    sd[vt + "embeddings.class_embedding"] = rn(1, 1, C, std=1.0)
    sd[vt + "embeddings.position_embedding"] = rn(1, npos, C, std=1.0)
    sd[vt + "embeddings.patch_embedding.weight"] = rn(C, 3, 14, 14, std=0.05)
    sd[vt + "embeddings.patch_embedding.bias"] = rn(C, std=0.1)
    for li in range(c["vit_layers"]):
        p = f"{vt}encoder.layers.{li}."
        sd[p + "ls1"] = rn(C, std=0.03, mean=0.1)
        sd[p + "ls2"] = rn(C, std=0.03, mean=0.1)
        sd[p + "attn.qkv.weight"] = rn(3 * C, C, std=0.05)
        sd[p + "attn.q_norm.weight"] = rn(C, std=0.1, mean=1.0)
        sd[p + "attn.k_norm.weight"] = rn(C, std=0.1, mean=1.0)
        sd[p + "attn.proj.weight"] = rn(C, C, std=0.05)
        sd[p + "attn.proj.bias"] = rn(C, std=0.1)
        sd[p + "mlp.fc1.weight"] = rn(I, C, std=0.05)
        sd[p + "mlp.fc1.bias"] = rn(I, std=0.1)
        sd[p + "mlp.fc2.weight"] = rn(C, I, std=0.05)
        sd[p + "mlp.fc2.bias"] = rn(C, std=0.1)
        sd[p + "norm1.weight"] = rn(C, std=0.1, mean=1.0)
        sd[p + "norm2.weight"] = rn(C, std=0.1, mean=1.0)
    H = c["hidden"]
    sd["model.mm_projector.0.weight"] = rn(H, C, std=0.05)
    sd["model.mm_projector.0.bias"] = rn(H, std=0.1)
    sd["model.mm_projector.2.weight"] = rn(H, H, std=0.05)
    sd["model.mm_projector.2.bias"] = rn(H, std=0.1)
    sd["model.embed_tokens.weight"] = rn(c["vocab"], H, std=0.5)
    D = H // c["heads"]
    for li in range(c["layers"]):
        p = f"model.layers.{li}."
        sd[p + "self_attn.q_proj.weight"] = rn(c["heads"] * D, H, std=0.05)
        sd[p + "self_attn.q_proj.bias"] = rn(c["heads"] * D, std=0.1)
        sd[p + "self_attn.k_proj.weight"] = rn(c["kv_heads"] * D, H, std=0.05)
        sd[p + "self_attn.k_proj.bias"] = rn(c["kv_heads"] * D, std=0.1)
        sd[p + "self_attn.v_proj.weight"] = rn(c["kv_heads"] * D, H, std=0.05)
        sd[p + "self_attn.v_proj.bias"] = rn(c["kv_heads"] * D, std=0.1)
        sd[p + "self_attn.o_proj.weight"] = rn(H, c["heads"] * D, std=0.05)
        sd[p + "mlp.gate_proj.weight"] = rn(c["inter"], H, std=0.05)
        sd[p + "mlp.up_proj.weight"] = rn(c["inter"], H, std=0.05)
        sd[p + "mlp.down_proj.weight"] = rn(H, c["inter"], std=0.05)
        sd[p + "input_layernorm.weight"] = rn(H, std=0.1, mean=1.0)
        sd[p + "post_attention_layernorm.weight"] = rn(H, std=0.1, mean=1.0)
    sd["model.norm.weight"] = rn(H, std=0.1, mean=1.0)
    sd["lm_head.weight"] = rn(c["vocab"], H, std=0.05)
    return sd


def weights_checksum(sd: dict) -> float:
    return float(sum(v.double().abs().sum() for v in sd.values()))


def tiny_inputs(seed: int = 1):
    g = torch.Generator().manual_seed(seed)
    pixels = torch.randn(4, 3, TINY["image_size"], TINY["image_size"], generator=g)
    ids = torch.randint(0, TINY["vocab"] - 10, (3, 24), generator=g)
    return pixels, ids


# ---- InternViT-300M variant (intern_vit_300m/): LayerNorm with bias, no QK-norm, 64-dim heads, qkv bias on (the published
# InternViT-300M-448px checkpoint has qkv_bias = true; the reference config default is false - the bias path is the superset)
TINY_300M = dict(TINY, vit_heads=4, vit_norm_type="layer_norm", vit_qk_norm=False, vit_qkv_bias=True)


def tiny_state_dict_300m(seed: int = 0) -> dict:
    sd = tiny_state_dict(seed)
    g = torch.Generator().manual_seed(seed + 1000)
    vt = "model.vision_tower.vision_tower."
    C = TINY_300M["vit_hidden"]
    for li in range(TINY_300M["vit_layers"]):
        p = f"{vt}encoder.layers.{li}."
        del sd[p + "attn.q_norm.weight"], sd[p + "attn.k_norm.weight"]
        sd[p + "norm1.bias"] = torch.randn(C, generator=g) * 0.1
        sd[p + "norm2.bias"] = torch.randn(C, generator=g) * 0.1
        sd[p + "attn.qkv.bias"] = torch.randn(3 * C, generator=g) * 0.1
    return sd


# ---- Qwen2-MoE language model variant (omchat/model/language_model/omchat_qwen2_moe.py over transformers Qwen2MoeForCausalLM)
TINY_MOE = dict(TINY, num_experts=8, top_k=2, moe_inter=128, shared_inter=256)


def tiny_state_dict_moe(seed: int = 0, dense_layers=()) -> dict:
    """Checkpoint-format names (one gate/up/down matrix per expert, as the published Qwen-MoE safetensors hold them). Layers in
    `dense_layers` (config.mlp_only_layers) keep the dense MLP of tiny_state_dict; the others get router + experts + shared expert."""
    sd = tiny_state_dict(seed)
    g = torch.Generator().manual_seed(seed + 2000)
    c = TINY_MOE
    H = c["hidden"]

    def rn(*shape, std=0.05):
        return torch.randn(*shape, generator=g) * std

    for li in range(c["layers"]):
        if li in dense_layers:
            continue
        p = f"model.layers.{li}.mlp."
        for k in ("gate_proj", "up_proj", "down_proj"):
            del sd[p + k + ".weight"]
        sd[p + "gate.weight"] = rn(c["num_experts"], H, std=0.5)  # wide router logits: top-k margins well above bf16 noise
        for e in range(c["num_experts"]):
            sd[p + f"experts.{e}.gate_proj.weight"] = rn(c["moe_inter"], H)
            sd[p + f"experts.{e}.up_proj.weight"] = rn(c["moe_inter"], H)
            sd[p + f"experts.{e}.down_proj.weight"] = rn(H, c["moe_inter"])
        sd[p + "shared_expert.gate_proj.weight"] = rn(c["shared_inter"], H)
        sd[p + "shared_expert.up_proj.weight"] = rn(c["shared_inter"], H)
        sd[p + "shared_expert.down_proj.weight"] = rn(H, c["shared_inter"])
        sd[p + "shared_expert_gate.weight"] = rn(1, H, std=0.3)
    return sd


def fuse_experts_for_transformers5(sd: dict, num_experts: int) -> dict:
    """transformers >= 5 keeps the experts as 3-D parameters (modeling_qwen2_moe.py:298-305): experts.gate_up_proj
    [E, 2 I, H] = cat(gate, up) per expert, experts.down_proj [E, H, I]."""
    out = {k: v for k, v in sd.items() if ".mlp.experts." not in k}
    layers = sorted({k.split(".mlp.experts.")[0] for k in sd if ".mlp.experts." in k})
    for p in layers:
        gu = [torch.cat([sd[f"{p}.mlp.experts.{e}.gate_proj.weight"], sd[f"{p}.mlp.experts.{e}.up_proj.weight"]], 0)
              for e in range(num_experts)]
        out[p + ".mlp.experts.gate_up_proj"] = torch.stack(gu)
        out[p + ".mlp.experts.down_proj"] = torch.stack([sd[f"{p}.mlp.experts.{e}.down_proj.weight"] for e in range(num_experts)])
    return out
