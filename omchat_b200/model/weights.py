"""Device-resident bf16 weights in the layouts the kernels consume, built from a state dict that uses the reference's own
parameter names (`model.vision_tower.vision_tower.*`, `model.mm_projector.{0,2}.*`, `model.layers.*`, `model.norm.weight`,
`model.embed_tokens.weight`, `lm_head.weight`; HF-hub layout names are remapped with the table of
convert_omchat_to_hf.py:26-35 read backwards), or random-initialised directly on the device for benchmarks.

Layout decisions (all K-major = nn.Linear's own [out, in] layout, so no transposes):
  - patch-embed conv weight [C,3,14,14] -> [C, 640] (K = 588 zero-padded to a TMA-legal row pitch)
  - Qwen2 q/k/v -> one [Hq*128 + 2*Hkv*128, hidden] matrix + one bias vector
  - Qwen2 gate/up -> one [2*I, hidden] matrix with rows alternating gate_i, up_i, so adjacent accumulator columns of
    the GEMM (or adjacent rows of a decode GEMV warp) are a matching gate/up pair for the fused SwiGLU epilogue, and any
    even row split is a valid tensor-parallel shard
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from ..config import OmChatQwen2Config

VT = "model.vision_tower.vision_tower."
PATCH_K = 640

# convert_omchat_to_hf.py:26-35 (omchat name fragment -> HF-hub name fragment)
KEYS_TO_MODIFY_MAPPING = {
    "model.vision_tower.vision_tower.": "vision_tower.",
    "model.mm_projector.0": "multi_modal_projector.linear_1",
    "model.mm_projector.2": "multi_modal_projector.linear_2",
    "model.": "language_model.model.",
    "lm_head.": "language_model.lm_head.",
}


def from_hf_names(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Map an HF-hub style state dict (OmChatForConditionalGeneration) back to the omchat names used here."""
    out = {}
    for k, v in sd.items():
        if k.startswith("vision_tower."):
            k = VT + k[len("vision_tower."):]
        elif k.startswith("multi_modal_projector.linear_1"):
            k = "model.mm_projector.0" + k[len("multi_modal_projector.linear_1"):]
        elif k.startswith("multi_modal_projector.linear_2"):
            k = "model.mm_projector.2" + k[len("multi_modal_projector.linear_2"):]
        elif k.startswith("language_model.model."):
            k = "model." + k[len("language_model.model."):]
        elif k.startswith("language_model.lm_head."):
            k = "lm_head." + k[len("language_model.lm_head."):]
        out[k] = v
    return out


def interleave_gate_up(gate: torch.Tensor, up: torch.Tensor) -> torch.Tensor:
    I, K = gate.shape
    return torch.stack([gate, up], dim=1).reshape(2 * I, K)


def pad_head_rows(w: torch.Tensor, groups: int, heads: int, D: int, Dp: int) -> torch.Tensor:
    """[groups * heads * D, ...] -> [groups * heads * Dp, ...], every head's D rows followed by Dp - D zero rows: a tower whose
    head_dim is below the attention kernels' 128 (InternViT-300M: 64) runs on zero-padded heads - q.k and P.V are unchanged by
    the zero dims, the scale stays D^-0.5."""
    rest = w.shape[1:]
    out = torch.zeros(groups, heads, Dp, *rest, device=w.device, dtype=w.dtype)
    out[:, :, :D] = w.reshape(groups, heads, D, *rest)
    return out.reshape(groups * heads * Dp, *rest).contiguous()


def pad_head_cols(w: torch.Tensor, heads: int, D: int, Dp: int) -> torch.Tensor:
    """proj weight [N, heads * D] -> [N, heads * Dp] with zero columns under the padded dims of the attention output."""
    N = w.shape[0]
    out = torch.zeros(N, heads, Dp, device=w.device, dtype=w.dtype)
    out[:, :, :D] = w.reshape(N, heads, D)
    return out.reshape(N, heads * Dp).contiguous()


def fold_norm(w: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    """W'[n, k] = W[n, k] * g[k]: the RMSNorm weight in front of a linear layer folded into its columns (fp32 product, one
    bf16 rounding). The GEMM then runs on the raw residual stream and scales rows by rstd in its epilogue."""
    return (w.float() * g.float()[None, :]).to(torch.bfloat16).contiguous()


@dataclass
class VitLayerW:
    norm1: torch.Tensor
    qkv: torch.Tensor
    q_norm: torch.Tensor
    k_norm: torch.Tensor
    proj_w: torch.Tensor
    proj_b: torch.Tensor
    ls1: torch.Tensor
    norm2: torch.Tensor
    fc1_w: torch.Tensor
    fc1_b: torch.Tensor
    fc2_w: torch.Tensor
    fc2_b: torch.Tensor
    ls2: torch.Tensor
    # InternViT-300M variant: LayerNorm biases (norm_type = 'layer_norm') and the optional qkv bias (config.qkv_bias)
    norm1_b: Optional[torch.Tensor] = None
    norm2_b: Optional[torch.Tensor] = None
    qkv_b: Optional[torch.Tensor] = None


@dataclass
class VitW:
    patch_w: torch.Tensor  # [C, 640]
    patch_b: torch.Tensor
    cls: torch.Tensor  # [C]
    pos: torch.Tensor  # [P+1, C]
    layers: List[VitLayerW]


@dataclass
class ProjW:
    w0: torch.Tensor
    b0: torch.Tensor
    w2: torch.Tensor
    b2: torch.Tensor


@dataclass
class MoeLayerW:
    """Sparse MLP block of a Qwen2-MoE layer (transformers modeling_qwen2_moe.py:295-374) in the grouped GEMM's layout."""
    router_w: torch.Tensor        # [E, C]           mlp.gate.weight
    shared_gate_w: torch.Tensor   # [C]              mlp.shared_expert_gate.weight
    experts_gate_up: torch.Tensor  # [E * 2 I, C]     per expert: gate / up rows interleaved (SwiGLU epilogue layout)
    experts_down: torch.Tensor    # [E * C, I]
    shared_gate_up: torch.Tensor  # [2 Is, C]        interleaved
    shared_down: torch.Tensor     # [C, Is]


@dataclass
class LlmLayerW:
    ln1: torch.Tensor
    qkv_w: torch.Tensor
    qkv_b: torch.Tensor
    o_w: torch.Tensor
    ln2: torch.Tensor
    gate_up_w: Optional[torch.Tensor]  # None in a sparse (mixture-of-experts) layer
    down_w: Optional[torch.Tensor]
    moe: Optional[MoeLayerW] = None


@dataclass
class LlmW:
    embed: torch.Tensor
    layers: List[LlmLayerW]
    norm: torch.Tensor
    lm_head: torch.Tensor  # [V_local, C] (vocab-parallel under TP)
    q_heads_local: int = 28
    kv_heads_local: int = 4


@dataclass
class TPPlan:
    """Which slices of the Qwen2 weights one tensor-parallel rank owns.
    q_heads: global q-head indices (-1 = zero pad head), kv_heads: global kv-head indices, i_lo/i_hi: MLP rows,
    i_pad: zero rows appended so the local intermediate size is a multiple of 8 (GEMM N/K granularity),
    v_lo/v_hi: vocab rows of lm_head."""
    q_heads: List[int]
    kv_heads: List[int]
    i_lo: int
    i_hi: int
    i_pad: int
    v_lo: int
    v_hi: int


def tp_plan(cfg: OmChatQwen2Config, rank: int, size: int) -> TPPlan:
    Hq, Hkv, I, V = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.intermediate_size, cfg.vocab_size
    G = Hq // Hkv
    if size <= Hkv:
        if Hkv % size:
            raise ValueError(f"tp_size {size} must divide the {Hkv} kv heads (or be a multiple of it)")
        per = Hkv // size
        kv = list(range(rank * per, (rank + 1) * per))
        q = [g * G + j for g in kv for j in range(G)]
    else:
        if size % Hkv:
            raise ValueError(f"tp_size {size} must be a multiple of the {Hkv} kv heads")
        rep = size // Hkv  # ranks sharing one (replicated) kv head; its G q-heads are dealt to them, padded
        g, part = rank // rep, rank % rep
        chunk = (G + rep - 1) // rep
        q = [g * G + j if j < G else -1 for j in range(part * chunk, (part + 1) * chunk)]
        kv = [g]
    if I % size or V % size:
        raise ValueError("tp_size must divide intermediate_size and vocab_size")
    il = I // size
    return TPPlan(q_heads=q, kv_heads=kv, i_lo=rank * il, i_hi=(rank + 1) * il, i_pad=(-il) % 8,
                  v_lo=rank * (V // size), v_hi=(rank + 1) * (V // size))


def _take_heads(t: torch.Tensor, heads: List[int], D: int, dim: int) -> torch.Tensor:
    """Select head blocks (D rows/cols each) along `dim`; -1 inserts a zero block."""
    parts = []
    for hd in heads:
        if hd < 0:
            shp = list(t.shape)
            shp[dim] = D
            parts.append(torch.zeros(shp, dtype=t.dtype, device=t.device))
        else:
            parts.append(t.narrow(dim, hd * D, D))
    return torch.cat(parts, dim=dim)


def shard_llm_layer(q_w, q_b, k_w, k_b, v_w, v_b, o_w, gate, up, down, plan: TPPlan, D: int):
    """Column-parallel q/k/v + gate/up, row-parallel o/down for one rank; returns the fused kernel layouts."""
    qkv_w = torch.cat([_take_heads(q_w, plan.q_heads, D, 0), _take_heads(k_w, plan.kv_heads, D, 0),
                       _take_heads(v_w, plan.kv_heads, D, 0)], dim=0)
    qkv_b = torch.cat([_take_heads(q_b, plan.q_heads, D, 0), _take_heads(k_b, plan.kv_heads, D, 0),
                       _take_heads(v_b, plan.kv_heads, D, 0)], dim=0)
    o = _take_heads(o_w, plan.q_heads, D, 1)
    g, u, d = gate[plan.i_lo:plan.i_hi], up[plan.i_lo:plan.i_hi], down[:, plan.i_lo:plan.i_hi]
    if plan.i_pad:
        z = torch.zeros(plan.i_pad, g.shape[1], dtype=g.dtype, device=g.device)
        g, u = torch.cat([g, z]), torch.cat([u, z])
        d = torch.cat([d, torch.zeros(d.shape[0], plan.i_pad, dtype=d.dtype, device=d.device)], dim=1)
    return qkv_w, qkv_b, o.contiguous(), interleave_gate_up(g, u), d.contiguous()


@dataclass
class OmChatWeights:
    vit: Optional[VitW]
    proj: Optional[ProjW]
    llm: LlmW


def _dev(t: torch.Tensor, device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=torch.bfloat16).contiguous()


def from_state_dict(sd: Dict[str, torch.Tensor], cfg: OmChatQwen2Config, device="cuda", tp_rank: int = 0,
                    tp_size: int = 1) -> OmChatWeights:
    if any(k.startswith("language_model.") for k in sd):
        sd = from_hf_names(sd)
    vc = cfg.vision_config
    vit = proj = None
    if (VT + "embeddings.class_embedding") in sd:
        C = vc.hidden_size
        pw = sd[VT + "embeddings.patch_embedding.weight"].reshape(C, -1)
        patch_w = torch.zeros(C, PATCH_K, dtype=pw.dtype)
        patch_w[:, : pw.shape[1]] = pw
        layers = []
        for li in range(vc.num_hidden_layers):
            p = f"{VT}encoder.layers.{li}."
            opt = lambda k: _dev(sd[k], device) if k in sd else None  # noqa: E731
            if vc.norm_type == "layer_norm" and (p + "norm1.bias") not in sd:
                raise KeyError(f"norm_type 'layer_norm' needs {p}norm1.bias")
            layers.append(VitLayerW(
                norm1=_dev(sd[p + "norm1.weight"], device), qkv=_dev(sd[p + "attn.qkv.weight"], device),
                q_norm=opt(p + "attn.q_norm.weight"), k_norm=opt(p + "attn.k_norm.weight"),
                norm1_b=opt(p + "norm1.bias"), norm2_b=opt(p + "norm2.bias"), qkv_b=opt(p + "attn.qkv.bias"),
                proj_w=_dev(sd[p + "attn.proj.weight"], device), proj_b=_dev(sd[p + "attn.proj.bias"], device),
                ls1=_dev(sd[p + "ls1"], device), norm2=_dev(sd[p + "norm2.weight"], device),
                fc1_w=_dev(sd[p + "mlp.fc1.weight"], device), fc1_b=_dev(sd[p + "mlp.fc1.bias"], device),
                fc2_w=_dev(sd[p + "mlp.fc2.weight"], device), fc2_b=_dev(sd[p + "mlp.fc2.bias"], device),
                ls2=_dev(sd[p + "ls2"], device)))
        vit = VitW(patch_w=_dev(patch_w, device), patch_b=_dev(sd[VT + "embeddings.patch_embedding.bias"], device),
                   cls=_dev(sd[VT + "embeddings.class_embedding"].reshape(-1), device),
                   pos=_dev(sd[VT + "embeddings.position_embedding"].reshape(-1, C), device), layers=layers)
        proj = ProjW(w0=_dev(sd["model.mm_projector.0.weight"], device), b0=_dev(sd["model.mm_projector.0.bias"], device),
                     w2=_dev(sd["model.mm_projector.2.weight"], device), b2=_dev(sd["model.mm_projector.2.bias"], device))
    plan = tp_plan(cfg, tp_rank, tp_size)
    D = cfg.head_dim
    layers = []
    for li in range(cfg.num_hidden_layers):
        p = f"model.layers.{li}."
        a = p + "self_attn."
        if getattr(cfg, "num_experts", 0) > 0 and cfg.layer_is_sparse(li):
            layers.append(_moe_layer_from_sd(sd, p, cfg, tp_size, device))
            continue
        qkv_w, qkv_b, o_w, gu, down = shard_llm_layer(
            sd[a + "q_proj.weight"], sd[a + "q_proj.bias"], sd[a + "k_proj.weight"], sd[a + "k_proj.bias"],
            sd[a + "v_proj.weight"], sd[a + "v_proj.bias"], sd[a + "o_proj.weight"], sd[p + "mlp.gate_proj.weight"],
            sd[p + "mlp.up_proj.weight"], sd[p + "mlp.down_proj.weight"], plan, D)
        layers.append(LlmLayerW(
            ln1=_dev(sd[p + "input_layernorm.weight"], device), qkv_w=_dev(qkv_w, device), qkv_b=_dev(qkv_b, device),
            o_w=_dev(o_w, device), ln2=_dev(sd[p + "post_attention_layernorm.weight"], device),
            gate_up_w=_dev(gu, device), down_w=_dev(down, device)))
    llm = LlmW(embed=_dev(sd["model.embed_tokens.weight"], device), layers=layers,
               norm=_dev(sd["model.norm.weight"], device), lm_head=_dev(sd["lm_head.weight"][plan.v_lo:plan.v_hi], device),
               q_heads_local=len(plan.q_heads), kv_heads_local=len(plan.kv_heads))
    return OmChatWeights(vit=vit, proj=proj, llm=llm)


def moe_experts(sd: Dict[str, torch.Tensor], p: str, E: int):
    """Per-expert (gate, up, down) matrices of layer prefix `p` from either checkpoint layout: one matrix per expert
    (`mlp.experts.{e}.gate_proj.weight` ..., the published safetensors) or transformers >= 5's fused 3-D parameters
    (`mlp.experts.gate_up_proj` [E, 2 I, C] = cat(gate, up), `mlp.experts.down_proj` [E, C, I]; modeling_qwen2_moe.py:298-305)."""
    if (p + "mlp.experts.gate_up_proj") in sd:
        gu, dn = sd[p + "mlp.experts.gate_up_proj"], sd[p + "mlp.experts.down_proj"]
        I = gu.shape[1] // 2
        return [(gu[e, :I], gu[e, I:], dn[e]) for e in range(E)]
    return [(sd[f"{p}mlp.experts.{e}.gate_proj.weight"], sd[f"{p}mlp.experts.{e}.up_proj.weight"],
             sd[f"{p}mlp.experts.{e}.down_proj.weight"]) for e in range(E)]


def _moe_layer_from_sd(sd, p: str, cfg, tp_size: int, device) -> "LlmLayerW":
    if tp_size != 1:
        raise NotImplementedError("tensor parallelism is not built for the Qwen2-MoE variant")
    a = p + "self_attn."
    qkv_w = torch.cat([sd[a + "q_proj.weight"], sd[a + "k_proj.weight"], sd[a + "v_proj.weight"]], 0)
    qkv_b = torch.cat([sd[a + "q_proj.bias"], sd[a + "k_proj.bias"], sd[a + "v_proj.bias"]], 0)
    ex = moe_experts(sd, p, cfg.num_experts)
    moe = MoeLayerW(
        router_w=_dev(sd[p + "mlp.gate.weight"], device), shared_gate_w=_dev(sd[p + "mlp.shared_expert_gate.weight"].reshape(-1), device),
        experts_gate_up=_dev(torch.cat([interleave_gate_up(g, u) for g, u, _ in ex], 0), device),
        experts_down=_dev(torch.cat([d for _, _, d in ex], 0), device),
        shared_gate_up=_dev(interleave_gate_up(sd[p + "mlp.shared_expert.gate_proj.weight"], sd[p + "mlp.shared_expert.up_proj.weight"]), device),
        shared_down=_dev(sd[p + "mlp.shared_expert.down_proj.weight"], device))
    return LlmLayerW(ln1=_dev(sd[p + "input_layernorm.weight"], device), qkv_w=_dev(qkv_w, device), qkv_b=_dev(qkv_b, device),
                     o_w=_dev(sd[a + "o_proj.weight"], device), ln2=_dev(sd[p + "post_attention_layernorm.weight"], device),
                     gate_up_w=None, down_w=None, moe=moe)


def random_init(cfg: OmChatQwen2Config, device="cuda", seed: int = 0, vision: bool = True, text: bool = True,
                tp_rank: int = 0, tp_size: int = 1) -> OmChatWeights:
    """Random weights of the configured architecture created directly on the device (no host copy of the 13B model).
    Scales follow HF `_init_weights` (normal std 0.02 for Linear/Embedding; ViT cls/pos randn; layer-scale 0.1; norm
    weights 1) with small perturbations so every parameter matters. Layout identical to from_state_dict(). Under TP
    every rank draws the same full matrices from the same seed and keeps its shard."""
    g = torch.Generator(device=device).manual_seed(seed)
    vc = cfg.vision_config

    def rn(*shape, std=0.02, mean=0.0):
        return (torch.randn(*shape, generator=g, device=device, dtype=torch.float32) * std + mean).to(torch.bfloat16)

    vit = proj = None
    if vision:
        C, I = vc.hidden_size, vc.intermediate_size
        patch_w = torch.zeros(C, PATCH_K, device=device, dtype=torch.bfloat16)
        patch_w[:, :588] = rn(C, 588)
        ln = vc.norm_type == "layer_norm"
        layers = [VitLayerW(norm1=rn(C, std=0.02, mean=1.0), qkv=rn(3 * C, C),
                            q_norm=rn(C, std=0.02, mean=1.0) if vc.qk_normalization else None,
                            k_norm=rn(C, std=0.02, mean=1.0) if vc.qk_normalization else None,
                            proj_w=rn(C, C), proj_b=rn(C), ls1=rn(C, std=0.01, mean=0.1),
                            norm2=rn(C, std=0.02, mean=1.0), fc1_w=rn(I, C), fc1_b=rn(I), fc2_w=rn(C, I), fc2_b=rn(C),
                            ls2=rn(C, std=0.01, mean=0.1), norm1_b=rn(C) if ln else None, norm2_b=rn(C) if ln else None,
                            qkv_b=rn(3 * C) if vc.qkv_bias else None) for _ in range(vc.num_hidden_layers)]
        vit = VitW(patch_w=patch_w, patch_b=rn(C), cls=rn(C, std=1.0), pos=rn(vc.num_patches + 1, C, std=1.0), layers=layers)
        H = cfg.hidden_size
        Cin = C * cfg.pixel_shuffle_down ** 2
        proj = ProjW(w0=rn(H, Cin), b0=rn(H), w2=rn(H, H), b2=rn(H))
    H, I, D = cfg.hidden_size, cfg.intermediate_size, cfg.head_dim
    nq, nkv = cfg.num_attention_heads, cfg.num_key_value_heads
    plan = tp_plan(cfg, tp_rank, tp_size)
    layers = []
    if text:
        for li in range(cfg.num_hidden_layers):
            ln1, ln2 = rn(H, std=0.02, mean=1.0), rn(H, std=0.02, mean=1.0)
            if getattr(cfg, "num_experts", 0) > 0 and cfg.layer_is_sparse(li):
                if tp_size != 1:
                    raise NotImplementedError("tensor parallelism is not built for the Qwen2-MoE variant")
                E, Im, Is = cfg.num_experts, cfg.moe_intermediate_size, cfg.shared_expert_intermediate_size
                moe = MoeLayerW(router_w=rn(E, H, std=0.3), shared_gate_w=rn(H, std=0.1),
                                experts_gate_up=rn(E * 2 * Im, H), experts_down=rn(E * H, Im),
                                shared_gate_up=rn(2 * Is, H), shared_down=rn(H, Is))
                layers.append(LlmLayerW(ln1=ln1, qkv_w=rn((nq + 2 * nkv) * D, H), qkv_b=rn((nq + 2 * nkv) * D), o_w=rn(H, nq * D),
                                        ln2=ln2, gate_up_w=None, down_w=None, moe=moe))
                continue
            qkv_w, qkv_b, o_w, gu, down = shard_llm_layer(
                rn(nq * D, H), rn(nq * D), rn(nkv * D, H), rn(nkv * D), rn(nkv * D, H), rn(nkv * D), rn(H, nq * D),
                rn(I, H), rn(I, H), rn(H, I), plan, D)
            layers.append(LlmLayerW(ln1=ln1, qkv_w=qkv_w.contiguous(), qkv_b=qkv_b.contiguous(), o_w=o_w, ln2=ln2,
                                    gate_up_w=gu.contiguous(), down_w=down))
    llm = LlmW(embed=rn(cfg.vocab_size, H) if text else torch.zeros(1, H, device=device, dtype=torch.bfloat16),
               layers=layers, norm=rn(H, std=0.02, mean=1.0),
               lm_head=rn(cfg.vocab_size, H)[plan.v_lo:plan.v_hi].contiguous() if text
               else torch.zeros(8, H, device=device, dtype=torch.bfloat16),
               q_heads_local=len(plan.q_heads), kv_heads_local=len(plan.kv_heads))
    return OmChatWeights(vit=vit, proj=proj, llm=llm)


def to_reference_state_dict(w: OmChatWeights, cfg: OmChatQwen2Config) -> Dict[str, torch.Tensor]:
    """Inverse of from_state_dict (bf16 tensors on their current device) — lets the parity tests hand the SAME weights
    to the checker under the reference's parameter names."""
    sd: Dict[str, torch.Tensor] = {}
    vc = cfg.vision_config
    if w.vit is not None:
        C = vc.hidden_size
        sd[VT + "embeddings.class_embedding"] = w.vit.cls.view(1, 1, C)
        sd[VT + "embeddings.position_embedding"] = w.vit.pos.view(1, -1, C)
        sd[VT + "embeddings.patch_embedding.weight"] = w.vit.patch_w[:, :588].reshape(C, 3, 14, 14)
        sd[VT + "embeddings.patch_embedding.bias"] = w.vit.patch_b
        for li, l in enumerate(w.vit.layers):
            p = f"{VT}encoder.layers.{li}."
            for k, v in (("attn.q_norm.weight", l.q_norm), ("attn.k_norm.weight", l.k_norm), ("norm1.bias", l.norm1_b),
                         ("norm2.bias", l.norm2_b), ("attn.qkv.bias", l.qkv_b)):
                if v is not None:
                    sd[p + k] = v
            sd.update({p + "norm1.weight": l.norm1, p + "attn.qkv.weight": l.qkv,
                       p + "attn.proj.weight": l.proj_w, p + "attn.proj.bias": l.proj_b,
                       p + "ls1": l.ls1, p + "norm2.weight": l.norm2, p + "mlp.fc1.weight": l.fc1_w,
                       p + "mlp.fc1.bias": l.fc1_b, p + "mlp.fc2.weight": l.fc2_w, p + "mlp.fc2.bias": l.fc2_b, p + "ls2": l.ls2})
        sd.update({"model.mm_projector.0.weight": w.proj.w0, "model.mm_projector.0.bias": w.proj.b0,
                   "model.mm_projector.2.weight": w.proj.w2, "model.mm_projector.2.bias": w.proj.b2})
    D = cfg.head_dim
    nq, nkv = cfg.num_attention_heads * D, cfg.num_key_value_heads * D
    for li, l in enumerate(w.llm.layers):
        p = f"model.layers.{li}."
        if l.moe is not None:
            m, E = l.moe, l.moe.router_w.shape[0]
            K = m.router_w.shape[1]
            egu = m.experts_gate_up.view(E, -1, 2, K)
            edn = m.experts_down.view(E, K, -1)
            for e in range(E):
                sd.update({f"{p}mlp.experts.{e}.gate_proj.weight": egu[e, :, 0], f"{p}mlp.experts.{e}.up_proj.weight": egu[e, :, 1],
                           f"{p}mlp.experts.{e}.down_proj.weight": edn[e]})
            sgu = m.shared_gate_up.view(-1, 2, K)
            sd.update({p + "mlp.gate.weight": m.router_w, p + "mlp.shared_expert_gate.weight": m.shared_gate_w.view(1, K),
                       p + "mlp.shared_expert.gate_proj.weight": sgu[:, 0], p + "mlp.shared_expert.up_proj.weight": sgu[:, 1],
                       p + "mlp.shared_expert.down_proj.weight": m.shared_down})
            sd.update({p + "input_layernorm.weight": l.ln1, p + "post_attention_layernorm.weight": l.ln2,
                       p + "self_attn.q_proj.weight": l.qkv_w[:nq], p + "self_attn.k_proj.weight": l.qkv_w[nq:nq + nkv],
                       p + "self_attn.v_proj.weight": l.qkv_w[nq + nkv:], p + "self_attn.q_proj.bias": l.qkv_b[:nq],
                       p + "self_attn.k_proj.bias": l.qkv_b[nq:nq + nkv], p + "self_attn.v_proj.bias": l.qkv_b[nq + nkv:],
                       p + "self_attn.o_proj.weight": l.o_w})
            continue
        I2, K = l.gate_up_w.shape
        gu = l.gate_up_w.view(I2 // 2, 2, K)
        sd.update({p + "input_layernorm.weight": l.ln1, p + "post_attention_layernorm.weight": l.ln2,
                   p + "self_attn.q_proj.weight": l.qkv_w[:nq], p + "self_attn.k_proj.weight": l.qkv_w[nq:nq + nkv],
                   p + "self_attn.v_proj.weight": l.qkv_w[nq + nkv:], p + "self_attn.q_proj.bias": l.qkv_b[:nq],
                   p + "self_attn.k_proj.bias": l.qkv_b[nq:nq + nkv], p + "self_attn.v_proj.bias": l.qkv_b[nq + nkv:],
                   p + "self_attn.o_proj.weight": l.o_w, p + "mlp.gate_proj.weight": gu[:, 0].reshape(I2 // 2, K),
                   p + "mlp.up_proj.weight": gu[:, 1].reshape(I2 // 2, K), p + "mlp.down_proj.weight": l.down_w})
    sd["model.embed_tokens.weight"] = w.llm.embed
    sd["model.norm.weight"] = w.llm.norm
    sd["lm_head.weight"] = w.llm.lm_head
    return sd
