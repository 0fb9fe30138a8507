"""CPU oracle for the OmChat multimodal forward pass — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain restatement (torch CPU tensors, fp32 by default) of the reference's algorithm for the hot path: InternViT-6B
tower -> feature select (-> pixel shuffle) -> mm_projector -> image-token splice -> Qwen2 decoder prefill + greedy
decode. Every function cites the reference file:line it follows (paths relative to the reference root; the Qwen2
decoder lives in the third-party dependency `transformers` (reference pin ==4.41.2, pyproject.toml:22; the copy this
oracle was checked against is transformers 5.5.0, models/qwen2/modeling_qwen2.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module, and
only as the checker / the CPU baseline; the product path (omchat_b200/) never does.

Pinning: the reference ships no tests, golden vectors or fixtures for this path, so the oracle is
pinned against outputs of the reference itself: tests/golden/make_golden.py imports the real reference modules
(with timm/peft/accelerate shims) at a tiny configuration, runs them on seeded inputs and stores inputs, weights and
outputs in tests/golden/*.pt; tests/test_oracle.py checks this restatement against those files. The pixel-shuffle
(ratio 0.5) has NO reference symbol ("parity unpinned" for that one function): it is pinned against the InternVL
view/permute formulation quoted in SURVEY.md §8 a7.
The two lighter model families are pinned the same way: the InternViT-300M branch (vit_norm / LayerNorm, no QK-norm, qkv bias)
on tests/golden/golden_tiny_300m.pt = outputs of the reference's InternVIT300mVisionTower (make_golden_300m.py,
tests/test_oracle_300m.py); the Qwen2-MoE branch (moe_route / qwen2_moe_block; third-party transformers
models/qwen2_moe/modeling_qwen2_moe.py, checked copy 5.5.0) on golden_tiny_moe.pt = outputs of the reference's
OmChatQwen2MoeForCausalLM in two configurations (make_golden_moe.py, tests/test_oracle_moe.py). ROUTE_TRACE exposes the
routing margins to the GPU tests: a bf16 run may legitimately send a token on a router near-tie to the other expert.

The tower / projector / decoder functions are device- and dtype-agnostic plain torch: the parity tests also run them with
bf16 CUDA tensors ("what the reference's own 16-bit PyTorch path gives on this box") to calibrate how much of an error
against the fp32 gold is bf16 itself.

Weights are passed as a flat dict keyed by the reference's own state-dict names
(`model.vision_tower.vision_tower.*`, `model.mm_projector.{0,2}.*`, `model.layers.*`, `model.norm.weight`,
`model.embed_tokens.weight`, `lm_head.weight`).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

IMAGE_TOKEN_INDEX = -200  # omchat/constants.py:8
VT = "model.vision_tower.vision_tower."


@dataclass
class OracleConfig:
    # vision (intern_vit_6b/configuration_intern_vit.py:63-83)
    vit_hidden: int = 3200
    vit_heads: int = 25
    vit_inter: int = 12800
    vit_layers: int = 45
    vit_eps: float = 1e-6
    # InternViT-300M variant (intern_vit_300m/configuration_intern_vit.py:60-80): 'layer_norm', no QK-norm, optional qkv bias
    vit_norm_type: str = "rms_norm"
    vit_qk_norm: bool = True
    image_size: int = 448
    patch_size: int = 14
    select_layer: int = -1
    select_feature: str = "patch"
    pixel_shuffle_down: int = 1  # 1 = reference behaviour (no shuffle); 2 = ratio 0.5
    # text (Qwen2-7B)
    hidden: int = 3584
    heads: int = 28
    kv_heads: int = 4
    inter: int = 18944
    layers: int = 28
    vocab: int = 152064
    rms_eps: float = 1e-6
    rope_theta: float = 1e6
    # Qwen2-MoE language model (omchat_qwen2_moe.py:14-17 = transformers Qwen2MoeConfig); num_experts 0 = the dense Qwen2
    num_experts: int = 0
    top_k: int = 4
    norm_topk_prob: bool = False
    decoder_sparse_step: int = 1
    mlp_only_layers: Tuple[int, ...] = ()
    max_len: Optional[int] = None  # tokenizer_model_max_length
    padding_side: str = "right"

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads


# ----------------------------------------------------------------------------------------------- vision tower
def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """InternRMSNorm.forward intern_vit_6b/modeling_intern_vit.py:39-44 (== Qwen2RMSNorm modeling_qwen2.py:258-263):
    statistics in fp32, cast back to the input dtype, THEN multiply by the weight."""
    dt = x.dtype
    h = x.to(torch.float32)
    var = h.pow(2).mean(-1, keepdim=True)
    h = h * torch.rsqrt(var + eps)
    return w * h.to(dt)


def vit_norm(x: torch.Tensor, sd: Dict[str, torch.Tensor], name: str, cfg: OracleConfig) -> torch.Tensor:
    """norm1 / norm2 of an encoder layer: NORM2FN[config.norm_type] (intern_vit_300m/modeling_intern_vit.py:61-64,209-210) -
    InternRMSNorm, or torch.nn.LayerNorm (weight + bias, biased variance) for the 300M tower."""
    if cfg.vit_norm_type == "layer_norm":
        return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], cfg.vit_eps)
    return rms_norm(x, sd[name + ".weight"], cfg.vit_eps)


def vit_embeddings(pixels: torch.Tensor, sd: Dict[str, torch.Tensor], cfg: OracleConfig) -> torch.Tensor:
    """InternVisionEmbeddings.forward modeling_intern_vit.py:90-102. The bicubic resize of the position grid (:82-88)
    is applied like the reference does (identity at the native 32x32 grid)."""
    w = sd[VT + "embeddings.patch_embedding.weight"]
    b = sd[VT + "embeddings.patch_embedding.bias"]
    pe = F.conv2d(pixels.to(w.dtype), w, b, stride=cfg.patch_size)  # [B, C, gh, gw]
    B, C, gh, gw = pe.shape
    pe = pe.flatten(2).transpose(1, 2)
    cls = sd[VT + "embeddings.class_embedding"].expand(B, 1, -1).to(w.dtype)
    emb = torch.cat([cls, pe], dim=1)
    pos = sd[VT + "embeddings.position_embedding"]
    g = cfg.image_size // cfg.patch_size
    grid = pos[:, 1:, :].float().reshape(1, g, g, -1).permute(0, 3, 1, 2)
    grid = F.interpolate(grid, size=(gh, gw), mode="bicubic", align_corners=False)
    grid = grid.reshape(1, -1, gh * gw).permute(0, 2, 1).to(pos.dtype)
    pos = torch.cat([pos[:, :1, :], grid], dim=1)
    return emb + pos.to(w.dtype)


def vit_attention(x: torch.Tensor, sd: Dict[str, torch.Tensor], pre: str, cfg: OracleConfig) -> torch.Tensor:
    """InternAttention._naive_attn modeling_intern_vit.py:138-155: qkv (no bias), q/k RMS-normed over ALL heads
    flattened (:143-146), non-causal softmax attention, proj + bias."""
    B, N, C = x.shape
    H = cfg.vit_heads
    qkv = F.linear(x, sd[pre + "attn.qkv.weight"], sd.get(pre + "attn.qkv.bias"))  # bias iff config.qkv_bias (:131)
    qkv = qkv.reshape(B, N, 3, H, C // H).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.unbind(0)
    if cfg.vit_qk_norm:  # config.qk_normalization (:136-140,143-146): off in the 300M tower
        q = rms_norm(q.transpose(1, 2).flatten(-2, -1), sd[pre + "attn.q_norm.weight"], cfg.vit_eps).view(B, N, H, C // H).transpose(1, 2)
        k = rms_norm(k.transpose(1, 2).flatten(-2, -1), sd[pre + "attn.k_norm.weight"], cfg.vit_eps).view(B, N, H, C // H).transpose(1, 2)
    scale = (C // H) ** -0.5
    attn = (q * scale) @ k.transpose(-2, -1)
    attn = attn.softmax(dim=-1)
    out = (attn @ v).transpose(1, 2).reshape(B, N, C)
    return F.linear(out, sd[pre + "attn.proj.weight"], sd[pre + "attn.proj.bias"])


def vit_mlp(x: torch.Tensor, sd: Dict[str, torch.Tensor], pre: str) -> torch.Tensor:
    """InternMLP.forward modeling_intern_vit.py:187-191, hidden_act='gelu' = exact erf GELU."""
    h = F.linear(x, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"])
    h = F.gelu(h)
    return F.linear(h, sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"])


def vit_layer(x: torch.Tensor, sd: Dict[str, torch.Tensor], li: int, cfg: OracleConfig) -> torch.Tensor:
    """InternVisionEncoderLayer.forward modeling_intern_vit.py:218-220 (drop_path is Identity at rate 0, :207-208)."""
    pre = f"{VT}encoder.layers.{li}."
    x = x + vit_attention(vit_norm(x, sd, pre + "norm1", cfg), sd, pre, cfg) * sd[pre + "ls1"]
    x = x + vit_mlp(vit_norm(x, sd, pre + "norm2", cfg), sd, pre) * sd[pre + "ls2"]
    return x


def vit_tower(pixels: torch.Tensor, sd: Dict[str, torch.Tensor], cfg: OracleConfig,
              return_all: bool = False):
    """InternVisionEncoder.forward modeling_intern_vit.py:268-279 + InternVITVisionTower.feature_select
    internVIT_encoder.py:35-43: hidden_states = [embeddings, layer1, ..., layerL]; pick select_layer; drop CLS."""
    h = vit_embeddings(pixels, sd, cfg)
    states = [h]
    for li in range(cfg.vit_layers):
        h = vit_layer(h, sd, li, cfg)
        states.append(h)
    feats = states[cfg.select_layer]
    if cfg.select_feature == "patch":
        feats = feats[:, 1:]
    elif cfg.select_feature != "cls_patch":
        raise ValueError(f"Unexpected select feature: {cfg.select_feature}")
    return (feats, states) if return_all else feats


def pixel_shuffle(feats: torch.Tensor, down: int) -> torch.Tensor:
    """North-star addition, absent from the reference: InternVL pixel_shuffle(scale=1/down), v2
    ordering, restated with the original view/permute chain. feats [B, G*G, C] -> [B, (G/down)^2, C*down^2]."""
    if down == 1:
        return feats
    B, L, C = feats.shape
    G = int(math.isqrt(L))
    r = 1.0 / down
    x = feats.reshape(B, G, G, C)
    x = x.view(B, G, int(G * r), int(C / r))
    x = x.permute(0, 2, 1, 3).contiguous()
    x = x.view(B, int(G * r), int(G * r), int(C / (r * r)))
    x = x.permute(0, 2, 1, 3).contiguous()
    return x.reshape(B, -1, x.shape[-1])


def projector(feats: torch.Tensor, sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """mlp2x_gelu = Linear, GELU, Linear — multimodal_projector/builder.py:54-61."""
    h = F.linear(feats, sd["model.mm_projector.0.weight"], sd["model.mm_projector.0.bias"])
    h = F.gelu(h)
    return F.linear(h, sd["model.mm_projector.2.weight"], sd["model.mm_projector.2.bias"])


def encode_images(pixels: torch.Tensor, sd: Dict[str, torch.Tensor], cfg: OracleConfig) -> torch.Tensor:
    """OmChatMetaForCausalLM.encode_images omchat_arch.py:50-53 (+ optional pixel shuffle between tower and projector)."""
    return projector(pixel_shuffle(vit_tower(pixels, sd, cfg), cfg.pixel_shuffle_down), sd)


# ----------------------------------------------------------------------------------------------- splice
def splice_plan(input_ids: Sequence[Sequence[int]], n_img: int, L: int, max_len: Optional[int] = None):
    """Integer placement of prepare_inputs_labels_for_multimodal omchat_arch.py:115-164, pure Python (small cases).
    Returns per sequence a list of (kind, index, row): kind 0 = text token id `index`, kind 1 = row `row` of image
    block `index`. A sequence with no placeholder still consumes one image block (:122-129)."""
    plans: List[List[Tuple[int, int, int]]] = []
    cur = 0
    for ids in input_ids:
        plan: List[Tuple[int, int, int]] = []
        if sum(1 for t in ids if t == IMAGE_TOKEN_INDEX) == 0:
            if cur >= n_img:  # the reference indexes image_features[cur_image_idx] here (:123)
                raise IndexError("image-less sequence still consumes an image block")
            plan = [(0, int(t), 0) for t in ids]
            cur += 1
        else:
            for t in ids:
                if t == IMAGE_TOKEN_INDEX:
                    if cur >= n_img:
                        raise IndexError("more image placeholders than images")
                    plan.extend((1, cur, r) for r in range(L))
                    cur += 1
                else:
                    plan.append((0, int(t), 0))
        if max_len is not None:
            plan = plan[:max_len]
        plans.append(plan)
    return plans


def splice(input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor], image_feats: torch.Tensor,
           embed_table: torch.Tensor, cfg: OracleConfig):
    """prepare_inputs_labels_for_multimodal omchat_arch.py:55-209 (inference subset: no labels): strip padding by
    mask (:115), split at -200 and embed text (:131-139), interleave image features (:145-155), truncate (:161-164),
    pad to the batch max and rebuild mask / position ids (:166-195). Returns (embeds [b,T,C], mask [b,T] bool,
    position_ids [b,T] long, lengths)."""
    b = input_ids.shape[0]
    if attention_mask is None:
        attention_mask = torch.ones_like(input_ids, dtype=torch.bool)
    rows = [input_ids[i][attention_mask[i].bool()].tolist() for i in range(b)]
    plans = splice_plan(rows, image_feats.shape[0], image_feats.shape[1], cfg.max_len)
    seqs = []
    for plan in plans:
        parts = [embed_table[idx] if kind == 0 else image_feats[idx, row] for kind, idx, row in plan]
        seqs.append(torch.stack(parts) if parts else embed_table[:0])
    T = max(s.shape[0] for s in seqs)
    C = embed_table.shape[1]
    embeds = torch.zeros(b, T, C, dtype=embed_table.dtype)
    mask = torch.zeros(b, T, dtype=torch.bool)
    pos = torch.zeros(b, T, dtype=torch.long)
    for i, s in enumerate(seqs):
        n = s.shape[0]
        if n == 0:
            continue
        if cfg.padding_side == "left":
            embeds[i, -n:] = s; mask[i, -n:] = True; pos[i, -n:] = torch.arange(n)
        else:
            embeds[i, :n] = s; mask[i, :n] = True; pos[i, :n] = torch.arange(n)
    return embeds, mask, pos, [s.shape[0] for s in seqs]


# ----------------------------------------------------------------------------------------------- Qwen2 decoder
def rope_inv_freq(cfg: OracleConfig) -> torch.Tensor:
    """Qwen2RotaryEmbedding default init: inv_freq = 1 / theta^(arange(0,dim,2)/dim) (modeling_qwen2.py:51-100)."""
    d = cfg.head_dim
    return 1.0 / (cfg.rope_theta ** (torch.arange(0, d, 2, dtype=torch.int64).to(torch.float32) / d))


def rope_cos_sin(position_ids: torch.Tensor, cfg: OracleConfig, dtype: torch.dtype):
    """Qwen2RotaryEmbedding.forward modeling_qwen2.py:102-113: fp32 angles, emb = cat(freqs, freqs), cast to dtype."""
    inv = rope_inv_freq(cfg).to(position_ids.device)
    freqs = position_ids.to(torch.float32)[..., None] * inv  # [b, T, d/2]
    emb = torch.cat([freqs, freqs], dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    """modeling_qwen2.py:117-121."""
    x1 = x[..., : x.shape[-1] // 2]
    x2 = x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def apply_rope(q: torch.Tensor, k: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor):
    """apply_rotary_pos_emb modeling_qwen2.py:124-146 (unsqueeze over the head dim)."""
    cos, sin = cos.unsqueeze(1), sin.unsqueeze(1)
    return q * cos + rotate_half(q) * sin, k * cos + rotate_half(k) * sin


def qwen2_attention(x: torch.Tensor, sd, pre: str, cfg: OracleConfig, cos, sin, past: Optional[Tuple[torch.Tensor, torch.Tensor]],
                    key_mask: Optional[torch.Tensor]):
    """Qwen2Attention.forward modeling_qwen2.py:206-246 with eager_attention_forward :161-184: q/k/v with bias, RoPE,
    cache append, repeat_kv (:149-158), causal mask (+ key padding mask), softmax in fp32, o_proj without bias."""
    b, T, _ = x.shape
    H, KV, D = cfg.heads, cfg.kv_heads, cfg.head_dim
    q = F.linear(x, sd[pre + "self_attn.q_proj.weight"], sd[pre + "self_attn.q_proj.bias"]).view(b, T, H, D).transpose(1, 2)
    k = F.linear(x, sd[pre + "self_attn.k_proj.weight"], sd[pre + "self_attn.k_proj.bias"]).view(b, T, KV, D).transpose(1, 2)
    v = F.linear(x, sd[pre + "self_attn.v_proj.weight"], sd[pre + "self_attn.v_proj.bias"]).view(b, T, KV, D).transpose(1, 2)
    q, k = apply_rope(q, k, cos, sin)
    if past is not None:
        k = torch.cat([past[0], k], dim=2)
        v = torch.cat([past[1], v], dim=2)
    new_past = (k, v)
    ctx = k.shape[2]
    kr = k[:, :, None].expand(b, KV, H // KV, ctx, D).reshape(b, H, ctx, D)
    vr = v[:, :, None].expand(b, KV, H // KV, ctx, D).reshape(b, H, ctx, D)
    w = (q @ kr.transpose(2, 3)) * (D ** -0.5)
    qpos = torch.arange(ctx - T, ctx, device=q.device)[:, None]
    causal = torch.arange(ctx, device=q.device)[None, :] <= qpos  # [T, ctx]
    allow = causal[None, None]
    if key_mask is not None:
        allow = allow & key_mask[:, None, None, :].bool()
    w = w.masked_fill(~allow, torch.finfo(w.dtype).min)
    w = torch.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
    out = (w @ vr).transpose(1, 2).reshape(b, T, H * D)
    return F.linear(out, sd[pre + "self_attn.o_proj.weight"]), new_past


def qwen2_mlp(x: torch.Tensor, sd, pre: str) -> torch.Tensor:
    """Qwen2MLP.forward modeling_qwen2.py:46-48: down(silu(gate(x)) * up(x)), no biases."""
    return F.linear(F.silu(F.linear(x, sd[pre + "mlp.gate_proj.weight"])) * F.linear(x, sd[pre + "mlp.up_proj.weight"]),
                    sd[pre + "mlp.down_proj.weight"])


def moe_layer_is_sparse(li: int, cfg: OracleConfig) -> bool:
    """Qwen2MoeDecoderLayer.__init__ transformers modeling_qwen2_moe.py:381-386."""
    return (li not in cfg.mlp_only_layers) and cfg.num_experts > 0 and (li + 1) % cfg.decoder_sparse_step == 0


ROUTE_TRACE: Optional[list] = None  # tests set this to a list: every moe_route call appends log(p_k / p_(k+1)) per row


def moe_route(x: torch.Tensor, sd, pre: str, cfg: OracleConfig):
    """Qwen2MoeTopKRouter.forward modeling_qwen2_moe.py:343-352: softmax over ALL experts in fp32, top-k, optional
    renormalisation of the k weights. x [T, C] -> (weights [T, k], expert ids [T, k])."""
    probs = torch.softmax(F.linear(x, sd[pre + "mlp.gate.weight"]), dim=-1, dtype=torch.float32)
    w, idx = torch.topk(probs, cfg.top_k, dim=-1)
    if ROUTE_TRACE is not None and cfg.top_k < cfg.num_experts:
        # how far the routing decision is from flipping: a bf16 run may legitimately pick the other expert when this is ~ 0
        srt = torch.sort(probs, dim=-1, descending=True).values
        ROUTE_TRACE.append(torch.log(srt[:, cfg.top_k - 1] / srt[:, cfg.top_k].clamp_min(1e-30)))
    if cfg.norm_topk_prob:
        w = w / w.sum(dim=-1, keepdim=True)
    return w.to(x.dtype), idx


def qwen2_moe_block(x: torch.Tensor, sd, pre: str, cfg: OracleConfig) -> torch.Tensor:
    """Qwen2MoeSparseMoeBlock.forward modeling_qwen2_moe.py:363-374 with Qwen2MoeExperts.forward :307-331: every token goes
    through its k routed experts (SwiGLU MLPs of moe_intermediate_size, weighted by the routing weights) plus the shared
    expert scaled by sigmoid(shared_expert_gate(x)). Expert weights under their checkpoint names (one matrix per expert)."""
    shp = x.shape
    x = x.reshape(-1, shp[-1])
    w, idx = moe_route(x, sd, pre, cfg)
    out = torch.zeros_like(x)
    for e in range(cfg.num_experts):
        tok, slot = torch.where(idx == e)
        if tok.numel() == 0:
            continue
        q = f"{pre}mlp.experts.{e}."
        y = F.linear(F.silu(F.linear(x[tok], sd[q + "gate_proj.weight"])) * F.linear(x[tok], sd[q + "up_proj.weight"]),
                     sd[q + "down_proj.weight"])
        out.index_add_(0, tok, y * w[tok, slot, None])
    q = pre + "mlp.shared_expert."
    shared = F.linear(F.silu(F.linear(x, sd[q + "gate_proj.weight"])) * F.linear(x, sd[q + "up_proj.weight"]),
                      sd[q + "down_proj.weight"])
    out = out + torch.sigmoid(F.linear(x, sd[pre + "mlp.shared_expert_gate.weight"])) * shared
    return out.reshape(shp)


def qwen2_forward(embeds: torch.Tensor, position_ids: torch.Tensor, sd, cfg: OracleConfig,
                  past: Optional[List[Tuple[torch.Tensor, torch.Tensor]]] = None,
                  key_mask: Optional[torch.Tensor] = None, return_hidden: bool = False):
    """Qwen2Model.forward modeling_qwen2.py:353-414 + Qwen2DecoderLayer.forward :280-310 + lm_head :470-472.
    embeds [b,T,C]; position_ids [b,T]; key_mask [b, ctx] (True = attend) covers past + current keys.
    Returns (logits [b,T,V], new_past[, hidden states per layer])."""
    h = embeds
    cos, sin = rope_cos_sin(position_ids, cfg, embeds.dtype)
    new_past = []
    hiddens = [h]
    for li in range(cfg.layers):
        pre = f"model.layers.{li}."
        a, kv = qwen2_attention(rms_norm(h, sd[pre + "input_layernorm.weight"], cfg.rms_eps), sd, pre, cfg, cos, sin,
                                None if past is None else past[li], key_mask)
        new_past.append(kv)
        h = h + a
        xn = rms_norm(h, sd[pre + "post_attention_layernorm.weight"], cfg.rms_eps)
        h = h + (qwen2_moe_block(xn, sd, pre, cfg) if moe_layer_is_sparse(li, cfg) else qwen2_mlp(xn, sd, pre))
        hiddens.append(h)
    h = rms_norm(h, sd["model.norm.weight"], cfg.rms_eps)
    logits = F.linear(h, sd["lm_head.weight"])
    if return_hidden:
        return logits, new_past, hiddens
    return logits, new_past


# ----------------------------------------------------------------------------------------------- boundary
def forward_multimodal(input_ids: torch.Tensor, images: Optional[torch.Tensor], sd, cfg: OracleConfig,
                       attention_mask: Optional[torch.Tensor] = None):
    """OmChatQwen2ForCausalLM.forward omchat_qwen2.py:45-89 for a prefill call: glue (omchat_arch.py:55-209) then
    Qwen2ForCausalLM.forward on inputs_embeds. Returns (logits [b,T,V], past, mask [b,T], lengths)."""
    table = sd["model.embed_tokens.weight"]
    if images is None:
        embeds = table[input_ids]
        b, T = input_ids.shape
        mask = torch.ones(b, T, dtype=torch.bool) if attention_mask is None else attention_mask.bool()
        pos = torch.arange(T)[None].expand(b, T)
        lens = [T] * b
    else:
        feats = encode_images(images, sd, cfg)
        embeds, mask, pos, lens = splice(input_ids, attention_mask, feats, table, cfg)
    logits, past = qwen2_forward(embeds, pos, sd, cfg, None, mask)
    return logits, past, mask, lens


def greedy_generate(input_ids: torch.Tensor, images: Optional[torch.Tensor], sd, cfg: OracleConfig, max_new_tokens: int,
                    eos_token_id: Optional[int] = None):
    """Greedy loop as driven by cli.py:60-70 (do_sample=False, use_cache=True), batch 1 (the reference's generate()
    is broken on transformers 5.5.0, SURVEY.md §8c, so this is the manual loop through forward). Returns
    (new token ids list, list of last-position logits per step)."""
    assert input_ids.shape[0] == 1
    logits, past, mask, lens = forward_multimodal(input_ids, images, sd, cfg)
    T = lens[0]
    out, step_logits = [], []
    last = logits[0, T - 1]
    table = sd["model.embed_tokens.weight"]
    for step in range(max_new_tokens):
        step_logits.append(last.clone())
        tok = int(torch.argmax(last))
        out.append(tok)
        if eos_token_id is not None and tok == eos_token_id:
            break
        if step == max_new_tokens - 1:
            break
        emb = table[torch.tensor([[tok]])]
        pos = torch.tensor([[T + step]])
        lg, past = qwen2_forward(emb, pos, sd, cfg, past, None)
        last = lg[0, 0]
    return out, step_logits
