"""CPU: host-side logic of the two lighter model families - weight layouts (zero-padded heads, stacked / interleaved experts,
both checkpoint layouts of the experts), the router GEMM operand, the sparse-layer rule. No kernels are called."""
import torch

from oracle import omchat_oracle as O
from tiny import TINY_MOE, fuse_experts_for_transformers5, tiny_state_dict_300m, tiny_state_dict_moe


def test_pad_heads_is_exact_in_fp32():
    """Zero-padded heads: q.k and P.V of the padded layout equal the unpadded ones exactly (the padding adds zeros only)."""
    from omchat_b200.model.weights import pad_head_cols, pad_head_rows
    g = torch.Generator().manual_seed(0)
    H, D, Dp, C, T = 3, 8, 16, 24, 5
    qkv_w = torch.randn(3 * H * D, C, generator=g)
    proj_w = torch.randn(C, H * D, generator=g)
    x = torch.randn(T, C, generator=g)

    def attn(qkv_w, proj_w, d):
        q, k, v = (x @ qkv_w.t()).view(T, 3, H, d).permute(1, 2, 0, 3)
        p = torch.softmax(q @ k.transpose(-1, -2) * D ** -0.5, dim=-1)
        return (p @ v).permute(1, 0, 2).reshape(T, H * d) @ proj_w.t()

    ref = attn(qkv_w, proj_w, D)
    got = attn(pad_head_rows(qkv_w, 3, H, D, Dp), pad_head_cols(proj_w, H, D, Dp), Dp)
    assert torch.allclose(ref, got, atol=1e-5)
    b = torch.randn(3 * H * D, generator=g)
    pb = pad_head_rows(b, 3, H, D, Dp)
    assert pb.shape == (3 * H * Dp,) and torch.equal(pb.view(3, H, Dp)[:, :, :D].reshape(-1), b) and not pb.view(3, H, Dp)[:, :, D:].any()


def test_moe_weights_round_trip_both_expert_layouts():
    from omchat_b200.config import InternVisionConfig, OmChatQwen2MoeConfig
    from omchat_b200.model.weights import from_state_dict, to_reference_state_dict
    T = TINY_MOE
    vc = InternVisionConfig(hidden_size=T["vit_hidden"], num_attention_heads=T["vit_heads"], intermediate_size=T["vit_inter"],
                            num_hidden_layers=T["vit_layers"], image_size=T["image_size"])
    cfg = OmChatQwen2MoeConfig(vocab_size=T["vocab"], hidden_size=T["hidden"], intermediate_size=T["inter"],
                               num_hidden_layers=T["layers"], num_attention_heads=T["heads"], num_key_value_heads=T["kv_heads"],
                               mm_hidden_size=T["vit_hidden"], vision_config=vc, num_experts=T["num_experts"],
                               num_experts_per_tok=T["top_k"], moe_intermediate_size=T["moe_inter"],
                               shared_expert_intermediate_size=T["shared_inter"], mlp_only_layers=[0])
    sd = {k: v.to(torch.bfloat16) for k, v in tiny_state_dict_moe(0, (0,)).items()}
    w = from_state_dict(sd, cfg, device="cpu")
    assert w.llm.layers[0].moe is None and w.llm.layers[0].gate_up_w is not None          # dense layer (mlp_only_layers)
    m = w.llm.layers[1].moe
    E, I, C = T["num_experts"], T["moe_inter"], T["hidden"]
    assert m.experts_gate_up.shape == (E * 2 * I, C) and m.experts_down.shape == (E * C, I) and m.router_w.shape == (E, C)
    # interleaved rows: gate_i, up_i alternate inside every expert's block
    e = 3
    blk = m.experts_gate_up.view(E, I, 2, C)[e]
    assert torch.equal(blk[:, 0], sd[f"model.layers.1.mlp.experts.{e}.gate_proj.weight"])
    assert torch.equal(blk[:, 1], sd[f"model.layers.1.mlp.experts.{e}.up_proj.weight"])
    back = to_reference_state_dict(w, cfg)
    for k, v in sd.items():
        if "rotary" in k or "inv_freq" in k:
            continue
        assert k in back and torch.equal(back[k].reshape(v.shape), v), k
    # transformers >= 5 fused 3-D expert parameters load to the same tensors
    w5 = from_state_dict(fuse_experts_for_transformers5(sd, E), cfg, device="cpu")
    m5 = w5.llm.layers[1].moe
    assert torch.equal(m5.experts_gate_up, m.experts_gate_up) and torch.equal(m5.experts_down, m.experts_down)


def test_router_gemm_operand_and_sparse_rule():
    from omchat_b200 import lib
    from omchat_b200.config import OmChatQwen2MoeConfig
    rw, sg = torch.randn(60, 64).to(torch.bfloat16), torch.randn(64).to(torch.bfloat16)
    cat = lib.router_cat(rw, sg)
    assert cat.shape == (128, 64) and torch.equal(cat[:60], rw) and torch.equal(cat[60], sg) and not cat[61:].any()
    assert lib.router_cat(torch.zeros(128, 64, dtype=torch.bfloat16), sg).shape == (256, 64)
    assert lib.router_cat(rw, None).shape == (128, 64)
    for step, only in ((1, []), (2, [3]), (3, [0, 5])):
        c = OmChatQwen2MoeConfig(num_hidden_layers=8, decoder_sparse_step=step, mlp_only_layers=only)
        oc = O.OracleConfig(layers=8, num_experts=c.num_experts, decoder_sparse_step=step, mlp_only_layers=tuple(only))
        assert [c.layer_is_sparse(i) for i in range(8)] == [O.moe_layer_is_sparse(i, oc) for i in range(8)]


def test_300m_weights_round_trip():
    from omchat_b200.config import InternVisionConfig, OmChatQwen2Config
    from omchat_b200.model.weights import from_state_dict, to_reference_state_dict
    from tiny import TINY_300M as T
    vc = InternVisionConfig.intern_vit_300m(hidden_size=T["vit_hidden"], num_attention_heads=T["vit_heads"],
                                            intermediate_size=T["vit_inter"], num_hidden_layers=T["vit_layers"],
                                            image_size=T["image_size"], qkv_bias=True)
    cfg = OmChatQwen2Config(vocab_size=T["vocab"], hidden_size=T["hidden"], intermediate_size=T["inter"],
                            num_hidden_layers=T["layers"], num_attention_heads=T["heads"], num_key_value_heads=T["kv_heads"],
                            mm_hidden_size=T["vit_hidden"], vision_config=vc, mm_vision_tower="InternViT-300M-448px")
    sd = {k: v.to(torch.bfloat16) for k, v in tiny_state_dict_300m(0).items()}
    w = from_state_dict(sd, cfg, device="cpu")
    l = w.vit.layers[0]
    assert l.q_norm is None and l.k_norm is None and l.norm1_b is not None and l.qkv_b is not None
    back = to_reference_state_dict(w, cfg)
    for k, v in sd.items():
        if "vision_tower" in k:
            assert k in back and torch.equal(back[k].reshape(v.shape), v), k
    assert not any("q_norm" in k for k in back)
