"""Generate tests/golden/golden_tiny_moe.pt: the REAL reference's Qwen2-MoE wrapper (/root/reference
omchat/model/language_model/omchat_qwen2_moe.py over transformers' Qwen2MoeForCausalLM) at the tiny configuration.

Run in the build container only:  python tests/golden/make_golden_moe.py
Shims as make_golden.py. Two variants: A = every layer sparse, norm_topk_prob False (the published Qwen-MoE setting);
B = layer 0 dense (mlp_only_layers = [0]), norm_topk_prob True.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from tiny import TINY_MOE as T, fuse_experts_for_transformers5, tiny_inputs, tiny_state_dict_moe, weights_checksum  # noqa: E402
import make_golden  # noqa: E402


def build(dense_layers, norm_topk):
    make_golden.import_reference()  # stubs + tiny vision config + fp16-cast shim (idempotent)
    from omchat.model.language_model.omchat_qwen2_moe import OmChatQwen2MoeConfig, OmChatQwen2MoeForCausalLM
    cfg = OmChatQwen2MoeConfig(
        mm_vision_tower="InternViT-6B-448px-V1-5", mm_projector_type="mlp2x_gelu", mm_hidden_size=T["vit_hidden"],
        mm_vision_select_layer=-1, mm_vision_select_feature="patch", delay_load=False, hidden_size=T["hidden"],
        intermediate_size=T["inter"], num_hidden_layers=T["layers"], num_attention_heads=T["heads"],
        num_key_value_heads=T["kv_heads"], vocab_size=T["vocab"], max_position_embeddings=8192, rope_theta=T["rope_theta"],
        rms_norm_eps=1e-6, attn_implementation="eager", num_experts=T["num_experts"], num_experts_per_tok=T["top_k"],
        moe_intermediate_size=T["moe_inter"], shared_expert_intermediate_size=T["shared_inter"], norm_topk_prob=norm_topk,
        decoder_sparse_step=1, mlp_only_layers=list(dense_layers), qkv_bias=True, tie_word_embeddings=False)
    model = OmChatQwen2MoeForCausalLM(cfg).eval()
    sd = tiny_state_dict_moe(0, dense_layers)
    missing, unexpected = model.load_state_dict(fuse_experts_for_transformers5(sd, T["num_experts"]), strict=False)
    assert not unexpected, unexpected
    assert all("inv_freq" in m or "rotary" in m for m in missing), missing
    return model, sd


@torch.no_grad()
def run(model, pixels, ids):
    out = {}
    ids_c = ids[:1].clone()
    ids_c[0, 5] = -200
    res = model(input_ids=ids_c, images=pixels[:1], use_cache=True)
    out["prefill_ids"] = ids_c
    out["prefill_logits_sub"] = res.logits[0, ::16, :].clone()
    out["prefill_logits_last"] = res.logits[0, -1, :].clone()
    past, last = res.past_key_values, res.logits[0, -1]
    toks, margins, scales = [], [], []
    for _ in range(8):
        top2 = torch.topk(last, 2).values
        margins.append(float(top2[0] - top2[1]))
        scales.append(float(last.abs().max()))
        tok = int(torch.argmax(last))
        toks.append(tok)
        r2 = model(input_ids=torch.tensor([[tok]]), past_key_values=past, use_cache=True)
        past, last = r2.past_key_values, r2.logits[0, -1]
    out["greedy_tokens"], out["greedy_margins"], out["greedy_scales"] = toks, margins, scales
    # a padded batch of 3 (2 / 0 / 1 image placeholders) through the same forward
    ids_e = ids.clone()
    ids_e[0, 3] = -200
    ids_e[0, 17] = -200
    ids_e[2, 0] = -200
    mask_e = torch.ones_like(ids_e, dtype=torch.bool)
    mask_e[1, 20:] = False
    mask_e[2, 22:] = False
    res_f = model(input_ids=ids_e, attention_mask=mask_e, images=pixels, use_cache=False)
    out["batch_ids"], out["batch_mask"] = ids_e, mask_e
    out["batch_logits_sub"] = res_f.logits[:, ::32, ::4].clone()
    return out


def main():
    torch.manual_seed(0)
    pixels, ids = tiny_inputs(1)
    out = {"torch": torch.__version__}
    for name, dense, norm in (("A", (), False), ("B", (0,), True)):
        model, sd = build(dense, norm)
        o = run(model, pixels, ids)
        o["weights_checksum"] = weights_checksum(sd)
        o["dense_layers"], o["norm_topk_prob"] = list(dense), norm
        out[name] = o
        print(name, "greedy", o["greedy_tokens"], "margin/scale", [round(m / s, 4) for m, s in zip(o["greedy_margins"], o["greedy_scales"])])
    path = os.path.join(HERE, "golden_tiny_moe.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
