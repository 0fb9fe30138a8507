"""CPU: pin the oracle's Qwen2-MoE branch (router softmax / top-k / optional renormalisation, routed experts, sigmoid-gated
shared expert, dense mlp_only_layers) against outputs of the REAL reference's OmChatQwen2MoeForCausalLM
(omchat/model/language_model/omchat_qwen2_moe.py) stored by tests/golden/make_golden_moe.py."""
import os

import pytest
import torch

from oracle import omchat_oracle as O
from tiny import TINY_MOE as T, tiny_inputs, tiny_state_dict_moe, weights_checksum

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def golden_moe():
    return torch.load(os.path.join(HERE, "golden", "golden_tiny_moe.pt"), weights_only=False)


def moe_cfg(g, **kw):
    c = dict(vit_hidden=T["vit_hidden"], vit_heads=T["vit_heads"], vit_inter=T["vit_inter"], vit_layers=T["vit_layers"],
             image_size=T["image_size"], hidden=T["hidden"], heads=T["heads"], kv_heads=T["kv_heads"], inter=T["inter"],
             layers=T["layers"], vocab=T["vocab"], rope_theta=T["rope_theta"], num_experts=T["num_experts"], top_k=T["top_k"],
             norm_topk_prob=g["norm_topk_prob"], mlp_only_layers=tuple(g["dense_layers"]))
    c.update(kw)
    return O.OracleConfig(**c)


def close(a, b, tol=5e-4):
    a, b = a.float(), b.float()
    err, ref = (a - b).abs().max().item(), b.abs().max().item()
    assert err <= tol * max(ref, 1.0), f"max abs err {err} (ref scale {ref})"


@pytest.mark.parametrize("variant", ["A", "B"])
def test_moe_prefill_greedy_and_batch_match_reference(golden_moe, variant):
    g = golden_moe[variant]
    sd = tiny_state_dict_moe(0, g["dense_layers"])
    assert abs(weights_checksum(sd) - g["weights_checksum"]) < 1e-6 * g["weights_checksum"]
    cfg = moe_cfg(g)
    pixels, _ = tiny_inputs(1)
    logits, _, _, lens = O.forward_multimodal(g["prefill_ids"], pixels[:1], sd, cfg)
    assert lens == [24 - 1 + 256]
    close(logits[0, ::16, :], g["prefill_logits_sub"])
    close(logits[0, -1, :], g["prefill_logits_last"])
    toks, _ = O.greedy_generate(g["prefill_ids"], pixels[:1], sd, cfg, max_new_tokens=8)
    assert toks == g["greedy_tokens"]
    lb, _, mask, _ = O.forward_multimodal(g["batch_ids"], pixels, sd, cfg, attention_mask=g["batch_mask"])
    sel = mask[:, ::32]
    close(lb[:, ::32, ::4][sel], g["batch_logits_sub"][sel])


def test_sparse_layer_rule():
    c = O.OracleConfig(num_experts=8, decoder_sparse_step=2, mlp_only_layers=(3,), layers=6)
    assert [O.moe_layer_is_sparse(i, c) for i in range(6)] == [False, True, False, False, False, True]
    assert not O.moe_layer_is_sparse(1, O.OracleConfig())
