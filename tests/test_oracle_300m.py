"""CPU: pin the oracle's InternViT-300M branch (LayerNorm, no QK-norm, qkv bias, 64-dim heads) against outputs of the REAL
reference's 300M tower (internVIT300m_encoder.py + intern_vit_300m/) stored by tests/golden/make_golden_300m.py."""
import os

import pytest
import torch

from oracle import omchat_oracle as O
from tiny import TINY_300M as T, tiny_inputs, tiny_state_dict_300m, weights_checksum

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def golden300():
    return torch.load(os.path.join(HERE, "golden", "golden_tiny_300m.pt"), weights_only=False)


@pytest.fixture(scope="module")
def sd300():
    return tiny_state_dict_300m(0)


def cfg300(**kw):
    c = dict(vit_hidden=T["vit_hidden"], vit_heads=T["vit_heads"], vit_inter=T["vit_inter"], vit_layers=T["vit_layers"],
             image_size=T["image_size"], hidden=T["hidden"], heads=T["heads"], kv_heads=T["kv_heads"], inter=T["inter"],
             layers=T["layers"], vocab=T["vocab"], rope_theta=T["rope_theta"], vit_norm_type="layer_norm", vit_qk_norm=False)
    c.update(kw)
    return O.OracleConfig(**c)


def close(a, b, tol=2e-4):
    a, b = a.float(), b.float()
    err, ref = (a - b).abs().max().item(), b.abs().max().item()
    assert err <= tol * max(ref, 1.0), f"max abs err {err} (ref scale {ref})"


def test_weights_reproducible(golden300, sd300):
    assert abs(weights_checksum(sd300) - golden300["weights_checksum"]) < 1e-6 * golden300["weights_checksum"]


def test_300m_tower_matches_reference(golden300, sd300):
    pixels, _ = tiny_inputs(1)
    feats, states = O.vit_tower(pixels[:2], sd300, cfg300(), return_all=True)
    assert len(states) == len(golden300["vit_hidden_states_sub"]) == T["vit_layers"] + 1
    for mine, ref in zip(states, golden300["vit_hidden_states_sub"]):
        close(mine[:, ::16, ::4], ref)
    close(feats[:, ::8, :], golden300["vit_features_sub"])
    close(O.encode_images(pixels[:2], sd300, cfg300())[:, ::8, :], golden300["encode_images_sub"])


def test_300m_prefill_and_greedy_match_reference(golden300, sd300):
    pixels, _ = tiny_inputs(1)
    ids = golden300["prefill_ids"]
    logits, _, _, lens = O.forward_multimodal(ids, pixels[:1], sd300, cfg300())
    assert lens == [24 - 1 + 256]
    close(logits[0, ::16, :], golden300["prefill_logits_sub"], 5e-4)
    close(logits[0, -1, :], golden300["prefill_logits_last"], 5e-4)
    toks, _ = O.greedy_generate(ids, pixels[:1], sd300, cfg300(), max_new_tokens=8)
    assert toks == golden300["greedy_tokens"]


def test_rms_branch_is_the_default():
    assert O.OracleConfig().vit_norm_type == "rms_norm" and O.OracleConfig().vit_qk_norm
