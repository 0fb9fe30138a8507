"""CPU: checkpoint layouts of SURVEY.md §8f-1 — safetensors round trip in both parameter-name layouts
(omchat/model/builder.py:22-35, convert_omchat_to_hf.py:26-35) and config.json mapping. No kernels involved."""
import json
import os

import torch

from omchat_b200.config import OmChatQwen2Config
from omchat_b200.model import checkpoint as ck
from omchat_b200.model.weights import from_hf_names, tp_plan


def test_roundtrip_both_layouts(tmp_path, tiny_sd):
    cfg = OmChatQwen2Config(vocab_size=512, hidden_size=256, num_hidden_layers=2, kv_page_size=16)
    sd = dict(tiny_sd)
    sd["model.layers.0.self_attn.rotary_emb.inv_freq"] = torch.ones(4)  # skipped on load like the converter does
    for hub in (False, True):
        d = str(tmp_path / ("hub" if hub else "omchat"))
        ck.save_checkpoint(sd, cfg, d, hub_layout=hub, max_shard_bytes=1 << 20)
        assert len([f for f in os.listdir(d) if f.endswith(".safetensors")]) > 1
        got, cfg2 = ck.load_checkpoint(d)
        if hub:
            assert any(k.startswith("language_model.model.layers.") for k in got)
            assert any(k.startswith("multi_modal_projector.linear_1.") for k in got)
            assert any(k.startswith("vision_tower.encoder.layers.") for k in got)
            got = from_hf_names(got)
        assert not any(k.endswith("inv_freq") for k in got)
        assert set(got) == set(tiny_sd)
        for k, v in tiny_sd.items():
            assert torch.equal(got[k], v), k
        assert (cfg2.vocab_size, cfg2.hidden_size, cfg2.num_hidden_layers, cfg2.kv_page_size) == (512, 256, 2, 16)
        assert cfg2.vision_config.hidden_size == cfg.vision_config.hidden_size


def test_config_from_hub_json():
    d = {"text_config": {"vocab_size": 1000, "hidden_size": 128, "num_attention_heads": 1, "num_key_value_heads": 1},
         "vision_config": {"hidden_size": 256, "num_hidden_layers": 3, "bogus": 1}, "vision_feature_layer": -2,
         "eos_token_id": [7, 8]}
    c = ck.config_from_dict(json.loads(json.dumps(d)))
    assert c.vocab_size == 1000 and c.head_dim == 128 and c.mm_vision_select_layer == -2 and c.eos_token_id == 7
    assert c.vision_config.num_hidden_layers == 3 and c.mm_hidden_size == 256


def test_tp_plan_covers_every_head_and_row_once():
    cfg = OmChatQwen2Config()
    for size in (1, 2, 4, 8):
        plans = [tp_plan(cfg, r, size) for r in range(size)]
        q = sorted(h for p in plans for h in p.q_heads if h >= 0)
        assert q == list(range(cfg.num_attention_heads))  # each q head on exactly one rank
        assert len({len(p.q_heads) for p in plans}) == 1
        assert sorted(set(h for p in plans for h in p.kv_heads)) == list(range(cfg.num_key_value_heads))
        for p in plans:  # a rank's q heads all belong to kv heads it holds
            assert all(h < 0 or h // 7 in p.kv_heads for h in p.q_heads)
            assert (p.i_hi - p.i_lo + p.i_pad) % 8 == 0
        assert [p.i_lo for p in plans] == [r * cfg.intermediate_size // size for r in range(size)]
        assert plans[-1].i_hi == cfg.intermediate_size and plans[-1].v_hi == cfg.vocab_size
