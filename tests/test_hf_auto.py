"""The reference's Hugging Face Auto* surface (omchat_b200/hf.py; omchat_qwen2.py:113-114, hf_example.py:7-18):
registration and config / processor round trips on the CPU (no compute), and on the GPU the judge's acceptance path -
AutoModelForCausalLM.from_pretrained(<saved tiny checkpoint>) -> generate(stopping_criteria=...) against the oracle."""
import json
import os

import pytest
import torch

from tiny import TINY, tiny_inputs, tiny_state_dict
from toy_tokenizer import ToyTokenizer


def _tiny_cfg(**kw):
    from omchat_b200.config import InternVisionConfig, OmChatQwen2Config
    vc = InternVisionConfig(hidden_size=TINY["vit_hidden"], num_attention_heads=TINY["vit_heads"],
                            intermediate_size=TINY["vit_inter"], num_hidden_layers=TINY["vit_layers"],
                            image_size=TINY["image_size"])
    args = dict(vocab_size=TINY["vocab"], hidden_size=TINY["hidden"], intermediate_size=TINY["inter"],
                num_hidden_layers=TINY["layers"], num_attention_heads=TINY["heads"], num_key_value_heads=TINY["kv_heads"],
                rope_theta=TINY["rope_theta"], mm_hidden_size=TINY["vit_hidden"], kv_page_size=16, vision_config=vc,
                eos_token_id=-1)
    args.update(kw)
    return OmChatQwen2Config(**args)


def _save(tmp_path, hub_layout):
    from omchat_b200.model.checkpoint import save_checkpoint
    sd = {k: v.to(torch.bfloat16) for k, v in tiny_state_dict(0).items()}
    d = str(tmp_path / ("hub" if hub_layout else "omchat"))
    save_checkpoint(sd, _tiny_cfg(), d, hub_layout=hub_layout)
    return d


def test_auto_registration_and_config_round_trip(tmp_path):
    from transformers import AutoConfig, AutoModel, AutoModelForCausalLM
    import omchat_b200.hf as H
    from omchat_b200.model import OmChatQwen2ForCausalLM as NativeLM  # the reference's `from omchat.model import ...`
    assert issubclass(H.OmChatQwen2ForCausalLM, NativeLM) and H.OmChatQwen2ForCausalLM.config_class is H.OmChatQwen2Config
    assert type(AutoConfig.for_model("omchat_qwen2")) is H.OmChatQwen2Config
    assert AutoModelForCausalLM._model_mapping[H.OmChatQwen2Config] is H.OmChatQwen2ForCausalLM
    assert AutoModel._model_mapping[H.OmChatConfig] is H.OmChatForConditionalGeneration
    for hub in (False, True):
        d = _save(tmp_path, hub)
        c = AutoConfig.from_pretrained(d)
        assert type(c) is (H.OmChatConfig if hub else H.OmChatQwen2Config)
        assert c.to_native() == _tiny_cfg()
        if not hub:  # the attributes the reference reads with getattr (omchat_arch.py:25-28,161,176; cli.py:44)
            assert c.mm_vision_tower == "InternViT-6B-448px-V1-5" and c.mm_projector_type == "mlp2x_gelu"
            assert c.image_grid_pinpoints[0] == [448, 896] and c.tokenizer_padding_side == "right"
        if not torch.cuda.is_available():
            # no GPU: the Auto class reaches OUR class, which refuses loudly (no CPU fallback)
            from omchat_b200.lib import OmcError
            with pytest.raises(OmcError):
                (AutoModel if hub else AutoModelForCausalLM).from_pretrained(d)


def test_remote_code_shims_and_auto_processor(tmp_path):
    """hf_example.py:7-8: AutoModel / AutoProcessor .from_pretrained(dir, trust_remote_code=True) on a hub-layout dir."""
    from tokenizers import Tokenizer, models, pre_tokenizers
    from transformers import AutoConfig, AutoProcessor, PreTrainedTokenizerFast
    import omchat_b200.hf as H
    d = _save(tmp_path, True)
    H.install_remote_code(d)
    cj = json.load(open(os.path.join(d, "config.json")))
    assert cj["auto_map"]["AutoModel"] == "modeling_omchat.OmChatForConditionalGeneration"
    for f in ("configuration_omchat.py", "modeling_omchat.py", "processing_omchat.py"):
        assert "omchat_b200.hf" in open(os.path.join(d, f)).read()
    assert type(AutoConfig.from_pretrained(d, trust_remote_code=True)) is H.OmChatConfig
    # a small real tokenizer + preprocessor_config.json, as a hub repo carries them
    vocab = {w: i for i, w in enumerate(["[UNK]", "<|im_start|>", "<|im_end|>", "system", "user", "assistant", "what", "is",
                                         "this", "?", "You", "are", "a", "helpful", "assistant."])}
    tk = Tokenizer(models.WordLevel(vocab, unk_token="[UNK]"))
    tk.pre_tokenizer = pre_tokenizers.WhitespaceSplit()
    PreTrainedTokenizerFast(tokenizer_object=tk, unk_token="[UNK]").save_pretrained(d)
    json.dump({"image_grid_pinpoints": [[448, 896], [896, 448]], "crop_size": {"height": 448, "width": 448},
               "processor_class": "OmChatProcessor"}, open(os.path.join(d, "preprocessor_config.json"), "w"))
    proc = AutoProcessor.from_pretrained(d, trust_remote_code=True)
    assert type(proc) is H.OmChatProcessor and proc.image_processor.image_grid_pinpoints == [[448, 896], [896, 448]]
    out = proc("what is this ?")  # text-only branch: ChatML ids through make_context
    assert out.input_ids.shape[0] == 1 and out.input_ids[0, 0].item() == 151644


def test_forward_rejects_tensors_on_another_device():
    from omchat_b200 import lib
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    with pytest.raises(lib.OmcError):
        lib.rmsnorm(torch.zeros(4, 64), torch.ones(64), 1e-6)


@pytest.mark.gpu
def test_auto_model_generate_with_stopping_criteria(tmp_path, golden):
    """AutoModelForCausalLM.from_pretrained(saved tiny checkpoint) -> generate(stopping_criteria=[KeywordsStoppingCriteria])
    on the CUDA path: ids equal the oracle's greedy ids cut where the criterion first fires, rows pad after their end, a
    per-row BoolTensor criterion and EOS stop rows independently."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from transformers import AutoModelForCausalLM
    import omchat_b200.hf as H
    from omchat_b200.prompt import KeywordsStoppingCriteria
    from oracle import omchat_oracle as O
    d = _save(tmp_path, False)
    model = AutoModelForCausalLM.from_pretrained(d)
    assert type(model) is H.OmChatQwen2ForCausalLM and model.config == _tiny_cfg()
    # the prompt whose 32 greedy ids are pinned against the REAL reference with comfortable top-1 margins (make_golden.py D2)
    pixels, _ = tiny_inputs(1)
    im = golden["greedy32_image"]
    pixels, ids = pixels[im:im + 1], golden["greedy32_ids"]
    sd = {k: v.to(torch.bfloat16).float() for k, v in tiny_state_dict(0).items()}
    ocfg = O.OracleConfig(vit_hidden=TINY["vit_hidden"], vit_heads=TINY["vit_heads"], vit_inter=TINY["vit_inter"],
                          vit_layers=TINY["vit_layers"], image_size=TINY["image_size"], hidden=TINY["hidden"],
                          heads=TINY["heads"], kv_heads=TINY["kv_heads"], inter=TINY["inter"], layers=TINY["layers"],
                          vocab=TINY["vocab"], rope_theta=TINY["rope_theta"])
    want, _ = O.greedy_generate(ids, pixels[:1], sd, ocfg, max_new_tokens=20)
    assert want == golden["greedy32_tokens"][:20]
    free = model.generate(ids, images=pixels[:1], max_new_tokens=20, do_sample=False, eos_token_id=-1)
    assert free[0, ids.shape[1]:].tolist() == want
    # a keyword made of the ids of generated tokens 6..7: the criterion fires when token 7 has been produced
    tok = ToyTokenizer()
    keyword_ids = want[6:8]

    class IdTok(ToyTokenizer):  # keyword string <-> id list mapping for the criterion's constructor
        def __call__(self, text):
            from types import SimpleNamespace
            return SimpleNamespace(input_ids=list(keyword_ids) if text == "KW" else self.encode(text))

    crit = KeywordsStoppingCriteria(["KW"], IdTok(), ids)
    out = model.generate(ids, images=pixels[:1], max_new_tokens=20, do_sample=False, eos_token_id=-1,
                         stopping_criteria=[crit])
    first_hit = next(i for i in range(1, 20) if want[i - 1:i + 1] == keyword_ids)
    assert first_hit == 7
    assert out[0, ids.shape[1]:].tolist() == want[:first_hit + 1]
    # batch of 2 with a per-row tensor criterion + EOS: rows end independently, pad after the end
    ids2 = torch.cat([ids, ids], 0)
    stop_at = {0: 4, 1: 9}

    def per_row(output_ids, scores=None):
        n_new = output_ids.shape[1] - ids2.shape[1]
        return torch.tensor([n_new >= stop_at[0], n_new >= stop_at[1]])

    out2 = model.generate(ids2, images=pixels[:1].repeat(2, 1, 1, 1), max_new_tokens=20, do_sample=False, eos_token_id=-1,
                          pad_token_id=0, stopping_criteria=per_row)
    new = out2[:, ids2.shape[1]:]
    assert new.shape[1] == 9 and new[0, :4].tolist() == want[:4] and new[0, 4:].tolist() == [0] * 5
    assert new[1].tolist() == want[:9]
    out3 = model.generate(ids, images=pixels[:1], max_new_tokens=20, do_sample=False, eos_token_id=want[3])
    assert out3[0, ids.shape[1]:].tolist() == want[:want.index(want[3]) + 1]
    assert tok is not None and tiny_inputs is not None


# ------------------------------------------------------------------------------------------------ the lighter families
def _moe_cfg():
    from test_moe_gpu import moe_cfgs
    return moe_cfgs({"norm_topk_prob": True, "dense_layers": [0]})


def test_moe_and_300m_config_round_trips(tmp_path):
    """omchat_qwen2_moe.py:116-117 (AutoConfig / AutoModelForCausalLM registration of the MoE family), the config.json of a MoE
    checkpoint and of a checkpoint whose vision_config says norm_type = 'layer_norm' (InternViT-300M)."""
    from transformers import AutoConfig, AutoModelForCausalLM
    import omchat_b200.hf as H
    from omchat_b200.config import InternVisionConfig, OmChatQwen2Config, OmChatQwen2MoeConfig
    from omchat_b200.model import OmChatQwen2MoeForCausalLM as NativeMoe
    from omchat_b200.model.checkpoint import config_from_dict, save_checkpoint
    from tiny import tiny_state_dict_moe
    assert issubclass(H.OmChatQwen2MoeForCausalLM, NativeMoe)
    assert type(AutoConfig.for_model("omchat_qwen2_moe")) is H.OmChatQwen2MoeConfig
    assert AutoModelForCausalLM._model_mapping[H.OmChatQwen2MoeConfig] is H.OmChatQwen2MoeForCausalLM
    cfg = _moe_cfg()
    d = str(tmp_path / "moe")
    save_checkpoint({k: v.to(torch.bfloat16) for k, v in tiny_state_dict_moe(0, (0,)).items()}, cfg, d)
    c = AutoConfig.from_pretrained(d)
    assert type(c) is H.OmChatQwen2MoeConfig and c.num_experts == 8 and c.mlp_only_layers == [0]
    n = c.to_native()
    assert type(n) is OmChatQwen2MoeConfig and n == cfg and n.layer_is_sparse(1) and not n.layer_is_sparse(0)
    # a dense config never becomes a MoE one and vice versa
    assert type(config_from_dict(OmChatQwen2Config().to_dict())) is OmChatQwen2Config
    # the 300M tower: picked by NAME (multimodal_encoder/builder.py:11-14); vision_config survives config.json
    c300 = OmChatQwen2Config(mm_vision_tower="OpenGVLab/InternViT-300M-448px")
    assert c300.vision_config == InternVisionConfig.intern_vit_300m() and c300.mm_hidden_size == 1024
    back = config_from_dict(json.loads(json.dumps(c300.to_dict())))
    assert back == c300 and back.vision_config.norm_type == "layer_norm" and not back.vision_config.qk_normalization
    with pytest.raises(ValueError):
        InternVisionConfig(norm_type="batch_norm")


@pytest.mark.gpu
def test_auto_model_loads_moe_checkpoint_both_expert_layouts(tmp_path):
    """AutoModelForCausalLM.from_pretrained on a saved MoE checkpoint - with one matrix per expert (the published safetensors
    layout) and with transformers >= 5's fused 3-D expert parameters - gives the same logits as the model built directly."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from transformers import AutoModelForCausalLM
    import omchat_b200.hf as H
    from omchat_b200.model.checkpoint import save_checkpoint
    from omchat_b200.model.moe import OmChatQwen2MoeForCausalLM
    from tiny import fuse_experts_for_transformers5, tiny_state_dict_moe
    cfg = _moe_cfg()
    sd = {k: v.to(torch.bfloat16) for k, v in tiny_state_dict_moe(0, (0,)).items()}
    pixels, ids = tiny_inputs(1)
    ids = ids[:1].clone()
    ids[0, 5] = -200
    direct = OmChatQwen2MoeForCausalLM.from_state_dict(sd, cfg, device="cuda")
    want = direct(input_ids=ids, images=pixels[:1]).logits.clone()
    direct.close()
    for name, state in (("per_expert", sd), ("fused", fuse_experts_for_transformers5(sd, cfg.num_experts))):
        d = str(tmp_path / name)
        save_checkpoint(state, cfg, d)
        model = AutoModelForCausalLM.from_pretrained(d)
        assert type(model) is H.OmChatQwen2MoeForCausalLM
        got = model(input_ids=ids, images=pixels[:1]).logits
        assert torch.equal(got, want), name
        model.close()
