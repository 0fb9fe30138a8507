"""GPU parity of the InternViT-300M tower variant (LayerNorm + bias, no QK-norm, qkv bias, 16 x 64-dim heads on the head_dim-64
instantiation of the attention kernel; the zero-padded-to-128 path that serves other head widths is kept alive by one test)
through the same drop-in boundary, against (1) golden vectors of the REAL reference's 300M
tower (tests/golden/golden_tiny_300m.pt, fp32) and (2) the CPU oracle on bf16-rounded weights - at the tiny size and at the
real 300M width (hidden 1024, 16 heads, inter 4096, 448 px; reduced depth).

Tolerances as tests/test_model_gpu.py: cosine >= 0.999 per token, max-abs <= 4 % of the tensor's scale (bf16 vs fp32).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import omchat_oracle as O  # noqa: E402  (checker only)
from tiny import TINY_300M as T, tiny_inputs, tiny_state_dict_300m  # noqa: E402
from test_model_gpu import check  # noqa: E402


def cfg300():
    from omchat_b200.config import InternVisionConfig, OmChatQwen2Config
    vc = InternVisionConfig.intern_vit_300m(hidden_size=T["vit_hidden"], num_attention_heads=T["vit_heads"],
                                            intermediate_size=T["vit_inter"], num_hidden_layers=T["vit_layers"],
                                            image_size=T["image_size"], qkv_bias=True)
    return OmChatQwen2Config(vocab_size=T["vocab"], hidden_size=T["hidden"], intermediate_size=T["inter"],
                             num_hidden_layers=T["layers"], num_attention_heads=T["heads"], num_key_value_heads=T["kv_heads"],
                             rope_theta=T["rope_theta"], mm_hidden_size=T["vit_hidden"], kv_page_size=16, vision_config=vc,
                             eos_token_id=-1, mm_vision_tower="InternViT-300M-448px")


def oracle300(**kw):
    c = dict(vit_hidden=T["vit_hidden"], vit_heads=T["vit_heads"], vit_inter=T["vit_inter"], vit_layers=T["vit_layers"],
             image_size=T["image_size"], hidden=T["hidden"], heads=T["heads"], kv_heads=T["kv_heads"], inter=T["inter"],
             layers=T["layers"], vocab=T["vocab"], rope_theta=T["rope_theta"], vit_norm_type="layer_norm", vit_qk_norm=False)
    c.update(kw)
    return O.OracleConfig(**c)


@pytest.fixture(scope="module")
def golden300():
    import os
    return torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_tiny_300m.pt"),
                      weights_only=False)


@pytest.fixture(scope="module")
def sd_bf16():
    return {k: v.to(torch.bfloat16).float() for k, v in tiny_state_dict_300m(0).items()}


@pytest.fixture(scope="module")
def model(sd_bf16):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from omchat_b200.model.omchat import OmChatQwen2ForCausalLM
    return OmChatQwen2ForCausalLM.from_state_dict(sd_bf16, cfg300(), device="cuda")


def test_layernorm_kernel_vs_torch():
    from omchat_b200 import lib
    g = torch.Generator(device="cuda").manual_seed(3)
    for rows, C in ((1, 8), (5, 256), (1025, 1024), (77, 3200), (33, 4096)):
        x = (torch.randn(rows, C, generator=g, device="cuda") * 3 + 1.5).to(torch.bfloat16)
        w = (torch.randn(C, generator=g, device="cuda") * 0.1 + 1).to(torch.bfloat16)
        b = (torch.randn(C, generator=g, device="cuda") * 0.1).to(torch.bfloat16)
        for bias in (b, None):
            got = lib.layernorm(x, w, bias, 1e-6)
            ref = torch.nn.functional.layer_norm(x.float(), (C,), w.float(), None if bias is None else bias.float(), 1e-6)
            # fp32 statistics, ONE rounding: within half a bf16 ulp of the fp32 result (+ fp32 summation-order noise)
            err = (got.float() - ref).abs()
            assert (err <= ref.abs() * 2 ** -8 + 1e-5).all(), (rows, C, err.max().item())
    # strided rows (a column slice of a wider buffer), in place
    wide = torch.randn(64, 3 * 1024, generator=g, device="cuda").to(torch.bfloat16)
    w = torch.ones(1024, device="cuda", dtype=torch.bfloat16)
    ref = torch.nn.functional.layer_norm(wide[:, 1024:2048].float(), (1024,), None, None, 1e-6)
    keep = wide.clone()
    lib.layernorm(wide[:, 1024:2048], w, None, 1e-6, out=wide[:, 1024:2048])
    assert (wide[:, 1024:2048].float() - ref).abs().max().item() <= 2 ** -6
    assert torch.equal(wide[:, :1024], keep[:, :1024]) and torch.equal(wide[:, 2048:], keep[:, 2048:])


def test_tower_class_follows_the_name(model):
    from omchat_b200.model.vision import InternVIT300mVisionTower
    tower = model.get_vision_tower()
    assert isinstance(tower, InternVIT300mVisionTower) and not tower.fold_norms
    assert tower.hidden_size == T["vit_hidden"] and tower.vc.head_dim == 64 and tower.attn_dim == 64 and not tower.pad_heads


def test_300m_tower_vs_reference_golden(model, golden300):
    pixels, _ = tiny_inputs(1)
    tower = model.get_vision_tower()
    _, states = tower.hidden_states(pixels[:2].cuda(), collect=True)
    S = (T["image_size"] // 14) ** 2 + 1
    assert len(states) == len(golden300["vit_hidden_states_sub"])
    for li, (mine, ref) in enumerate(zip(states, golden300["vit_hidden_states_sub"])):
        check(mine.view(2, S, -1)[:, ::16, ::4], ref, f"300m hidden state {li}")
    check(tower(pixels[:2].cuda())[:, ::8, :], golden300["vit_features_sub"], "300m features")
    check(model.encode_images(pixels[:2])[:, ::8, :], golden300["encode_images_sub"], "300m encode_images")


def test_300m_tower_zero_padded_heads_path(model, golden300):
    """pad_heads = True: the heads run zero-padded to 128 dims (the path a tower with, say, 32-dim heads takes) - same golden."""
    pixels, _ = tiny_inputs(1)
    tower = model.get_vision_tower()
    native = tower(pixels[:2].cuda()).clone()
    tower.pad_heads = True
    try:
        assert tower.attn_dim == 128
        padded = tower(pixels[:2].cuda())
        check(padded[:, ::8, :], golden300["vit_features_sub"], "300m features (zero-padded heads)")
        check(padded, native, "zero-padded vs native head_dim 64", rel=2 ** -6)
    finally:
        tower.pad_heads = False


def test_300m_prefill_and_greedy_vs_reference_golden(model, golden300, sd_bf16):
    pixels, _ = tiny_inputs(1)
    ids = golden300["prefill_ids"]
    res = model(input_ids=ids, images=pixels[:1], use_cache=True)
    check(res.logits[0, ::16, :], golden300["prefill_logits_sub"], "300m prefill logits (sub)")
    check(res.logits[0, -1:, :], golden300["prefill_logits_last"][None], "300m prefill logits (last)")
    n = len(golden300["greedy_tokens"])
    out = model.generate(ids, images=pixels[:1], max_new_tokens=n, do_sample=False, eos_token_id=-1)
    got = out[0, ids.shape[1]:].tolist()
    want, step_logits = O.greedy_generate(ids, pixels[:1], sd_bf16, oracle300(), max_new_tokens=n)
    print("cuda:", got, "oracle:", want, "reference:", golden300["greedy_tokens"])
    for i in range(n):
        if got[i] != want[i]:  # teacher-free greedy: a flipped near-tie changes everything after it
            top2 = torch.topk(step_logits[i], 2).values
            margin, scale = float(top2[0] - top2[1]), float(step_logits[i].abs().max())
            assert margin < 0.02 * scale, f"greedy token {i}: {got[i]} vs {want[i]} (margin {margin:.4g}, scale {scale:.4g})"
            break


def test_300m_real_width_tower_vs_oracle():
    """hidden 1024 / 16 heads of 64 / inter 4096 / 448 px (1025 rows per crop), 3 of the 24 layers, 2 crops: the head_dim-64
    attention and the GEMM shapes of the real 300M tower against the fp32 oracle."""
    from omchat_b200.config import InternVisionConfig, OmChatQwen2Config
    from omchat_b200.model.vision import build_vision_tower
    from omchat_b200.model.weights import random_init, to_reference_state_dict
    vc = InternVisionConfig.intern_vit_300m(num_hidden_layers=3, qkv_bias=True)
    cfg = OmChatQwen2Config(vision_config=vc, mm_vision_tower="InternViT-300M-448px", mm_hidden_size=1024, hidden_size=256,
                            intermediate_size=512, num_hidden_layers=1, num_attention_heads=2, num_key_value_heads=1,
                            vocab_size=1000)
    w = random_init(cfg, device="cuda", seed=5, text=False)
    tower = build_vision_tower(cfg, w.vit)
    g = torch.Generator().manual_seed(9)
    pixels = torch.randn(2, 3, 448, 448, generator=g)
    _, states = tower.hidden_states(pixels.cuda(), collect=True)
    sd = {k: v.float().cpu() for k, v in to_reference_state_dict(w, cfg).items() if "vision_tower" in k}
    oc = O.OracleConfig(vit_hidden=1024, vit_heads=16, vit_inter=4096, vit_layers=3, image_size=448, vit_norm_type="layer_norm",
                        vit_qk_norm=False)
    _, ref_states = O.vit_tower(pixels.to(torch.bfloat16).float(), sd, oc, return_all=True)
    for li, (mine, ref) in enumerate(zip(states, ref_states)):
        check(mine.view(2, 1025, -1), ref, f"300m full-width hidden state {li}")


def test_300m_model_level_c_entry_equals_host_path(model, monkeypatch):
    """omc_vit_forward with norm_type = 1 / attn_head_dim = 128 / qkv_b: same kernels, same order, same bits as vision.py."""
    from omchat_b200 import lib
    monkeypatch.setattr(lib, "GEMM_AUTOTUNE", False, raising=False)
    pixels, _ = tiny_inputs(1)
    tower = model.get_vision_tower()
    want = model.encode_images(pixels[:3])
    fwd = lib.VitForward(model.weights.vit, model.weights.proj, model.config.vision_config, 1, mats=tower._layer_mats(),
                         mats_folded=tower.fold_norms)
    assert fwd.desc.attn_head_dim == 0 and fwd.desc.norm_type == 1
    got = fwd(pixels[:3].cuda().contiguous())
    assert got.shape == want.shape and torch.equal(got, want)
    # plain weights (no mats): the native head_dim-64 path needs no padded copies
    got2 = lib.VitForward(model.weights.vit, model.weights.proj, model.config.vision_config, 1)(pixels[:3].cuda().contiguous())
    assert torch.equal(got2, want)
    # and the zero-padded layout through the same entry (attn_head_dim = 128)
    tower.pad_heads = True
    try:
        fwd_p = lib.VitForward(model.weights.vit, model.weights.proj, model.config.vision_config, 1, mats=tower._layer_mats())
        assert fwd_p.desc.attn_head_dim == 128
        got_p = fwd_p(pixels[:3].cuda().contiguous())
        assert torch.equal(got_p, model.encode_images(pixels[:3]))
    finally:
        tower.pad_heads = False
