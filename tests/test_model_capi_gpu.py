"""Model-level C entry points (csrc/model_capi.cu: omc_vit_forward, omc_decoder_prefill - what a non-Python host binds,
INTEGRATION.md) against the Python host path that issues the same kernels one by one: bit-identical features / KV cache,
logits within summation-order tolerance (the Python path uses the GEMV kernel for lm_head on a few rows), at the tiny
configuration and at full InternViT-6B / Qwen2-7B width; and the oracle on top, so the entry points are pinned on their own."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _default_tile_configs(monkeypatch):
    """The C entry points use the library's default tile configuration; the Python path would pick one per shape by timing.
    With the folded RMSNorm the N-tile width decides how the rows' sums of squares are partitioned (fp32 partial sums), so
    bit-identity is only defined for equal configurations: switch the run-time tuning off for these comparisons."""
    from omchat_b200 import lib
    monkeypatch.setattr(lib, "GEMM_AUTOTUNE", False)

from oracle import omchat_oracle as O  # noqa: E402  (checker only)
from tiny import TINY, tiny_inputs, tiny_state_dict  # noqa: E402
from test_model_gpu import check, oracle_cfg, tiny_cfgs  # noqa: E402


@pytest.mark.parametrize("down", [1, 2])
def test_vit_forward_entry_point_tiny(down):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from omchat_b200 import lib
    from omchat_b200.model.omchat import OmChatQwen2ForCausalLM
    sd = {k: v.to(torch.bfloat16).float() for k, v in tiny_state_dict(0).items()}
    if down == 2:
        g = torch.Generator().manual_seed(3)
        sd["model.mm_projector.0.weight"] = (torch.randn(TINY["hidden"], TINY["vit_hidden"] * 4, generator=g) * 0.03).to(torch.bfloat16).float()
    cfg = tiny_cfgs(mm_pixel_shuffle_ratio=1.0 / down)
    m = OmChatQwen2ForCausalLM.from_state_dict(sd, cfg, device="cuda")
    pixels, _ = tiny_inputs(1)
    px = pixels[:3].cuda()
    tower = m.get_vision_tower()
    fwd = lib.VitForward(m.weights.vit, m.weights.proj, cfg.vision_config, down, folded=tower._folded())
    got = fwd(px)
    want = m.encode_images(px)  # the product's default: norm1 / norm2 folded into the GEMMs
    assert tower.fold_norms and got.shape == want.shape and torch.equal(got, want), \
        "C entry point and Python host path must produce the same bits"
    # the unfolded loop (stand-alone RMSNorm kernels) is the same math with the norm-weight rounding elsewhere
    plain = lib.VitForward(m.weights.vit, m.weights.proj, cfg.vision_config, down)(px)
    check(plain, want, "unfolded vs folded tower", rel=0.02)
    check(got, O.encode_images(pixels[:3], sd, oracle_cfg(pixel_shuffle_down=down)), f"omc_vit_forward vs oracle (down {down})")
    # capturable: no allocation / synchronisation inside the call
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fwd(px)
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out_g = fwd(px)
    out_g.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out_g, want)


def test_vit_forward_entry_point_full_width():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from omchat_b200 import lib
    from omchat_b200.config import InternVisionConfig, OmChatQwen2Config
    from omchat_b200.model.vision import InternVITVisionTower, MMProjector
    from omchat_b200.model.weights import random_init
    cfg = OmChatQwen2Config(vision_config=InternVisionConfig(num_hidden_layers=2))
    w = random_init(cfg, device="cuda", seed=0, text=False)
    px = torch.randn(2, 3, 448, 448, generator=torch.Generator().manual_seed(1)).cuda()
    tower = InternVITVisionTower(cfg, w.vit)
    want = MMProjector(w.proj)(tower(px))
    got = lib.VitForward(w.vit, w.proj, cfg.vision_config, 1, folded=tower._folded())(px)
    assert got.shape == (2, 1024, 3584) and torch.equal(got, want)


@pytest.mark.parametrize("full_width", [False, True])
def test_decoder_prefill_entry_point(full_width):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from omchat_b200 import lib
    from omchat_b200.config import OmChatQwen2Config
    from omchat_b200.model.decoder import Qwen2Decoder
    from omchat_b200.model.weights import random_init
    kw = {} if full_width else dict(hidden_size=256, num_attention_heads=2, num_key_value_heads=1, intermediate_size=512,
                                    vocab_size=1000, kv_page_size=16)
    cfg = OmChatQwen2Config(num_hidden_layers=2, **kw)
    w = random_init(cfg, device="cuda", seed=0, vision=False)
    dec = Qwen2Decoder(cfg, w.llm)
    lens = [70, 3, 129]
    T = sum(lens)
    g = torch.Generator(device="cuda").manual_seed(5)
    emb = (torch.randn(T, cfg.hidden_size, generator=g, device="cuda") * 0.02).to(torch.bfloat16)
    pos = torch.cat([torch.arange(n, dtype=torch.int32) for n in lens]).cuda()
    seq = torch.cat([torch.full((n,), i, dtype=torch.int32) for i, n in enumerate(lens)]).cuda()
    offs = [0, 70, 73, 202]
    cache_a, cache_b = dec.new_cache(3, 160), dec.new_cache(3, 160)
    want = dec.prefill(emb.clone(), pos, seq, offs, cache_a, logits="last")
    cu = torch.tensor(offs, dtype=torch.int32).cuda()
    last = torch.tensor([o - 1 for o in offs[1:]], dtype=torch.int64).cuda()
    got = lib.decoder_prefill(w.llm.layers, w.llm.norm, w.llm.lm_head,
                              (cfg.hidden_size, dec.Hq, dec.Hkv, dec.I_local, dec.V_local), dec.eps, dec.scale, dec.inv_freq,
                              emb.clone(), pos, seq, cu, max(lens), cache_b.pool, cache_b.block_table, cache_b.page_size, last,
                              folded=dec._folded_prefill())
    assert torch.equal(cache_a.pool, cache_b.pool), "the paged KV cache must hold the same bits"
    check(got, want, "omc_decoder_prefill logits vs Python host path", rel=2 ** -7)
