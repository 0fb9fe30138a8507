"""CPU: pin the oracle (oracle/omchat_oracle.py) against outputs of the real reference stored in tests/golden."""
import torch
import pytest

from oracle import omchat_oracle as O
from tiny import TINY, tiny_inputs, weights_checksum


def tiny_cfg(**kw):
    c = dict(vit_hidden=TINY["vit_hidden"], vit_heads=TINY["vit_heads"], vit_inter=TINY["vit_inter"],
             vit_layers=TINY["vit_layers"], image_size=TINY["image_size"], hidden=TINY["hidden"], heads=TINY["heads"],
             kv_heads=TINY["kv_heads"], inter=TINY["inter"], layers=TINY["layers"], vocab=TINY["vocab"],
             rope_theta=TINY["rope_theta"])
    c.update(kw)
    return O.OracleConfig(**c)


def close(a, b, tol=2e-4):
    a, b = a.float(), b.float()
    err = (a - b).abs().max().item()
    ref = b.abs().max().item()
    assert err <= tol * max(ref, 1.0), f"max abs err {err} (ref scale {ref})"


def test_weights_reproducible(golden, tiny_sd):
    assert abs(weights_checksum(tiny_sd) - golden["weights_checksum"]) < 1e-6 * golden["weights_checksum"]


def test_vit_tower_matches_reference(golden, tiny_sd):
    pixels, _ = tiny_inputs(1)
    cfg = tiny_cfg()
    feats, states = O.vit_tower(pixels[:2], tiny_sd, cfg, return_all=True)
    assert len(states) == len(golden["vit_hidden_states_sub"]) == TINY["vit_layers"] + 1
    for mine, ref in zip(states, golden["vit_hidden_states_sub"]):
        close(mine[:, ::16, ::4], ref)
    close(feats[:, ::8, :], golden["vit_features_sub"])
    assert feats.shape == (2, 256, TINY["vit_hidden"])


def test_encode_images_matches_reference(golden, tiny_sd):
    pixels, _ = tiny_inputs(1)
    close(O.encode_images(pixels[:2], tiny_sd, tiny_cfg())[:, ::8, :], golden["encode_images_sub"])


def test_prefill_logits_and_greedy_match_reference(golden, tiny_sd):
    pixels, _ = tiny_inputs(1)
    cfg = tiny_cfg()
    ids = golden["prefill_ids"]
    logits, past, mask, lens = O.forward_multimodal(ids, pixels[:1], tiny_sd, cfg)
    assert lens == [24 - 1 + 256]
    close(logits[0, ::16, :], golden["prefill_logits_sub"], 5e-4)
    close(logits[0, -1, :], golden["prefill_logits_last"], 5e-4)
    toks, step_logits = O.greedy_generate(ids, pixels[:1], tiny_sd, cfg, max_new_tokens=8)
    assert toks == golden["greedy_tokens"]


def test_32_greedy_ids_match_reference(golden, tiny_sd):
    """north_star: 'greedy token IDs must match for the first 32 generated tokens' - the oracle against the REAL
    reference's manual greedy loop (tests/golden/make_golden.py D2; prompt searched for comfortable top-1 margins)."""
    pixels, _ = tiny_inputs(1)
    im = golden["greedy32_image"]
    toks, step_logits = O.greedy_generate(golden["greedy32_ids"], pixels[im:im + 1], tiny_sd, tiny_cfg(), max_new_tokens=32)
    assert len(golden["greedy32_tokens"]) == 32 and toks == golden["greedy32_tokens"]
    for lg, m in zip(step_logits, golden["greedy32_margins"]):
        top2 = torch.topk(lg, 2).values
        assert abs(float(top2[0] - top2[1]) - m) < 2e-3


@pytest.mark.parametrize("side", ["right", "left"])
@pytest.mark.parametrize("max_len", [None, 300])
def test_splice_matches_reference(golden, tiny_sd, side, max_len):
    pixels, _ = tiny_inputs(1)
    cfg = tiny_cfg(padding_side=side, max_len=max_len)
    feats = O.encode_images(pixels, tiny_sd, cfg)
    emb, mask, pos, lens = O.splice(golden["splice_ids"], golden["splice_mask"], feats,
                                    tiny_sd["model.embed_tokens.weight"], cfg)
    key = f"splice_{side}_{max_len}"
    assert torch.equal(mask, golden[key + "_mask"].bool())
    assert torch.equal(pos, golden[key + "_pos"])
    close(emb[:, :, ::32], golden[key + "_embeds_sub"])


def test_batched_prefill_matches_reference(golden, tiny_sd):
    pixels, _ = tiny_inputs(1)
    logits, _, mask, lens = O.forward_multimodal(golden["splice_ids"], pixels, tiny_sd, tiny_cfg(),
                                                 attention_mask=golden["splice_mask"])
    ref = golden["batch_logits_sub"]
    mine = logits[:, ::32, ::4]
    sel = mask[:, ::32]
    close(mine[sel], ref[sel], 5e-4)


def test_pixel_shuffle_closed_form():
    # the only function with no reference symbol: pin the view/permute chain against the closed-form index map
    B, G, C, d = 2, 8, 16, 2
    x = torch.arange(B * G * G * C, dtype=torch.float32).reshape(B, G * G, C)
    y = O.pixel_shuffle(x, d)
    assert y.shape == (B, (G // d) ** 2, C * d * d)
    for b in range(B):
        for i in range(G):
            for j in range(G):
                row = (i // d) * (G // d) + j // d
                col = (i % d) * d * C + (j % d) * C
                assert torch.equal(y[b, row, col:col + C], x[b, i * G + j])
    assert torch.equal(O.pixel_shuffle(x, 1), x)


def test_splice_plan_edge_cases():
    # empty rows, placeholder at the ends, truncation inside an image block
    plans = O.splice_plan([[-200, 5, -200], [], [7]], n_img=4, L=4, max_len=6)
    assert [len(p) for p in plans] == [6, 0, 1]
    assert plans[0][:5] == [(1, 0, 0), (1, 0, 1), (1, 0, 2), (1, 0, 3), (0, 5, 0)]
    assert plans[0][5] == (1, 1, 0)
    with pytest.raises(IndexError):
        O.splice_plan([[-200, -200]], n_img=1, L=2)
