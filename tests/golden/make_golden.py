"""Generate tests/golden/golden_tiny.pt by running the REAL reference (/root/reference) at the tiny configuration.

Run in the build container only (the GPU box has no /root/reference):  python tests/golden/make_golden.py
The reference is imported unmodified; three documented shims make it importable/runnable on CPU fp32 here
(SURVEY.md §8c): stub modules for timm/peft/accelerate (never executed on this path), InternVisionConfig replaced by
a tiny one with use_flash_attn=False, and the tower wrapper's hard fp16 cast (internVIT_encoder.py:53) replaced by
the weight dtype. Outputs are sub-sampled to keep the fixture small.
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from tiny import TINY, tiny_inputs, tiny_state_dict, weights_checksum  # noqa: E402


def import_reference():
    import transformers  # noqa: F401  (must be imported before the stubs are registered)

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m

    class _Never(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

    stub("timm"); stub("timm.models"); stub("timm.models.layers", DropPath=_Never)
    stub("timm.layers", LayerNorm=torch.nn.LayerNorm, LayerNorm2d=_Never)
    stub("timm.models.regnet", RegStage=_Never)
    stub("peft", PeftModel=_Never); stub("accelerate", Accelerator=_Never)
    sys.path.insert(0, "/root/reference")
    import omchat.model.multimodal_encoder.intern_vit_6b.configuration_intern_vit as cfgmod
    import omchat.model.multimodal_encoder.internVIT_encoder as enc
    from omchat.model import OmChatQwen2Config, OmChatQwen2ForCausalLM

    orig = cfgmod.InternVisionConfig
    enc.InternVisionConfig = lambda *a, **k: orig(
        hidden_size=TINY["vit_hidden"], num_attention_heads=TINY["vit_heads"], intermediate_size=TINY["vit_inter"],
        num_hidden_layers=TINY["vit_layers"], image_size=TINY["image_size"], use_flash_attn=False)

    def tower_forward(self, images):  # internVIT_encoder.py:45-56 with the fp16 cast replaced by the weight dtype
        dt = self.vision_tower.embeddings.patch_embedding.weight.dtype
        outs = self.vision_tower(images.to(device=self.device, dtype=dt), output_hidden_states=True)
        return self.feature_select(outs).to(images.dtype)

    enc.InternVITVisionTower.forward = tower_forward
    cfg = OmChatQwen2Config(
        mm_vision_tower="InternViT-6B-448px-V1-5", mm_projector_type="mlp2x_gelu", mm_hidden_size=TINY["vit_hidden"],
        mm_vision_select_layer=-1, mm_vision_select_feature="patch", delay_load=False, hidden_size=TINY["hidden"],
        intermediate_size=TINY["inter"], num_hidden_layers=TINY["layers"], num_attention_heads=TINY["heads"],
        num_key_value_heads=TINY["kv_heads"], vocab_size=TINY["vocab"], max_position_embeddings=8192,
        rope_theta=TINY["rope_theta"], rms_norm_eps=1e-6, attn_implementation="eager")
    model = OmChatQwen2ForCausalLM(cfg).eval()
    return model


@torch.no_grad()
def main():
    torch.manual_seed(0)
    model = import_reference()
    sd = tiny_state_dict(0)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("inv_freq" in m or "rotary" in m for m in missing), missing
    pixels, ids = tiny_inputs(1)
    out = {"weights_checksum": weights_checksum(sd), "torch": torch.__version__}

    # A. tower: all hidden states + selected features
    tower = model.get_vision_tower()
    vout = tower.vision_tower(pixels[:2], output_hidden_states=True)
    out["vit_hidden_states_sub"] = [h[:, ::16, ::4].clone() for h in vout.hidden_states]
    feats = tower(pixels[:2])
    out["vit_features_sub"] = feats[:, ::8, :].clone()
    # B. tower + projector
    proj = model.encode_images(pixels[:2])
    out["encode_images_sub"] = proj[:, ::8, :].clone()

    # C. multimodal prefill, batch 1, one placeholder
    ids_c = ids[:1].clone()
    ids_c[0, 5] = -200
    res = model(input_ids=ids_c, images=pixels[:1], use_cache=True)
    out["prefill_ids"] = ids_c
    out["prefill_logits_sub"] = res.logits[0, ::16, :].clone()
    out["prefill_logits_last"] = res.logits[0, -1, :].clone()
    # D. manual greedy loop through the reference forward (generate() is broken on transformers 5.5.0)
    past = res.past_key_values
    last = res.logits[0, -1]
    toks, margins = [], []
    for step in range(8):
        top2 = torch.topk(last, 2).values
        margins.append(float(top2[0] - top2[1]))
        tok = int(torch.argmax(last))
        toks.append(tok)
        r2 = model(input_ids=torch.tensor([[tok]]), past_key_values=past, use_cache=True)
        past = r2.past_key_values
        last = r2.logits[0, -1]
    out["greedy_tokens"] = toks
    out["greedy_margins"] = margins
    # D2. north_star: greedy ids must match for the first 32 generated tokens. Random-init logits are nearly flat, so the
    # prompt is SEARCHED (seeds 100, 101, ...) for the one whose smallest of 32 top-1 margins is largest relative to the logit scale (>= 2 %, bf16 noise at this
    # depth is ~0.3 %): then a
    # mismatch on the bf16 path is an error, not a tie, and the GPU test can assert the 32 ids without a margin escape.
    best = None
    for seed in range(100, 260):
        g = torch.Generator().manual_seed(seed)
        ids_d = torch.randint(0, TINY["vocab"] - 10, (1, 24), generator=g)
        ids_d[0, 7] = -200
        res = model(input_ids=ids_d, images=pixels[1:2], use_cache=True)
        past, last = res.past_key_values, res.logits[0, -1]
        toks32, margins32, scales32 = [], [], []
        for step in range(32):
            top2 = torch.topk(last, 2).values
            margins32.append(float(top2[0] - top2[1]))
            scales32.append(float(last.abs().max()))
            tok = int(torch.argmax(last))
            toks32.append(tok)
            r2 = model(input_ids=torch.tensor([[tok]]), past_key_values=past, use_cache=True)
            past, last = r2.past_key_values, r2.logits[0, -1]
        worst = min(m / s for m, s in zip(margins32, scales32))
        if best is None or worst > best[0]:
            best = (worst, seed, ids_d, toks32, margins32, scales32)
    worst, seed, ids_d, toks32, margins32, scales32 = best
    assert worst >= 0.02, f"no prompt with 32 comfortable margins found (best {worst})"
    out["greedy32_seed"] = seed
    out["greedy32_ids"] = ids_d
    out["greedy32_image"] = 1
    out["greedy32_tokens"] = toks32
    out["greedy32_margins"] = margins32
    out["greedy32_scales"] = scales32
    print("greedy32 seed", seed, toks32, "min margin/scale", min(m / s for m, s in zip(margins32, scales32)))

    # E. splice semantics: batch of 3 with padding, 2 / 0 / 1 placeholders (the image-less row still consumes a block)
    ids_e = ids.clone()
    ids_e[0, 3] = -200
    ids_e[0, 17] = -200
    ids_e[2, 0] = -200
    mask_e = torch.ones_like(ids_e, dtype=torch.bool)
    mask_e[1, 20:] = False
    mask_e[2, 22:] = False
    for side in ("right", "left"):
        model.config.tokenizer_padding_side = side
        for max_len in (None, 300):
            model.config.tokenizer_model_max_length = max_len
            _, pos, am, _, emb, _ = model.prepare_inputs_labels_for_multimodal(
                ids_e, torch.arange(ids_e.shape[1]), mask_e, None, None, pixels)
            key = f"splice_{side}_{max_len}"
            out[key + "_pos"] = pos.clone()
            out[key + "_mask"] = am.clone()
            out[key + "_embeds_sub"] = emb[:, :, ::32].clone()
    model.config.tokenizer_padding_side = "right"
    model.config.tokenizer_model_max_length = None
    out["splice_ids"] = ids_e
    out["splice_mask"] = mask_e
    # F. batched multimodal prefill logits with padding (right)
    res_f = model(input_ids=ids_e, attention_mask=mask_e, images=pixels, use_cache=False)
    out["batch_logits_sub"] = res_f.logits[:, ::32, ::4].clone()

    path = os.path.join(HERE, "golden_tiny.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")
    print("greedy", toks, "margins", [round(m, 4) for m in margins])


if __name__ == "__main__":
    main()
