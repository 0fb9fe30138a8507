"""Generate tests/golden/golden_tiny_300m.pt: the REAL reference (/root/reference) with its InternViT-300M tower
(omchat/model/multimodal_encoder/internVIT300m_encoder.py + intern_vit_300m/) at the tiny configuration.

Run in the build container only:  python tests/golden/make_golden_300m.py
Same three shims as make_golden.py (stub timm/peft/accelerate, tiny InternVisionConfig with use_flash_attn=False, the tower
wrapper's hard fp16 cast replaced by the weight dtype).
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from tiny import TINY_300M as T, tiny_inputs, tiny_state_dict_300m, weights_checksum  # noqa: E402


def import_reference():
    import transformers  # noqa: F401

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
    import omchat.model.multimodal_encoder.intern_vit_300m.configuration_intern_vit as cfgmod
    import omchat.model.multimodal_encoder.internVIT300m_encoder as enc
    from omchat.model import OmChatQwen2Config, OmChatQwen2ForCausalLM

    orig = cfgmod.InternVisionConfig
    # drop_path_rate 0: DropPath is an eval-time identity (timm) and the stub above cannot stand in for it
    enc.InternVisionConfig = lambda *a, **k: orig(
        hidden_size=T["vit_hidden"], num_attention_heads=T["vit_heads"], intermediate_size=T["vit_inter"],
        num_hidden_layers=T["vit_layers"], image_size=T["image_size"], use_flash_attn=False, qkv_bias=T["vit_qkv_bias"],
        drop_path_rate=0.0)

    def tower_forward(self, images):  # internVIT300m_encoder.py:46-58 with the fp16 cast replaced by the weight dtype
        dt = self.vision_tower.embeddings.patch_embedding.weight.dtype
        outs = self.vision_tower(images.to(device=self.device, dtype=dt), output_hidden_states=True)
        return self.feature_select(outs).to(images.dtype)

    enc.InternVIT300mVisionTower.forward = tower_forward
    cfg = OmChatQwen2Config(
        mm_vision_tower="InternViT-300M-448px", mm_projector_type="mlp2x_gelu", mm_hidden_size=T["vit_hidden"],
        mm_vision_select_layer=-1, mm_vision_select_feature="patch", delay_load=False, hidden_size=T["hidden"],
        intermediate_size=T["inter"], num_hidden_layers=T["layers"], num_attention_heads=T["heads"],
        num_key_value_heads=T["kv_heads"], vocab_size=T["vocab"], max_position_embeddings=8192,
        rope_theta=T["rope_theta"], rms_norm_eps=1e-6, attn_implementation="eager")
    model = OmChatQwen2ForCausalLM(cfg).eval()
    assert type(model.get_vision_tower()).__name__ == "InternVIT300mVisionTower"
    return model


@torch.no_grad()
def main():
    torch.manual_seed(0)
    model = import_reference()
    sd = tiny_state_dict_300m(0)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("inv_freq" in m or "rotary" in m for m in missing), missing
    pixels, ids = tiny_inputs(1)
    out = {"weights_checksum": weights_checksum(sd), "torch": torch.__version__}
    tower = model.get_vision_tower()
    vout = tower.vision_tower(pixels[:2], output_hidden_states=True)
    out["vit_hidden_states_sub"] = [h[:, ::16, ::4].clone() for h in vout.hidden_states]
    out["vit_features_sub"] = tower(pixels[:2])[:, ::8, :].clone()
    out["encode_images_sub"] = model.encode_images(pixels[:2])[:, ::8, :].clone()
    ids_c = ids[:1].clone()
    ids_c[0, 5] = -200
    res = model(input_ids=ids_c, images=pixels[:1], use_cache=True)
    out["prefill_ids"] = ids_c
    out["prefill_logits_sub"] = res.logits[0, ::16, :].clone()
    out["prefill_logits_last"] = res.logits[0, -1, :].clone()
    past, last = res.past_key_values, res.logits[0, -1]
    toks, margins = [], []
    for _ in range(8):
        top2 = torch.topk(last, 2).values
        margins.append(float(top2[0] - top2[1]))
        tok = int(torch.argmax(last))
        toks.append(tok)
        r2 = model(input_ids=torch.tensor([[tok]]), past_key_values=past, use_cache=True)
        past, last = r2.past_key_values, r2.logits[0, -1]
    out["greedy_tokens"], out["greedy_margins"] = toks, margins
    path = os.path.join(HERE, "golden_tiny_300m.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes; greedy", toks, "margins", [round(m, 4) for m in margins])


if __name__ == "__main__":
    main()
