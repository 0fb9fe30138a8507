#!/usr/bin/env python
"""InternViT-300M tower (24 layers, hidden 1024, 16 heads of 64 run zero-padded to 128, LayerNorm) on `--crops` 448 px crops:
crops/s and TFLOP/s on the ALGORITHMIC flops (64-dim heads; the padded attention executes 2x the QK^T / PV flops)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from omchat_b200 import lib  # noqa: E402
from omchat_b200.config import InternVisionConfig, OmChatQwen2Config  # noqa: E402
from omchat_b200.model.vision import build_vision_tower  # noqa: E402
from omchat_b200.model.weights import random_init  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--crops", type=int, default=64)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    print(json.dumps(measure(a.crops, a.iters)))


def measure(crops=64, iters=5):
    class a:  # noqa: N801
        pass
    a.crops, a.iters = crops, iters
    lib.load()
    vc = InternVisionConfig.intern_vit_300m()
    cfg = OmChatQwen2Config(vision_config=vc, mm_vision_tower="InternViT-300M-448px", mm_hidden_size=1024, hidden_size=256,
                            intermediate_size=512, num_hidden_layers=1, num_attention_heads=2, num_key_value_heads=1, vocab_size=1000)
    w = random_init(cfg, device="cuda", seed=0, text=False)
    tower = build_vision_tower(cfg, w.vit)
    px = torch.randn(a.crops, 3, 448, 448, device="cuda").to(torch.bfloat16)
    for _ in range(2):
        tower(px)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = lib.launch_count()
    e0.record()
    for _ in range(a.iters):
        tower(px)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    S, C, I, L = 1025, 1024, 4096, 24
    flops = a.crops * L * (2.0 * S * (4 * C * C + 2 * C * I) + 4.0 * S * S * C) + a.crops * 2.0 * 1024 * 588 * C
    return {"tower": "InternViT-300M-448px", "crops": a.crops, "ms": ms, "crops_per_sec": a.crops / ms * 1e3,
            "algorithmic_tflops": flops / ms / 1e9, "launches": (lib.launch_count() - n0) // a.iters}


if __name__ == "__main__":
    main()
