#!/usr/bin/env python
"""Short, eager (no CUDA graph) runs of one phase of the hot path, meant to be wrapped in ncu:

  python tools/prof_step.py decode  [--steps 2] [--layers 28] [--batch 1] [--ctx 1088]
  python tools/prof_step.py vit     [--crops 8] [--layers 4]
  python tools/prof_step.py prefill [--tokens 1088] [--layers 4]

Random-init weights of the full-width architecture; --layers trims the depth so a profiler pass stays short (every
layer launches the same kernels on the same shapes). Prints the kernel launches it issued.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from omchat_b200 import lib  # noqa: E402
from omchat_b200.config import InternVisionConfig, OmChatQwen2Config  # noqa: E402
from omchat_b200.model.decoder import Qwen2Decoder  # noqa: E402
from omchat_b200.model.vision import InternVITVisionTower, MMProjector  # noqa: E402
from omchat_b200.model.weights import random_init  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["decode", "vit", "prefill"])
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--layers", type=int, default=0)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--ctx", type=int, default=1088)
    ap.add_argument("--crops", type=int, default=8)
    ap.add_argument("--tokens", type=int, default=1088)
    a = ap.parse_args()
    dev = "cuda:0"
    torch.cuda.set_device(0)
    lib.load()
    if a.what == "vit":
        vc = InternVisionConfig(num_hidden_layers=a.layers or 4)
        cfg = OmChatQwen2Config(vision_config=vc)
        w = random_init(cfg, device=dev, text=False)
        tower, proj = InternVITVisionTower(cfg, w.vit), MMProjector(w.proj)
        px = torch.randn(a.crops, 3, 448, 448, device=dev)
        for _ in range(a.steps):
            n0 = lib.launch_count()
            proj(tower(px))
            torch.cuda.synchronize()
        print("vit launches per pass:", lib.launch_count() - n0)
        return
    cfg = OmChatQwen2Config(num_hidden_layers=a.layers or (28 if a.what == "decode" else 4))
    w = random_init(cfg, device=dev, vision=False)
    dec = Qwen2Decoder(cfg, w.llm)
    if a.what == "prefill":
        T = a.tokens
        emb = torch.randn(T, cfg.hidden_size, device=dev, dtype=torch.bfloat16) * 0.02
        pos = torch.arange(T, device=dev, dtype=torch.int32)
        seq = torch.zeros(T, device=dev, dtype=torch.int32)
        cache = dec.new_cache(1, T + 64)
        for _ in range(a.steps):
            n0 = lib.launch_count()
            dec.prefill(emb.clone(), pos, seq, [0, T], cache, logits="last")
            torch.cuda.synchronize()
        print("prefill launches per pass:", lib.launch_count() - n0)
        return
    B = a.batch
    cache = dec.new_cache(B, a.ctx + a.steps + 8)
    cache.host_lens = [a.ctx] * B
    cache.ctx_lens.fill_(a.ctx)
    cache.pool.normal_(0, 0.5)
    toks = torch.randint(0, cfg.vocab_size, (B,), device=dev)
    for _ in range(a.steps):
        n0 = lib.launch_count()
        dec.decode_step(toks, cache, sample=True)
        torch.cuda.synchronize()
    print("decode launches per step:", lib.launch_count() - n0)


if __name__ == "__main__":
    main()
