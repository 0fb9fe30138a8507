#!/usr/bin/env python
"""Batched decode step in isolation: full-depth Qwen2-7B decoder, B sequences at context `ctx`, the CUDA-graph replay path
generate() uses, CUDA events over `steps` steps. Prints ms/step and the fraction of the HBM roofline (weights + KV bytes).
Variants through the environment: OMCHAT_B200_NO_STREAM=1 (round-1 skinny path), OMCHAT_B200_PDL=0,
OMCHAT_B200_STREAM_CTAS_PER_SM=2.   python tools/bench_decode_batch.py --batch 32 --ctx 1024 --steps 64"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from omchat_b200 import lib  # noqa: E402
from omchat_b200.config import OmChatQwen2Config  # noqa: E402
from omchat_b200.model.decoder import Qwen2Decoder  # noqa: E402
from omchat_b200.model.weights import random_init  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--ctx", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--layers", type=int, default=28)
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    lib.load()
    cfg = OmChatQwen2Config(num_hidden_layers=a.layers)
    w = random_init(cfg, device="cuda:0", vision=False)
    dec = Qwen2Decoder(cfg, w.llm)
    B = a.batch
    cache = dec.new_cache(B, a.ctx + 2 * a.steps + 16)
    cache.host_lens = [a.ctx] * B
    cache.ctx_lens.fill_(a.ctx)
    cache.pool.normal_(0, 0.5)
    toks = torch.randint(0, cfg.vocab_size, (B,), device="cuda:0")
    dec.generate_greedy(toks, cache, 8)  # warm-up + graph capture
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dec.generate_greedy(toks, cache, a.steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    wbytes = sum(l.qkv_w.numel() + l.o_w.numel() + l.gate_up_w.numel() + l.down_w.numel() for l in w.llm.layers) * 2 \
        + w.llm.lm_head.numel() * 2
    kv = B * (a.ctx + 8 + a.steps / 2.0) * 2 * a.layers * dec.Hkv * 128 * 2
    gbs = (wbytes + kv) / (ms * 1e-3) / 1e9
    peak = 6535.4
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    print(json.dumps({"tag": a.tag, "batch": B, "ctx": a.ctx, "layers": a.layers, "ms_per_step": ms, "gbs": gbs,
                      "frac_hbm": gbs / peak, "stream": dec.use_stream(B), "pdl": lib.PDL_ENABLED,
                      "ctas_per_sm": os.environ.get("OMCHAT_B200_STREAM_CTAS_PER_SM", "1")}))


if __name__ == "__main__":
    main()
