#!/usr/bin/env python
"""Qwen2-MoE decoder (Qwen1.5-MoE-A2.7B sizes: hidden 2048, 24 layers, 60 experts, top-4, expert width 1408, shared 5632) in
isolation: prefill of `--prefill` x 1024-token sequences (tokens/s, expert + dense FLOPs) and the decode step at batch B and
context ctx through the CUDA-graph path (ms/step; HBM roofline on the ALGORITHMIC bytes = attention / shared-expert / router /
lm_head weights + the distinct routed experts' weights a step touches (expected value under uniform routing) + KV).
python tools/bench_moe.py --batch 1 --batch 32"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from omchat_b200 import lib  # noqa: E402
from omchat_b200.config import OmChatQwen2MoeConfig  # noqa: E402
from omchat_b200.model.moe import Qwen2MoeDecoder  # noqa: E402
from omchat_b200.model.weights import random_init  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, action="append")
    ap.add_argument("--ctx", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--layers", type=int, default=24)
    ap.add_argument("--prefill", type=int, default=8)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    for rec in measure(a.layers, a.batch or [1, 32], a.ctx, a.steps, a.prefill):
        print(json.dumps(rec))


def measure(layers=24, batches=(1, 32), ctx=1024, steps=64, prefill=8):
    """-> list of records (one for the prefill, one per decode batch size); see the module docstring."""
    class a:  # noqa: N801
        pass
    a.layers, a.batch, a.ctx, a.steps, a.prefill = layers, list(batches), ctx, steps, prefill
    out = []
    lib.load()
    cfg = OmChatQwen2MoeConfig(num_hidden_layers=a.layers, mm_vision_tower=None)
    w = random_init(cfg, device="cuda:0", vision=False)
    dec = Qwen2MoeDecoder(cfg, w.llm)
    peak_hbm, peak_tf = 6535.4, 1400.0
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak_hbm, peak_tf = pk.get("hbm_gbs", peak_hbm), pk.get("bf16_tflops_sustained", pk.get("bf16_tflops", peak_tf))
    except Exception:
        pass
    C, E, k, Im, Is = cfg.hidden_size, cfg.num_experts, cfg.num_experts_per_tok, cfg.moe_intermediate_size, cfg.shared_expert_intermediate_size
    Hq, Hkv = cfg.num_attention_heads, cfg.num_key_value_heads
    attn_w = (Hq + 2 * Hkv) * 128 * C + C * Hq * 128
    expert_w = 3 * Im * C
    shared_w = 3 * Is * C + E * C + C
    ev = lambda: (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))  # noqa: E731
    # ---- prefill
    n, L = a.prefill, 1024
    T = n * L
    emb = (torch.randn(T, C, device="cuda") * 0.5).to(torch.bfloat16)
    pos = torch.arange(L, dtype=torch.int32, device="cuda").repeat(n)
    seq = torch.arange(n, dtype=torch.int32, device="cuda").repeat_interleave(L)
    offs = [i * L for i in range(n + 1)]
    cache = dec.new_cache(n, L + 16)
    for _ in range(2):
        dec.prefill(emb.clone(), pos, seq, offs, cache, logits="last")
    torch.cuda.synchronize()
    e0, e1 = ev()
    n0 = lib.launch_count()
    e0.record()
    for _ in range(3):
        dec.prefill(emb.clone(), pos, seq, offs, cache, logits="last")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    flops = 2.0 * T * a.layers * (attn_w + k * expert_w + 3 * Is * C) + 4.0 * n * a.layers * Hq * 128 * L * L / 2
    out.append({"phase": "prefill", "tokens": T, "ms": ms, "tokens_per_sec": T / ms * 1e3, "tflops": flops / ms / 1e9,
                "frac_tensor": flops / ms / 1e9 / peak_tf, "launches": (lib.launch_count() - n0) // 3})
    del cache, emb
    # ---- decode
    for B in a.batch or [1, 32]:
        cache = dec.new_cache(B, a.ctx + 2 * a.steps + 16)
        cache.host_lens = [a.ctx] * B
        cache.ctx_lens.fill_(a.ctx)
        cache.pool.normal_(0, 0.5)
        toks = torch.randint(0, cfg.vocab_size, (B,), device="cuda:0")
        dec.generate_greedy(toks, cache, 8)
        torch.cuda.synchronize()
        e0, e1 = ev()
        n0 = lib.launch_count()
        e0.record()
        dec.generate_greedy(toks, cache, a.steps)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        distinct = E * (1.0 - (1.0 - k / E) ** B)  # expected number of distinct experts B tokens touch (uniform routing)
        wbytes = 2.0 * (a.layers * (attn_w + shared_w + distinct * expert_w) + w.llm.lm_head.numel())
        kv = B * (a.ctx + 8 + a.steps / 2.0) * 2 * a.layers * Hkv * 128 * 2
        gbs = (wbytes + kv) / (ms * 1e-3) / 1e9
        out.append({"phase": "decode", "batch": B, "ctx": a.ctx, "ms_per_step": ms, "tokens_per_sec": B / ms * 1e3,
                    "algorithmic_gb": (wbytes + kv) / 1e9, "gbs": gbs, "frac_hbm": gbs / peak_hbm,
                    "launches_per_step": (lib.launch_count() - n0) // a.steps})
        del cache
    dec.release()
    return out


if __name__ == "__main__":
    main()
