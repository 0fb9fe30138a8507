#!/usr/bin/env python
"""Continuous batching (omchat_b200/serving.py) on the full-size model: N single-image requests (448x448 crop + 64-token
prompt, T = 1088; mixed generation lengths) through 1..4 decode slots; generated tokens/s end to end (host ids and pixels
in, token ids out) against the same requests served one by one with generate().
  python tools/bench_serving.py [--requests 12] [--slots 4]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from omchat_b200.config import OmChatQwen2Config  # noqa: E402
from omchat_b200.model.omchat import OmChatQwen2ForCausalLM  # noqa: E402
from omchat_b200.serving import ContinuousBatcher  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--requests", type=int, default=12)
    ap.add_argument("--slots", type=int, default=4)
    a = ap.parse_args()
    cfg = OmChatQwen2Config(eos_token_id=-1)
    model = OmChatQwen2ForCausalLM(cfg, device="cuda:0", seed=0)
    g = torch.Generator().manual_seed(2)
    reqs = []
    for i in range(a.requests):
        ids = torch.randint(0, 151643, (1, 65), generator=g)
        ids[0, 16] = -200
        px = torch.randn(1, 3, 448, 448, generator=g).to(torch.bfloat16)
        reqs.append((ids, px, 48 + 16 * (i % 4)))  # 48..96 new tokens
    total = sum(r[2] for r in reqs)

    def serve_batched(slots):
        cb = ContinuousBatcher(model, slots=slots, max_ctx=1088 + 128, chunk=16)
        for ids, px, mn in reqs:
            cb.submit(ids, px, max_new_tokens=mn)
        return cb.run(), cb

    def serve_sequential():
        return [model.generate(ids, images=px, max_new_tokens=mn, do_sample=False).cpu() for ids, px, mn in reqs]

    serve_batched(a.slots)  # warm-up (plans, caches)
    serve_sequential()
    out = {}
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    serve_sequential()
    torch.cuda.synchronize()
    out["sequential_generate_tokens_per_s"] = total / (time.perf_counter() - t0)
    for slots in sorted({1, 2, 4, 8, a.slots}):
        serve_batched(slots)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res, cb = serve_batched(slots)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        assert sum(len(v) for v in res.values()) == total
        out[f"continuous_batching_slots{slots}_tokens_per_s"] = total / dt
        out[f"decode_steps_slots{slots}"] = cb.steps_run
    out.update({"requests": a.requests, "generated_tokens": total, "prompt_tokens_each": 1088})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
