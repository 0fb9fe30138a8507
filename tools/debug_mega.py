"""Debug driver for the persistent decode kernel: runs a few steps at increasing depth and prints the watchdog record."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from omchat_b200 import lib
from omchat_b200.config import OmChatQwen2Config
from omchat_b200.model.decoder import Qwen2Decoder
from omchat_b200.model.weights import random_init

def run(layers, B=1, ctx=200, steps=3, **kw):
    cfg = OmChatQwen2Config(num_hidden_layers=layers, **kw)
    w = random_init(cfg, device="cuda", vision=False)
    dec = Qwen2Decoder(cfg, w.llm)
    cache = dec.new_cache(B, ctx + 64)
    cache.pool.normal_(0, 0.5)
    cache.host_lens = [ctx] * B
    cache.ctx_lens.fill_(ctx)
    toks = torch.randint(0, cfg.vocab_size, (B,), device="cuda")
    st = dec._decode_state(B, cache.capacity)
    plan = dec._mega_plan(st, cache)
    try:
        for i in range(steps):
            lg = dec.decode_step(toks, cache).clone()
            torch.cuda.synchronize()
            toks = st.tokens.clone()
            print(f"layers={layers} B={B} step {i}: tokens {toks.tolist()} argmax(logits) {lg.argmax(-1).tolist()} ctx {cache.ctx_lens.tolist()}", flush=True)
        # compare with the per-op path on the same state
        dec.mega_enabled = False
        cache.ctx_lens.fill_(ctx); cache.host_lens = [ctx] * B
        t2 = torch.randint(0, cfg.vocab_size, (B,), device="cuda")
        a = dec.decode_step(t2, cache).clone()
        dec.mega_enabled = True
        cache.ctx_lens.fill_(ctx); cache.host_lens = [ctx] * B
        b = dec.decode_step(t2, cache).clone()
        torch.cuda.synchronize()
        cos = torch.nn.functional.cosine_similarity(a, b, dim=-1)
        print(f"  per-op vs mega logits: cos {cos.tolist()} maxerr {(a-b).abs().max().item():.4g} scale {a.abs().max().item():.4g}", flush=True)
    except Exception as e:
        print("FAILED:", type(e).__name__, str(e)[:200])
        print("watchdog {code, cta, detail, thread}:", plan.status.tolist(), flush=True)
        sys.exit(1)

if __name__ == "__main__":
    tiny = dict(vocab_size=1000, hidden_size=256, intermediate_size=512, num_attention_heads=2, num_key_value_heads=1, kv_page_size=16)
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "tiny0"): run(0, **tiny)
    if which in ("all", "tiny1"): run(1, **tiny)
    if which in ("all", "full0"): run(0)
    if which in ("all", "full1"): run(1)
    if which in ("all", "full2b3"): run(2, B=3)
