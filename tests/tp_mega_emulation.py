"""Single-GPU emulation of the tensor-parallel persistent decode kernel (run as a script by test_decode_mega_gpu.py so
that a device trap cannot poison the pytest process).

Two (or four) "ranks" live in ONE process on ONE GPU: each is a Qwen2Decoder holding its Megatron shard of the same
seed-0 full-width weights, each launches its own cooperative decode_mega_kernel on its own stream with grid = SMs / tp
CTAs so that all of them are co-resident, and their peer exchange buffers are plain device buffers of this process. The
cross-"GPU" protocol (8-byte {value, tag} pushes of the row-parallel partial sums, polling, the (max, index) exchange of
the vocab-parallel argmax) is exactly the one that runs over NVLink between processes.

Checks: every rank samples the same token each step; tokens equal the single-GPU persistent kernel's wherever its top-1
margin exceeds bf16 noise; concatenated vocab-shard logits match its logits (cosine >= 0.999, max-abs <= 2 % of scale).
"""
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from omchat_b200 import lib  # noqa: E402
from omchat_b200.config import OmChatQwen2Config  # noqa: E402
from omchat_b200.model.decoder import Qwen2Decoder, TPInfo  # noqa: E402
from omchat_b200.model.weights import random_init, tp_plan  # noqa: E402


def main(tp: int, lens, steps: int = 6, layers: int = 2):
    dev = "cuda"
    if tp <= 2:
        cfg = OmChatQwen2Config(num_hidden_layers=layers)  # full Qwen2-7B width
    else:  # SMs / tp CTAs must still own <= 63 residual rows each: a narrower model for tp = 4
        cfg = OmChatQwen2Config(num_hidden_layers=layers, hidden_size=2048, num_attention_heads=16, num_key_value_heads=4,
                                intermediate_size=8192, vocab_size=32000)
    B = len(lens)
    full = Qwen2Decoder(cfg, random_init(cfg, device=dev, seed=0, vision=False).llm)
    full.stream_min_b = 5  # the reference here is the single-GPU PERSISTENT kernel (batches 2..4 default to the stream GEMMs)
    ranks = [Qwen2Decoder(cfg, random_init(cfg, device=dev, seed=0, vision=False, tp_rank=r, tp_size=tp).llm,
                          TPInfo(rank=r, size=tp)) for r in range(tp)]
    # prefill once on the unsharded decoder; every rank's cache = its kv heads of that cache (same block table: same seed)
    g = torch.Generator(device=dev).manual_seed(1)
    T = sum(lens)
    emb = (torch.randn(T, cfg.hidden_size, generator=g, device=dev) * 0.02).to(torch.bfloat16)
    pos = torch.cat([torch.arange(n, dtype=torch.int32) for n in lens]).to(dev)
    seq = torch.cat([torch.full((n,), i, dtype=torch.int32) for i, n in enumerate(lens)]).to(dev)
    offs = [0]
    for n in lens:
        offs.append(offs[-1] + n)
    cache = full.new_cache(B, max(lens) + 40)
    first = full.prefill(emb, pos, seq, offs, cache, logits="last").argmax(-1)
    caches = []
    for r, d in enumerate(ranks):
        c = d.new_cache(B, max(lens) + 40)
        assert torch.equal(c.block_table, cache.block_table)
        kv = tp_plan(cfg, r, tp).kv_heads
        c.pool.copy_(cache.pool[:, :, :, kv])
        c.ctx_lens.copy_(cache.ctx_lens)
        c.host_lens = list(cache.host_lens)
        caches.append(c)
    # the "peer" exchange buffers and one plan per rank, grid = SMs / tp
    nbytes = lib.decode_xchg_bytes(B, cfg.hidden_size, tp)
    xbufs = [torch.zeros(nbytes, device=dev, dtype=torch.uint8) for _ in range(tp)]
    ptrs = [x.data_ptr() for x in xbufs]
    grid = lib.num_sms() // tp
    states, plans, streams = [], [], []
    for d, c in zip(ranks, caches):
        st = d._decode_state(B, c.capacity)
        st.tokens.copy_(first)
        states.append(st)
        plans.append(d._mega_plan(st, c, xchg_ptrs=ptrs, grid=grid))
        streams.append(torch.cuda.Stream())
    torch.cuda.synchronize()
    # reference: the single-GPU persistent kernel
    ref_tok, ref_logits = [], []
    cur = first.clone()
    for _ in range(steps):
        lg = full.decode_step(cur, cache).clone()
        cur = full._decode_state(B, cache.capacity).tokens.clone()
        ref_tok.append(cur)
        ref_logits.append(lg)
    torch.cuda.synchronize()
    for i in range(steps):
        for r in range(tp):
            with torch.cuda.stream(streams[r]):
                plans[r].step(i + 1)
        torch.cuda.synchronize()
        toks = [st.tokens.clone() for st in states]
        for r in range(1, tp):
            assert torch.equal(toks[0], toks[r]), f"step {i}: ranks disagree {toks}"
        lg = torch.cat([st.logits for st in states], dim=1)
        assert torch.equal(lg.argmax(-1), toks[0]), "exchanged argmax must equal the argmax of the concatenated logits"
        cos = torch.nn.functional.cosine_similarity(lg, ref_logits[i], dim=-1).min().item()
        err = (lg - ref_logits[i]).abs().max().item() / ref_logits[i].abs().max().item()
        assert cos >= 0.999 and err <= 0.02, (i, cos, err)
        top2 = torch.topk(ref_logits[i], 2, dim=-1).values
        for b in range(B):
            if int(toks[0][b]) != int(ref_tok[i][b]):
                assert float(top2[b, 0] - top2[b, 1]) < 0.02 * float(ref_logits[i][b].abs().max()), (i, b)
        # keep the two sequences identical even if a near-tie flipped
        for st in states:
            st.tokens.copy_(ref_tok[i])
        print(f"tp{tp} step {i}: tokens {toks[0].tolist()} cos {cos:.6f} err {err:.4f}")
    for r, c in enumerate(caches):
        assert c.ctx_lens.tolist() == [n + steps for n in lens]
        kv = tp_plan(cfg, r, tp).kv_heads
        a, b = c.pool.float(), cache.pool[:, :, :, kv].float()
        assert (a - b).abs().max().item() <= 0.02 * b.abs().max().item()
    print(f"tp{tp} emulation ok")


if __name__ == "__main__":
    tp = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    lens = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [130, 77]
    main(tp, lens)
