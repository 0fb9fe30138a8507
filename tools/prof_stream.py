#!/usr/bin/env python
"""In-situ timeline of the batched decode step's weight-streaming GEMMs (csrc/gemm_stream.cu): every launch stamps
%globaltimer per CTA at 8 points; this prints, per launch of one steady-state step, the times (us, relative to the first
stamp of the step) of: first / last CTA entry, dependency satisfied (first..last), first stage landed, last MMA issued,
accumulators read, last stores issued - i.e. where the step's wall time goes with programmatic dependent launch active.
   python tools/prof_stream.py --batch 32 --ctx 1024 --layers 4"""
import argparse
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
    ap.add_argument("--layers", type=int, default=4)
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    L = lib.load()
    cfg = OmChatQwen2Config(num_hidden_layers=a.layers)
    if world > 1:  # tensor-parallel step (torchrun): the timeline of rank 0 includes the waits for the peers' partial tiles
        import torch.distributed as dist
        from omchat_b200.model.decoder import TPInfo
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        w = random_init(cfg, device=f"cuda:{local}", vision=False, tp_rank=rank, tp_size=world)
        dec = Qwen2Decoder(cfg, w.llm, TPInfo(rank=rank, size=world, group=dist.group.WORLD))
    else:
        w = random_init(cfg, device="cuda:0", vision=False)
        dec = Qwen2Decoder(cfg, w.llm)
    B = a.batch
    cache = dec.new_cache(B, a.ctx + 64)
    cache.host_lens = [a.ctx] * B
    cache.ctx_lens.fill_(a.ctx)
    cache.pool.normal_(0, 0.5)
    toks = torch.randint(0, cfg.vocab_size, (B,), device=f"cuda:{local}")
    for _ in range(3):
        dec.decode_step(toks, cache, sample=True)
    torch.cuda.synchronize()
    per_step = 4 * a.layers + 1
    grid = 2 * lib.num_sms()
    buf = torch.zeros(2 * per_step, grid, 16, device=f"cuda:{local}", dtype=torch.int64)
    L.omc_gemm_stream_set_prof(buf.data_ptr(), 2 * per_step)
    for _ in range(2):
        dec.decode_step(toks, cache, sample=True)
    torch.cuda.synchronize()
    L.omc_gemm_stream_set_prof(None, 0)
    if rank != 0:
        torch.distributed.barrier()
        return
    t = buf[per_step:].cpu().double()  # second profiled step
    names = (["qkv", "o", "gate_up", "down"] * a.layers) + ["lm_head"]
    t0 = t[t > 0].min()
    print(f"{'launch':10s} {'ctas':>4s} {'entry':>15s} {'w issued':>9s} {'dep ok':>15s} {'1st stage':>15s} {'last mma':>15s} "
          f"{'acc read':>15s} {'gathered':>9s} {'pushed':>15s} {'peers in':>15s} {'done':>15s}   (us since the step's first stamp; first..last CTA)")
    for i, name in enumerate(names):
        x = t[i]
        live = x[:, 0] > 0
        n = int(live.sum())
        x = x[live]

        def rng(c):
            v = x[:, c]
            v = v[v > 0]
            return f"{(v.min() - t0) / 1e3:7.1f}..{(v.max() - t0) / 1e3:6.1f}" if v.numel() else "      -"

        def mx(c):
            v = x[:, c]
            v = v[v > 0]
            return f"{(v.max() - t0) / 1e3:9.1f}" if v.numel() else "        -"

        print(f"{name:10s} {n:4d} {rng(0)} {mx(1)} {rng(2)} {rng(3)} {rng(4)} {rng(5)} {mx(6)} {rng(8)} {rng(9)} {rng(7)}")
    if world > 1:
        torch.distributed.barrier()


if __name__ == "__main__":
    main()
