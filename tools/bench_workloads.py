"""The multi-GPU workloads of BASELINE.json besides the headline request (bench.py --workload c3 | c4 | c5).

c3 (configs[2]): InternViT-6B tower + mm_projector only, 64 synthetic 448x448 crops, DATA-parallel: rank r encodes crops
    r, r+N, r+2N, ... with replicated weights and NO collective on the data path (SURVEY.md §8e "independent crops").
    value = images/s over all ranks (strong scaling: the 64 crops are fixed). roofline: tensor (11.945 TF/crop + projector).
c4 (configs[3]): Qwen2-7B decoder TENSOR-parallel over N ranks: 32 sequences x 1024-token multimodal prefill (768 text
    ids + one placeholder -> 256 image tokens, i.e. pixel-shuffle 0.5 features, synthetic) and batch-32 greedy decode
    over the paged KV cache. value = decode tokens/s (all 32 sequences); prefill tokens/s is reported beside it.

c5 (configs[4]): OmChat-2.1-8B-style multi-image requests: 16 prompts x 8 images x 256 image tokens (pixel-shuffle 0.5)
    + 2048 text ids = 4096-token contexts, greedy decode. DATA-parallel replicas: rank r serves prompts r, r+N, ... in
    mini-batches of 2 (the per-GPU share at N = 8) through the public generate() call; no collective on the data path.
    value = generated tokens/s over all ranks; crops/s of the vision phase is reported beside it.

All print ONE JSON line on rank 0 with the same keys as bench.py's main line.
"""
from __future__ import annotations

import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _peaks():
    from bench import load_peaks
    return load_peaks()


def _finish(torch, dist, dec=None):
    import bench
    if dec is not None:
        bench._TEARDOWN.append(lambda: bench._release_decoder(dec))
    bench._finish(torch, dist)


def _barrier(torch, dist, world):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def _max_over_ranks(torch, dist, world, vals, dev):
    t = torch.tensor(vals, device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


# ------------------------------------------------------------------------------------------------------------ c3
def run_c3(args, rank, world, local, total_crops: int = 64):
    import torch
    import torch.distributed as dist
    from omchat_b200.config import OmChatQwen2Config
    from omchat_b200.model.vision import InternVITVisionTower, MMProjector
    from omchat_b200.model.weights import random_init

    dev = torch.device("cuda", local)
    cfg = OmChatQwen2Config(mm_pixel_shuffle_ratio=args.pixel_shuffle)
    w = random_init(cfg, device=dev, seed=0, vision=True, text=False)
    line = measure_c3(args, rank, world, local, cfg, InternVITVisionTower(cfg, w.vit), MMProjector(w.proj), total_crops)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        _finish(torch, dist)


def measure_c3(args, rank, world, local, cfg, tower, proj, total_crops: int = 64) -> dict:
    """Times config c3 on an existing tower + projector (also called by bench.py's main line: `workloads.c3`)."""
    import torch
    import torch.distributed as dist
    from bench import ClockSampler
    from omchat_b200 import lib

    dev = torch.device("cuda", local)
    mine = list(range(rank, total_crops, world))  # crop i -> GPU i mod N
    g = torch.Generator().manual_seed(1)
    pixels_all = torch.randn(total_crops, 3, 448, 448, generator=g)
    pixels_host = pixels_all[mine].contiguous().pin_memory()
    pixels_dev = pixels_host.to(dev)
    down = cfg.pixel_shuffle_down

    def step(px):
        return proj(tower(px, down))

    for _ in range(max(args.warmup, 3)):
        step(pixels_dev)
    _barrier(torch, dist, world)
    n0 = lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        _barrier(torch, dist, world)
        e0.record()
        for _ in range(args.steps):
            feats = step(pixels_dev)
        e1.record()
        _barrier(torch, dist, world)
    launches = lib.launch_count() - n0
    (ms,) = _max_over_ranks(torch, dist, world, [e0.elapsed_time(e1) / args.steps], dev)
    value = total_crops / (ms * 1e-3)
    # e2e: pinned host pixels -> device -> features -> host
    _barrier(torch, dist, world)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out_host = step(pixels_host.to(dev, non_blocking=True)).cpu()
    _barrier(torch, dist, world)
    (e2e_s,) = _max_over_ranks(torch, dist, world, [(time.perf_counter() - t0) / args.steps], dev)
    peaks = _peaks()
    vc = cfg.vision_config
    S, C, I = vc.num_patches + 1, vc.hidden_size, vc.intermediate_size
    flop_crop = vc.num_hidden_layers * (2.0 * S * (4 * C * C + 2 * C * I) + 4.0 * S * S * C) + 2.0 * vc.num_patches * 588 * C
    L = cfg.image_tokens_per_crop
    flop_crop += 2.0 * L * (C * down * down * cfg.hidden_size + cfg.hidden_size ** 2)
    tf = flop_crop * len(mine) / (ms * 1e-3) / 1e12  # per GPU (the slowest rank's time, this rank's crops)
    line = {
        "metric": "images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"c3: InternViT-6B tower + mm_projector, {total_crops} synthetic 448x448 crops, data-parallel",
                   "crops_per_gpu": len(mine), "parallelism": f"dp{world} (no collective)",
                   "pixel_shuffle_ratio": args.pixel_shuffle,
                   "l2": "no flush needed: each step streams 11 GB of weights and > 1 GB of activations per layer"},
        "e2e": {"value": total_crops / e2e_s, "unit": "images/s",
                "h2d_bytes_per_step": pixels_host.numel() * 4, "d2h_bytes_per_step": out_host.numel() * 2},
        "gpu_launches": launches, "clocks": clocks.summary(),
        "roofline": {"bound": "tensor", "kernel": "whole tower step (tcgen05 GEMMs + attention)", "achieved": tf,
                     "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": tf / peaks["bf16_tflops_sustained"],
                     "traffic": None, "peak_source": peaks["source"] + " (bf16_tflops_sustained)",
                     "flops_per_crop": flop_crop},
    }
    return line


# ------------------------------------------------------------------------------------------------------------ c4
def run_c4(args, rank, world, local, n_seq: int = 32, text_tokens: int = 768, image_tokens: int = 256):
    import torch
    import torch.distributed as dist
    from bench import ClockSampler
    from omchat_b200 import lib
    from omchat_b200.config import IMAGE_TOKEN_INDEX, OmChatQwen2Config
    from omchat_b200.model.decoder import Qwen2Decoder, TPInfo
    from omchat_b200.model.weights import random_init

    dev = torch.device("cuda", local)
    cfg = OmChatQwen2Config(mm_pixel_shuffle_ratio=0.5, eos_token_id=-1)
    w = random_init(cfg, device=dev, seed=0, vision=False, text=True, tp_rank=rank, tp_size=world)
    dec = Qwen2Decoder(cfg, w.llm, TPInfo(rank=rank, size=world, group=dist.group.WORLD if world > 1 else None))
    line = measure_c4(args, rank, world, local, cfg, dec, n_seq, text_tokens, image_tokens)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        _finish(torch, dist, dec)


def measure_c4(args, rank, world, local, cfg, dec, n_seq: int = 32, text_tokens: int = 768, image_tokens: int = 256) -> dict:
    """Times config c4 on an existing (tensor-parallel) decoder (also called by bench.py's main line: `workloads.c4`)."""
    import torch
    import torch.distributed as dist
    from bench import ClockSampler
    from omchat_b200 import lib
    from omchat_b200.config import IMAGE_TOKEN_INDEX

    dev = torch.device("cuda", local)
    w_llm = dec.w
    new_tokens = min(args.new_tokens, 256)
    T = text_tokens + image_tokens
    g = torch.Generator().manual_seed(2)
    ids = torch.randint(0, 151643, (n_seq, text_tokens + 1), generator=g)
    ids[:, 16] = IMAGE_TOKEN_INDEX
    feats = (torch.randn(n_seq, image_tokens, cfg.hidden_size, generator=g) * 0.02).to(torch.bfloat16)
    ids_host, feats_host = ids.pin_memory(), feats.pin_memory()
    seq_off = (torch.arange(n_seq + 1, dtype=torch.int32) * (text_tokens + 1)).to(dev)
    offsets = [i * T for i in range(n_seq + 1)]
    cache = dec.new_cache(n_seq, T + new_tokens)
    first = torch.empty(n_seq, device=dev, dtype=torch.int64)
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    def request(ids_d, feats_d, timing=None):
        e = [ev() for _ in range(3)] if timing is not None else None
        if e:
            e[0].record()
        embeds, pos, seq, _ = lib.splice(ids_d.reshape(-1), seq_off, w_llm.embed, feats_d, IMAGE_TOKEN_INDEX, 0, n_seq * T)
        logits = dec.prefill(embeds, pos, seq, offsets, cache, logits="last")
        st = dec._decode_state(n_seq, cache.capacity)
        st.logits.copy_(logits)
        dec._greedy(st)
        first.copy_(st.tokens)
        if e:
            e[1].record()
        toks = dec.generate_greedy(first, cache, new_tokens - 1)
        if e:
            e[2].record()
            timing.append(e)
        return toks

    ids_dev, feats_dev = ids_host.to(dev), feats_host.to(dev)
    for _ in range(max(args.warmup, 3)):
        request(ids_dev, feats_dev)
    _barrier(torch, dist, world)
    n0 = lib.launch_count()
    timings = []
    t0e, t1e = ev(), ev()
    with ClockSampler(local) as clocks:
        _barrier(torch, dist, world)
        t0e.record()
        for _ in range(args.steps):
            toks = request(ids_dev, feats_dev, timings)
        t1e.record()
        _barrier(torch, dist, world)
    launches = lib.launch_count() - n0
    tot = t0e.elapsed_time(t1e) / args.steps
    pf = sum(e[0].elapsed_time(e[1]) for e in timings) / args.steps
    dc = sum(e[1].elapsed_time(e[2]) for e in timings) / args.steps
    tot, pf, dc = _max_over_ranks(torch, dist, world, [tot, pf, dc], dev)
    _barrier(torch, dist, world)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out_host = request(ids_host.to(dev, non_blocking=True), feats_host.to(dev, non_blocking=True)).cpu()
    _barrier(torch, dist, world)
    (e2e_s,) = _max_over_ranks(torch, dist, world, [(time.perf_counter() - t0) / args.steps], dev)
    peaks = _peaks()
    gen = n_seq * new_tokens
    # decode bytes per step per GPU: local weights + local KV at the mean context
    wbytes = sum(l.qkv_w.numel() + l.o_w.numel() + l.gate_up_w.numel() + l.down_w.numel() for l in w_llm.layers) * 2 \
        + w_llm.lm_head.numel() * 2
    kv_tok = 2 * len(w_llm.layers) * dec.Hkv * 128 * 2
    step_bytes = wbytes + n_seq * (T + new_tokens / 2.0) * kv_tok
    us = 1000.0 * dc / max(new_tokens - 1, 1)
    gbs = step_bytes / (us * 1e-6) / 1e9
    p_mm = sum(l.qkv_w.numel() + l.o_w.numel() + l.gate_up_w.numel() + l.down_w.numel() for l in w_llm.layers)
    pf_flop = 2.0 * p_mm * n_seq * T + 2.0 * len(w_llm.layers) * dec.Hq * 128 * T * T * n_seq
    line = {
        "metric": "tokens/sec", "value": gen / (tot * 1e-3), "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": tot, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"c4: Qwen2-7B decoder, {n_seq} x {T}-token multimodal prefill ({image_tokens} image tokens) "
                               f"+ batch-{n_seq} greedy decode of {new_tokens} tokens, paged KV",
                   "parallelism": f"tp{world}" + (" (NCCL all-reduce x56/forward)" if world > 1 else ""),
                   "kv_cache": f"paged, page {cfg.kv_page_size}, shuffled block table",
                   "l2": "no flush needed: every step streams the rank's weights (>> 126 MB L2)"},
        "phases": {"prefill_ms": pf, "prefill_tokens_per_sec": n_seq * T / (pf * 1e-3),
                   "prefill_tflops_per_gpu": pf_flop / (pf * 1e-3) / 1e12,
                   "decode_ms": dc, "decode_ms_per_step": dc / max(new_tokens - 1, 1),
                   "decode_tokens_per_sec": n_seq * (new_tokens - 1) / (dc * 1e-3)},
        "e2e": {"value": gen / e2e_s, "unit": "tokens/s",
                "h2d_bytes_per_step": ids_host.numel() * 8 + feats_host.numel() * 2, "d2h_bytes_per_step": out_host.numel() * 8},
        "gpu_launches": launches, "clocks": clocks.summary(),
        "roofline": {"bound": "hbm", "kernel": "one batch-32 decode step (weight-streaming GEMMs + paged attention)",
                     "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                     "traffic": None, "peak_source": peaks["source"], "bytes_per_step_per_gpu": step_bytes,
                     "avg_step_us": us},
    }
    return line


# ------------------------------------------------------------------------------------------------------------ c5
def run_c5(args, rank, world, local, **kw):
    import torch
    import torch.distributed as dist
    line, model = measure_c5(args, rank, world, local, keep_model=True, **kw)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        _finish(torch, dist, model.model.decoder if model is not None else None)


def measure_c5(args, rank, world, local, n_prompts: int = 16, images_per_prompt: int = 8, text_tokens: int = 2048, mb: int = 2,
               keep_model: bool = False):
    """-> the c5 JSON line as a dict (and the model when keep_model). Builds its own model (pixel-shuffle 0.5 projector). In the
    data-parallel mode everything up to and including the warm-up runs without a collective, so a rank that fails there (out of
    memory next to another workload's buffers) is caught, agreed on by ONE all-reduce, and every rank returns {"error": ...}
    instead of leaving its peers in a barrier."""
    import torch
    import torch.distributed as dist
    from bench import ClockSampler
    from omchat_b200 import lib
    from omchat_b200.config import IMAGE_TOKEN_INDEX, OmChatQwen2Config
    from omchat_b200.model.omchat import OmChatQwen2ForCausalLM

    dev = torch.device("cuda", local)
    cfg = OmChatQwen2Config(mm_pixel_shuffle_ratio=0.5, eos_token_id=-1)
    tp_mode = getattr(args, "c5_mode", "dp") == "tp" and world > 1
    err, model = None, None
    try:
        if tp_mode:
            # SURVEY.md §8e row 3, second variant: ONE model over all GPUs - the 128 crops data-parallel over the N towers, one
            # all-gather of the projected features, then the decoder tensor-parallel N ways over all 16 prompts
            model = OmChatQwen2ForCausalLM(cfg, device=f"cuda:{local}", seed=0, tp_rank=rank, tp_size=world,
                                           tp_group=dist.group.WORLD)
            mb = getattr(args, "c5_mb", 0) or 8
        else:
            model = OmChatQwen2ForCausalLM(cfg, device=f"cuda:{local}", seed=0)
        new_tokens = min(args.new_tokens, 64)
        L = cfg.image_tokens_per_crop  # 256
        T = text_tokens + images_per_prompt * L
        mine = list(range(n_prompts)) if tp_mode else list(range(rank, n_prompts, world))
        g = torch.Generator().manual_seed(2)
        ids = torch.randint(0, 151643, (n_prompts, text_tokens + images_per_prompt), generator=g)
        step = (text_tokens + images_per_prompt) // images_per_prompt
        for j in range(images_per_prompt):
            ids[:, 8 + j * step] = IMAGE_TOKEN_INDEX  # placeholders evenly spaced
        g1 = torch.Generator().manual_seed(1)
        pixels = torch.randn(len(mine) * images_per_prompt, 3, 448, 448, generator=g1).to(torch.bfloat16)
        ids_host, px_host = ids[mine].contiguous().pin_memory(), pixels.pin_memory()
        groups = [list(range(i, min(i + mb, len(mine)))) for i in range(0, len(mine), mb)]
        ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

        def serve(ids_src, px_src):
            outs = []
            for grp in groups:
                i_d = ids_src[grp[0]:grp[-1] + 1].to(dev, non_blocking=True)
                p_d = px_src[grp[0] * images_per_prompt:(grp[-1] + 1) * images_per_prompt].to(dev, non_blocking=True)
                outs.append(model.generate(i_d, images=p_d, max_new_tokens=new_tokens, do_sample=False))
            return outs

        ids_dev, px_dev = ids_host.to(dev), px_host.to(dev)
        for _ in range(max(args.warmup, 3) if len(groups) <= 2 else 1):
            serve(ids_dev, px_dev)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        if tp_mode:
            raise
        err = f"{type(e).__name__}: {e}"[:300]
    if world > 1 and not tp_mode:
        flag = torch.tensor([0 if err else 1], device=dev, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0 and err is None:
            err = "another rank failed to set the workload up"
    if err is not None:
        if model is not None:
            model.close()
        line = {"error": err}
        return (line, None) if keep_model else line
    _barrier(torch, dist, world)
    n0 = lib.launch_count()
    t0e, t1e, v0, v1 = ev(), ev(), ev(), ev()
    with ClockSampler(local) as clocks:
        _barrier(torch, dist, world)
        t0e.record()
        for _ in range(args.steps):
            serve(ids_dev, px_dev)
        t1e.record()
        _barrier(torch, dist, world)
    launches = lib.launch_count() - n0
    v0.record()
    for grp in groups:
        model.encode_images(px_dev[grp[0] * images_per_prompt:(grp[-1] + 1) * images_per_prompt])
    v1.record()
    torch.cuda.synchronize()
    tot, vis = _max_over_ranks(torch, dist, world, [t0e.elapsed_time(t1e) / args.steps, v0.elapsed_time(v1)], dev)
    _barrier(torch, dist, world)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out_host = [o.cpu() for o in serve(ids_host, px_host)]
    _barrier(torch, dist, world)
    (e2e_s,) = _max_over_ranks(torch, dist, world, [(time.perf_counter() - t0) / args.steps], dev)
    gen = n_prompts * new_tokens
    crops = n_prompts * images_per_prompt
    line = {
        "metric": "tokens/sec", "value": gen / (tot * 1e-3), "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3) if len(groups) <= 2 else 1, "ms_per_step": tot, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"c5: {n_prompts} prompts x {images_per_prompt} images x {L} image tokens (pixel-shuffle 0.5) + "
                               f"{text_tokens} text ids = {T}-token contexts, {new_tokens} greedy tokens each",
                   "parallelism": (f"vision dp{world} -> feature all-gather -> decoder tp{world}, mini-batches of {mb} prompts"
                                   if tp_mode else f"dp{world} replicas, mini-batches of {mb} prompts per GPU (no collective)"),
                   "kv_cache": f"paged, page {cfg.kv_page_size}",
                   "l2": "no flush needed: every mini-batch streams 26 GB of weights"},
        "phases": {"vision_ms_per_rank": vis, "crops_per_sec_vision": crops / (vis * 1e-3),
                   "prompt_tokens_per_sec": n_prompts * T / (tot * 1e-3)},
        "e2e": {"value": gen / e2e_s, "unit": "tokens/s",
                "h2d_bytes_per_step": ids_host.numel() * 8 + px_host.numel() * 2,
                "d2h_bytes_per_step": sum(o.numel() for o in out_host) * 8},
        "gpu_launches": launches, "clocks": clocks.summary(),
    }
    if keep_model:
        return line, model
    model.close()
    del model
    torch.cuda.empty_cache()
    return line
