#!/usr/bin/env python
"""bench.py — the OmChat hot path on B200: one "step" = one c2 request (BASELINE.json configs[1]):
    1 synthetic 448x448 image + 64-token prompt (one -200 placeholder -> T = 1088) -> InternViT-6B -> mm_projector ->
    splice -> Qwen2-7B prefill -> 256 greedy tokens (255 decode steps over the paged KV cache), bf16, random-init weights.

  python bench.py --gpus N --steps K --warmup W            # this repository's sm_100a path
  python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU fp32 path (oracle port)

N > 1 (torchrun, one rank per GPU): the SAME request served by N GPUs — Qwen2 decoder tensor-parallel (column/row
parallel, 2 all-reduces per layer: NCCL at prefill, inside the persistent decode kernel over NVLink peer memory at decode),
the single crop's vision tower replicated ("scaling": "strong").
Other workloads: --workload c3 (vision tower + projector, 64 crops data-parallel) and --workload c4 (decoder TP, 1024-token
prefill + batch-32 decode).

Prints ONE JSON line (rank 0). `value` = generated tokens / s over the whole request with inputs resident in HBM;
`e2e` = the same through model.generate() with pinned HOST inputs (H2D of pixels+ids and D2H of the ids inside the
timed region); `roofline` = the dominant kernel (decode GEMV weight streaming, HBM-bound); `roofline_tensor` = the
ViT/prefill tcgen05 GEMM; `cpu_baseline` = the fp32 oracle on the host cores (bounded sample, N=1 only).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PROMPT_TEXT_TOKENS = 64
PLACEHOLDER_AT = 16
NEW_TOKENS = 256
METRIC = "tokens/sec"
WORKLOAD_C2 = ("c2: OmChat-2.0-13B arch (InternViT-6B 448px + Qwen2-7B) bf16, batch-1, 1 synthetic 448x448 image + "
               "64-token prompt (T=1088), prefill + 256 greedy tokens")


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one decode_mega_kernel launch from the committed ncu --set full
    capture (profiles/r02_mega_traffic.json, from profiles/r02_ncu_mega.txt); None if the capture is absent."""
    for name in ("r02_mega_traffic.json", "r01_mega_traffic.json"):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name)))["traffic_bytes_per_launch"]
        except Exception:
            continue
    return None


# --------------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    """Samples SM clock + throttle reasons through NVML every 100 ms while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


_TEARDOWN = []  # callables that release what still references the communicator (captured graphs, peer mappings)


def _finish(torch, dist):
    """End of a multi-rank run. destroy_process_group() hung in round 1 because CUDA graphs that had captured NCCL
    collectives (the batched decode step) were still alive when the communicator was torn down: release them first (and
    close the CUDA-IPC peer mappings of the tensor-parallel decode kernel), drain the device, then destroy the group. A
    watchdog turns a teardown that still blocks into a loud message on stderr + exit instead of a silent hang."""
    dist.barrier()
    torch.cuda.synchronize()
    for fn in _TEARDOWN:
        try:
            fn()
        except Exception as e:  # noqa: BLE001
            print(f"[bench] teardown step failed: {e!r}", file=sys.stderr, flush=True)
    _TEARDOWN.clear()
    import gc
    gc.collect()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()

    def _watchdog():
        print("[bench] destroy_process_group() did not return within 60 s; exiting", file=sys.stderr, flush=True)
        os._exit(0)

    t = threading.Timer(60.0, _watchdog)
    t.daemon = True
    t.start()
    dist.destroy_process_group()
    t.cancel()


def _release_decoder(dec):
    """Drop the captured decode graphs / plans and close the peer exchange buffers of a decoder."""
    dec.release()


# --------------------------------------------------------------------------------------------------- reference arm
def host_threads() -> int:
    """All host threads this process may use (cgroup / affinity aware). torch.distributed.run exports OMP_NUM_THREADS=1
    to its workers; the reference arm runs alone on rank 0, so it takes the whole host regardless."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:  # noqa: BLE001
        return max(1, os.cpu_count() or 1)


def run_reference(args):
    """The reference's own CPU fp32 implementation of the path, timed on ALL host cores (oracle port; see
    oracle/cpu_baseline.py for why it is a port). The full c2 request in fp32 is ~52 GB of weights and minutes of CPU time
    per step, so each step times a BOUNDED SAMPLE of it - a few full-width layers of each tower at the real sequence
    lengths + real decode steps + the full lm_head - and the full-depth request time is that sample scaled by the layer
    and token counts. Both are in the line: `ms_per_step` / `measured_seconds_per_step` are MEASURED wall time of the
    sample, `value` is the extrapolation and says so (`extrapolated: true`, `extrapolated_request_ms`). Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    from oracle.cpu_baseline import time_c2_sample
    cores = host_threads()
    torch.set_num_threads(cores)  # overrides the OMP_NUM_THREADS=1 that torchrun puts in the environment
    vals, walls = [], []
    for i in range(args.warmup + args.steps):
        quick = i < args.warmup
        t0 = time.perf_counter()
        r = time_c2_sample(vit_layers=1 if quick else 2, llm_layers=1 if quick else 2, decode_steps=2 if quick else 4,
                           threads=cores)
        if not quick:
            vals.append(r)
            walls.append(time.perf_counter() - t0)
    v = sum(x["tokens_per_sec_request"] for x in vals) / len(vals)
    req_ms = 1000.0 * sum(x["seconds"]["request"] for x in vals) / len(vals)
    wall = sum(walls) / len(walls)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "tokens/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * wall, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "extrapolated": True, "measured_seconds_per_step": wall, "extrapolated_request_ms": req_ms,
        "config": {"workload": WORKLOAD_C2, "device": "cpu", "threads": cores,
                   "note": "ms_per_step = measured wall time of one bounded sample (weight generation included); value = "
                           "256 tokens / extrapolated_request_ms, the sample's per-layer and per-token times scaled to 45 "
                           "ViT blocks, 28 decoder layers and 255 decode steps"},
        "cpu_baseline": {"value": v, "unit": "tokens/s", "cores": vals[-1]["cores"], "kind": "port",
                         "sample": vals[-1]["sample"], "extrapolated": True,
                         "decode_tokens_per_sec": vals[-1]["decode_tokens_per_sec"],
                         "images_per_sec_vit_prefill": vals[-1]["images_per_sec_vit_prefill"]},
        "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------- roofline probes
def probe_gemv_roofline(model, torch, lib, reps=3):
    """Decode weight streaming in isolation: every GEMV of one decode step (28 x {qkv, o, gate/up, down} + lm_head),
    back to back on the current stream, CUDA events around the whole train of launches. Algorithmic bytes = the bf16
    weights each launch must read (14.14 GB per step at TP=1)."""
    dec = model.model.decoder
    w = dec.w
    st = dec._decode_state(1, 2048)
    st.h.normal_(0, 0.02)
    st.attn.normal_(0, 0.02)
    launches = 4 * len(w.layers) + 1
    bytes_total = sum(l.qkv_w.numel() + l.o_w.numel() + l.gate_up_w.numel() + l.down_w.numel() for l in w.layers) * 2
    bytes_total += w.lm_head.numel() * 2

    def train():
        for l in w.layers:
            lib.gemv(st.h, l.qkv_w, out=st.qkv, norm_w=l.ln1, eps=dec.eps, bias=l.qkv_b)
            lib.gemv(st.attn, l.o_w, out=st.xn, res=st.h, epi=lib.EPI_RES)
            lib.gemv(st.h, l.gate_up_w, out=st.act, norm_w=l.ln2, eps=dec.eps, epi=lib.EPI_SWIGLU)
            lib.gemv(st.act, l.down_w, out=st.xn, res=st.h, epi=lib.EPI_RES)
        lib.gemv(st.h, w.lm_head, out=st.logits, norm_w=w.norm, eps=dec.eps, out_f32=True)

    train()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        train()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return {"launches": launches, "bytes_per_launch": bytes_total / launches, "avg_launch_us": 1000.0 * ms / launches,
            "gbs": bytes_total / (ms * 1e-3) / 1e9, "train_ms": ms}


def probe_gemm_roofline(model, torch, lib, reps=2):
    """The tcgen05 GEMM on the ViT block shapes for one crop batch of 8 (M = 8200): qkv, proj, fc1, fc2 per block,
    45 blocks, CUDA events around the train. Algorithmic FLOPs = 2*M*N*K per launch."""
    vt = model.get_vision_tower()
    if vt is None or vt.w is None:
        return None
    C, I = vt.vc.hidden_size, vt.vc.intermediate_size
    M = 8 * (vt.vc.num_patches + 1)
    dev = model.device
    x = torch.randn(M, C, device=dev, dtype=torch.bfloat16) * 0.5
    qkv = torch.empty(M, 3 * C, device=dev, dtype=torch.bfloat16)
    h = torch.zeros(M, C, device=dev, dtype=torch.bfloat16)
    act = torch.empty(M, I, device=dev, dtype=torch.bfloat16)
    layers = vt.w.layers
    flops = len(layers) * 2.0 * M * (3 * C * C + C * C + 2 * C * I)

    def train():
        for l in layers:
            lib.gemm(x, l.qkv, out=qkv)
            lib.gemm(x, l.proj_w, out=h, bias=l.proj_b, scale=l.ls1, res=h, epi=lib.EPI_RES)
            lib.gemm(x, l.fc1_w, out=act, bias=l.fc1_b, epi=lib.EPI_GELU)
            lib.gemm(act, l.fc2_w, out=h, bias=l.fc2_b, scale=l.ls2, res=h, epi=lib.EPI_RES)

    train()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        train()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    launches = 4 * len(layers)
    return {"launches": launches, "flops_per_launch": flops / launches, "avg_launch_us": 1000.0 * ms / launches,
            "tflops": flops / (ms * 1e-3) / 1e12, "train_ms": ms}


# --------------------------------------------------------------------------------------------------- main arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--new-tokens", type=int, default=NEW_TOKENS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--c5-mode", default="dp", choices=["dp", "tp"], help="c5: data-parallel replicas, or vision-DP -> "
                    "feature all-gather -> decoder-TP over all GPUs")
    ap.add_argument("--c5-mb", type=int, default=0, help="c5: prompts per mini-batch (0 = default of the mode)")
    ap.add_argument("--no-workloads", action="store_true", help="skip the c3 / c4 legs of the main line (`workloads`)")
    ap.add_argument("--no-c5", action="store_true", help="skip the c5 leg of the main line (N > 1 only)")
    ap.add_argument("--no-variants", action="store_true", help="skip the Qwen2-MoE / InternViT-300M legs (`variants`, N = 1 only)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--pixel-shuffle", type=float, default=1.0, help="mm_pixel_shuffle_ratio (1.0 = reference behaviour)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("for --gpus N > 1 launch with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL's environment (NCCL_DEBUG & co) is left exactly as the launcher set it: its log lines go where NCCL puts them
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from omchat_b200 import lib
    from omchat_b200.config import OmChatQwen2Config
    lib.load()

    if args.workload == "c3":
        from tools.bench_workloads import run_c3
        return run_c3(args, rank, world, local)
    if args.workload == "c4":
        from tools.bench_workloads import run_c4
        return run_c4(args, rank, world, local)
    if args.workload == "c5":
        from tools.bench_workloads import run_c5
        return run_c5(args, rank, world, local)

    from omchat_b200.model.omchat import OmChatQwen2ForCausalLM
    cfg = OmChatQwen2Config(mm_pixel_shuffle_ratio=args.pixel_shuffle, eos_token_id=-1)
    model = OmChatQwen2ForCausalLM(cfg, device=f"cuda:{local}", seed=0, tp_rank=rank, tp_size=world,
                                   tp_group=dist.group.WORLD if world > 1 else None)
    dec = model.model.decoder
    dev = model.device
    new_tokens = args.new_tokens
    # synthetic inputs: seeds 1 (pixels) and 2 (prompt)
    g1, g2 = torch.Generator().manual_seed(1), torch.Generator().manual_seed(2)
    pixels_host = torch.randn(1, 3, 448, 448, generator=g1).pin_memory()
    ids_host = torch.randint(0, 151643, (1, PROMPT_TEXT_TOKENS + 1), generator=g2)
    ids_host[0, PLACEHOLDER_AT] = -200
    ids_host = ids_host.pin_memory()
    pixels_dev, ids_dev = pixels_host.to(dev), ids_host.to(dev)
    L = cfg.image_tokens_per_crop
    T = PROMPT_TEXT_TOKENS + L

    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_request(timing=None):
        """Device-resident request; returns the generated ids. Phase events are recorded on the current stream."""
        e = [ev() for _ in range(3)] if timing is not None else None
        if e:
            e[0].record()
        embeds, pos, seq, offsets = model._splice_packed(ids_dev, None, pixels_dev)
        cache = one_request.cache
        logits = dec.prefill(embeds, pos, seq, offsets, cache, logits="last")
        first = one_request.first
        if world == 1:
            lib.argmax(logits, out=first)
        else:
            st = dec._decode_state(1, cache.capacity)
            st.logits.copy_(logits)
            dec._greedy(st)
            first.copy_(st.tokens)
        if e:
            e[1].record()
        toks = dec.generate_greedy(first, cache, new_tokens - 1, use_graph=not args.no_graph)
        if e:
            e[2].record()
            timing.append(e)
        return first, toks

    one_request.cache = dec.new_cache(1, T + new_tokens)  # one paged pool reused by every step (pages are overwritten)
    one_request.first = torch.empty(1, device=dev, dtype=torch.int64)

    for _ in range(max(args.warmup, 3)):
        one_request()
    barrier()
    n0 = lib.launch_count()
    timings = []
    with ClockSampler(local) as clocks:
        barrier()
        t_start, t_end = ev(), ev()
        t_start.record()
        for _ in range(args.steps):
            first, toks = one_request(timings)
        t_end.record()
        barrier()
    launches = lib.launch_count() - n0
    total_ms = t_start.elapsed_time(t_end)
    vp_ms = sum(e[0].elapsed_time(e[1]) for e in timings) / args.steps
    dec_ms = sum(e[1].elapsed_time(e[2]) for e in timings) / args.steps
    stats = torch.tensor([total_ms, vp_ms, dec_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    total_ms, vp_ms, dec_ms = stats.tolist()
    ms_per_step = total_ms / args.steps
    value = new_tokens / (ms_per_step * 1e-3)

    # ---- e2e: the public API with HOST inputs, host<->device copies inside the timed region
    model.generate(ids_host, images=pixels_host, max_new_tokens=new_tokens, eos_token_id=-1)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = model.generate(ids_host, images=pixels_host, max_new_tokens=new_tokens, eos_token_id=-1)
        out_host = out.cpu()
    barrier()
    e2e_s = torch.tensor([(time.perf_counter() - t0) / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_val = new_tokens / e2e_s.item()
    h2d = pixels_host.numel() * pixels_host.element_size() + ids_host.numel() * ids_host.element_size()
    d2h = out_host.numel() * out_host.element_size()

    peaks = load_peaks()
    line = {
        "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD_C2, "prefill_tokens": T, "image_tokens": L, "new_tokens": new_tokens,
                   "parallelism": "single GPU" if world == 1 else
                   f"decoder tp{world} (all-reduce x56/forward: NCCL at prefill, in-kernel NVLink exchange at decode), vision replicated",
                   "kv_cache": f"paged, page {cfg.kv_page_size}, shuffled block table",
                   "decode": "persistent megakernel, 1 launch/token" if dec.use_mega(1) else ("cuda graph" if not args.no_graph else "eager"),
                   "l2": "no flush needed: every decode step streams 14.2 GB of weights, the ViT pass 11 GB (>> 126 MB L2)"},
        "phases": {"vit_projector_prefill_ms": vp_ms, "images_per_sec_vit_prefill": 1000.0 / vp_ms,
                   "decode_ms": dec_ms, "decode_tokens_per_sec": (new_tokens - 1) / (dec_ms * 1e-3),
                   "decode_ms_per_token": dec_ms / (new_tokens - 1)},
        "e2e": {"value": e2e_val, "unit": "tokens/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "clocks": clocks.summary(),
    }
    # ---- rooflines (dominant kernels, CUDA events, live)
    if dec.use_mega(1):
        # the whole decode step is ONE persistent kernel launch: its duration is the decode-phase event time per token
        w = dec.w
        wbytes = sum(l.qkv_w.numel() + l.o_w.numel() + l.gate_up_w.numel() + l.down_w.numel() + l.qkv_b.numel()
                     + l.ln1.numel() + l.ln2.numel() for l in w.layers) * 2 + (w.lm_head.numel() + w.norm.numel()) * 2
        kv_per_tok = 2 * len(w.layers) * dec.Hkv * 128 * 2
        ctx_avg = T + (new_tokens - 1) / 2.0
        bytes_per_launch = wbytes + ctx_avg * kv_per_tok
        us = 1000.0 * dec_ms / (new_tokens - 1)
        gbs = bytes_per_launch / (us * 1e-6) / 1e9
        line["roofline"] = {"bound": "hbm", "kernel": "decode_mega_kernel<1> (one persistent launch per token)",
                            "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                            "traffic": load_traffic(), "peak_source": peaks["source"] + " (MEASURED_PEAKS.json hbm_gbs, burst copy)",
                            "launches_per_step": new_tokens - 1, "bytes_per_launch": bytes_per_launch,
                            "avg_launch_us": us, "share_of_step": dec_ms / ms_per_step,
                            "algorithmic_bytes": "bf16 weights of 28 layers + final norm + lm_head (14.14 GB) + KV cache "
                                                 "read at the mean context (57 344 B/token)"}
    else:
        gv = probe_gemv_roofline(model, torch, lib)
        line["roofline"] = {"bound": "hbm", "kernel": "gemv_bf16_kernel", "achieved": gv["gbs"], "peak": peaks["hbm_gbs"],
                            "unit": "GB/s", "frac": gv["gbs"] / peaks["hbm_gbs"], "traffic": None,
                            "peak_source": peaks["source"] + " (MEASURED_PEAKS.json hbm_gbs, burst copy)",
                            "launches_per_step": gv["launches"], "bytes_per_launch": gv["bytes_per_launch"],
                            "avg_launch_us": gv["avg_launch_us"],
                            "share_of_decode_step": gv["train_ms"] / (dec_ms / (new_tokens - 1)),
                            "decode_step_frac_of_hbm_peak": (gv["bytes_per_launch"] * gv["launches"]) /
                            (dec_ms / (new_tokens - 1) * 1e-3) / 1e9 / peaks["hbm_gbs"]}
    # tensor side of the SAME timed step: algorithmic FLOPs of the c2 ViT + projector + prefill phase over
    # its measured time - the number to quote for c2 (the probe below runs the ViT GEMM shapes at 8 crops instead)
    vc = cfg.vision_config
    Sv, Cv, Iv = vc.num_patches + 1, vc.hidden_size, vc.intermediate_size
    down = cfg.pixel_shuffle_down
    flop_vit = vc.num_hidden_layers * (2.0 * Sv * (4 * Cv * Cv + 2 * Cv * Iv) + 4.0 * Sv * Sv * Cv) + 2.0 * vc.num_patches * 588 * Cv
    flop_proj = 2.0 * L * (Cv * down * down * cfg.hidden_size + cfg.hidden_size ** 2)
    p_mm = sum(l.qkv_w.numel() + l.o_w.numel() + l.gate_up_w.numel() + l.down_w.numel() for l in dec.w.layers)
    flop_prefill = 2.0 * p_mm * T + 2.0 * len(dec.w.layers) * dec.Hq * 128 * T * T + 2.0 * dec.w.lm_head.numel()
    tf_phase = (flop_vit + flop_proj + flop_prefill) / (vp_ms * 1e-3) / 1e12
    line["roofline_tensor_c2_phase"] = {
        "bound": "tensor", "kernel": "ViT + projector + prefill phase of the timed request (tcgen05 GEMMs + attention, per GPU)",
        "achieved": tf_phase, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
        "frac": tf_phase / peaks["bf16_tflops_sustained"], "flops_per_step": flop_vit + flop_proj + flop_prefill,
        "phase_ms": vp_ms, "peak_source": peaks["source"] + " (bf16_tflops_sustained)"}
    gm = probe_gemm_roofline(model, torch, lib)
    if gm:
        line["roofline_tensor"] = {"bound": "tensor", "kernel": "gemm_bf16_kernel<256,2>", "achieved": gm["tflops"],
                                   "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                                   "frac": gm["tflops"] / peaks["bf16_tflops_sustained"],
                                   "peak_source": peaks["source"] + " (bf16_tflops_sustained: kernel timed inside a long train)",
                                   "flops_per_launch": gm["flops_per_launch"], "avg_launch_us": gm["avg_launch_us"],
                                   "shapes": "ViT block GEMMs, 8 crops (M=8200)"}
    # ---- the workloads that shard naturally, on the same weights, each with its own timing:
    # c3 = 64 crops data-parallel over the N towers (images/s, tensor roofline), c4 = 32 x 1024-token prefill + batch-32
    # decode with the decoder tensor-parallel over the N GPUs (tokens/s, HBM roofline)
    if not args.no_workloads:
        from tools.bench_workloads import measure_c3, measure_c4
        del one_request.cache
        torch.cuda.empty_cache()
        wl = {}
        wl["c3"] = measure_c3(args, rank, world, local, cfg, model.get_vision_tower(), model.get_model().mm_projector)
        torch.cuda.empty_cache()
        wl["c4"] = measure_c4(args, rank, world, local, cfg, dec)
        if world > 1 and not args.no_c5:
            # c5 (16 multi-image 4k-context requests) as data-parallel replicas over the N GPUs, through the public generate()
            from tools.bench_workloads import measure_c5
            torch.cuda.empty_cache()
            wl["c5"] = measure_c5(args, rank, world, local)
        keep = ("metric", "value", "unit", "ms_per_step", "scaling", "config", "phases", "e2e", "gpu_launches", "roofline", "error")
        line["workloads"] = {k: {kk: v[kk] for kk in keep if kk in v} for k, v in wl.items()}
    # ---- the reference's two lighter model families on one GPU, each in its own right:
    # random-init Qwen1.5-MoE-A2.7B-sized decoder (prefill + decode steps) and the InternViT-300M tower. Never allowed to
    # break the headline line.
    if world == 1 and not args.no_workloads and not args.no_variants:
        var = {}
        try:
            torch.cuda.empty_cache()
            from tools.bench_moe import measure as measure_moe
            recs = measure_moe(batches=(1, 32), steps=32)
            var["qwen2_moe_a2.7b"] = {"prefill": recs[0], "decode": recs[1:],
                                      "config": "24 layers, hidden 2048, 60 experts top-4 x 1408, shared 5632, random init, bf16"}
        except Exception as e:  # noqa: BLE001
            var["qwen2_moe_a2.7b"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        try:
            torch.cuda.empty_cache()
            from tools.bench_vit300m import measure as measure_300m
            var["internvit_300m"] = measure_300m(crops=64, iters=3)
        except Exception as e:  # noqa: BLE001
            var["internvit_300m"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        line["variants"] = var
    if world > 1:
        dist.barrier()
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            from oracle.cpu_baseline import time_c2_sample
            r = time_c2_sample(new_tokens=new_tokens, vit_layers=1, llm_layers=1, decode_steps=3, threads=host_threads())
            line["cpu_baseline"] = {"value": r["tokens_per_sec_request"], "unit": "tokens/s", "cores": r["cores"],
                                    "kind": "port", "sample": r["sample"], "extrapolated": True,
                                    "decode_tokens_per_sec": r["decode_tokens_per_sec"],
                                    "images_per_sec_vit_prefill": r["images_per_sec_vit_prefill"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        _TEARDOWN.append(lambda: _release_decoder(dec))
        _finish(torch, dist)


if __name__ == "__main__":
    main()
