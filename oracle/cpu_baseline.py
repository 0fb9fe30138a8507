"""CPU timing of the oracle (the fp32 restatement of the reference's PyTorch path) — TEST / BENCH INFRASTRUCTURE ONLY.

Used by bench.py's `cpu_baseline` leg and by `bench.py --impl reference`. The reference itself is Python calling
torch.nn modules in fp32 on CPU (SURVEY.md §8d "CPU baseline beside it"); it cannot be installed on the GPU box
(it needs timm/peft/accelerate shims and /root/reference is absent there), so the timed code is this repository's
restatement of the same torch ops (kind = "port"): identical ATen kernels (mkldnn/oneDNN sgemm, softmax, erf-GELU), the
same shapes and dtypes, all host threads.

The full model in fp32 is 22 GB (ViT) + 30 GB (decoder): a BOUNDED SAMPLE is timed instead — a few full-width layers
of each tower at the real sequence lengths — and scaled linearly to the full depth (every layer of a tower does
identical work). The sample that was timed is reported in the `sample` string.
"""
from __future__ import annotations

import time
from typing import Dict

import torch

from . import omchat_oracle as O


def _rand_vit_layers(cfg: O.OracleConfig, n: int, g) -> Dict[str, torch.Tensor]:
    C, I = cfg.vit_hidden, cfg.vit_inter
    sd = {}
    for li in range(n):
        p = f"{O.VT}encoder.layers.{li}."
        sd[p + "norm1.weight"] = torch.ones(C)
        sd[p + "norm2.weight"] = torch.ones(C)
        sd[p + "attn.qkv.weight"] = torch.randn(3 * C, C, generator=g) * 0.02
        sd[p + "attn.q_norm.weight"] = torch.ones(C)
        sd[p + "attn.k_norm.weight"] = torch.ones(C)
        sd[p + "attn.proj.weight"] = torch.randn(C, C, generator=g) * 0.02
        sd[p + "attn.proj.bias"] = torch.zeros(C)
        sd[p + "ls1"] = torch.full((C,), 0.1)
        sd[p + "ls2"] = torch.full((C,), 0.1)
        sd[p + "mlp.fc1.weight"] = torch.randn(I, C, generator=g) * 0.02
        sd[p + "mlp.fc1.bias"] = torch.zeros(I)
        sd[p + "mlp.fc2.weight"] = torch.randn(C, I, generator=g) * 0.02
        sd[p + "mlp.fc2.bias"] = torch.zeros(C)
    return sd


def _rand_llm_layers(cfg: O.OracleConfig, n: int, g) -> Dict[str, torch.Tensor]:
    H, I, D = cfg.hidden, cfg.inter, cfg.head_dim
    sd = {}
    for li in range(n):
        p = f"model.layers.{li}."
        sd[p + "self_attn.q_proj.weight"] = torch.randn(cfg.heads * D, H, generator=g) * 0.02
        sd[p + "self_attn.q_proj.bias"] = torch.zeros(cfg.heads * D)
        sd[p + "self_attn.k_proj.weight"] = torch.randn(cfg.kv_heads * D, H, generator=g) * 0.02
        sd[p + "self_attn.k_proj.bias"] = torch.zeros(cfg.kv_heads * D)
        sd[p + "self_attn.v_proj.weight"] = torch.randn(cfg.kv_heads * D, H, generator=g) * 0.02
        sd[p + "self_attn.v_proj.bias"] = torch.zeros(cfg.kv_heads * D)
        sd[p + "self_attn.o_proj.weight"] = torch.randn(H, cfg.heads * D, generator=g) * 0.02
        sd[p + "mlp.gate_proj.weight"] = torch.randn(I, H, generator=g) * 0.02
        sd[p + "mlp.up_proj.weight"] = torch.randn(I, H, generator=g) * 0.02
        sd[p + "mlp.down_proj.weight"] = torch.randn(H, I, generator=g) * 0.02
        sd[p + "input_layernorm.weight"] = torch.ones(H)
        sd[p + "post_attention_layernorm.weight"] = torch.ones(H)
    sd["model.norm.weight"] = torch.ones(H)
    return sd


@torch.no_grad()
def time_c2_sample(prefill_len: int = 1088, new_tokens: int = 256, vit_layers: int = 2, llm_layers: int = 2,
                   decode_steps: int = 3, threads: int = 0, full: O.OracleConfig = None) -> dict:
    """Times the c2 request (1 crop -> ViT -> projector -> prefill(prefill_len) -> new_tokens greedy tokens) on the CPU
    through the oracle, on a sample of `vit_layers` / `llm_layers` full-width layers, and extrapolates to the full
    depth. Returns seconds per phase for the FULL model and the derived throughput figures."""
    full = full or O.OracleConfig()
    if threads > 0:
        torch.set_num_threads(threads)
    cores = torch.get_num_threads()
    g = torch.Generator().manual_seed(0)
    C = full.vit_hidden
    S = (full.image_size // full.patch_size) ** 2 + 1
    # ---- ViT blocks on one crop
    cfg_v = O.OracleConfig(vit_layers=vit_layers)
    sd = _rand_vit_layers(cfg_v, vit_layers, g)
    x = torch.randn(1, S, C, generator=g)
    O.vit_layer(x, sd, 0, cfg_v)  # warm-up (thread pool, oneDNN primitives)
    t0 = time.perf_counter()
    for li in range(vit_layers):
        x = O.vit_layer(x, sd, li, cfg_v)
    t_vit_layer = (time.perf_counter() - t0) / vit_layers
    del sd
    # ---- projector
    H = full.hidden
    psd = {"model.mm_projector.0.weight": torch.randn(H, C, generator=g) * 0.02, "model.mm_projector.0.bias": torch.zeros(H),
           "model.mm_projector.2.weight": torch.randn(H, H, generator=g) * 0.02, "model.mm_projector.2.bias": torch.zeros(H)}
    t0 = time.perf_counter()
    O.projector(x[:, 1:], psd)
    t_proj = time.perf_counter() - t0
    del psd
    # ---- decoder layers: prefill at the real length, then single-token steps against the grown cache
    cfg_t = O.OracleConfig(layers=llm_layers)
    sd = _rand_llm_layers(cfg_t, llm_layers, g)
    sd["lm_head.weight"] = torch.randn(8, H, generator=g) * 0.02  # lm_head is timed separately below
    emb = torch.randn(1, prefill_len, H, generator=g) * 0.02
    pos = torch.arange(prefill_len)[None]
    O.qwen2_forward(emb[:, :64], pos[:, :64], sd, cfg_t)  # warm-up
    t0 = time.perf_counter()
    _, past = O.qwen2_forward(emb, pos, sd, cfg_t)
    t_prefill_layer = (time.perf_counter() - t0) / llm_layers
    tok = torch.randn(1, 1, H, generator=g) * 0.02
    t0 = time.perf_counter()
    for i in range(decode_steps):
        _, past = O.qwen2_forward(tok, torch.tensor([[prefill_len + i]]), sd, cfg_t, past)
    t_decode_layer = (time.perf_counter() - t0) / (decode_steps * llm_layers)
    del sd, past
    # ---- lm_head on one row (fp32 [V, H] = 2.2 GB)
    lm = torch.randn(full.vocab, H, generator=g) * 0.02
    h1 = torch.randn(1, H, generator=g)
    torch.nn.functional.linear(h1, lm)
    t0 = time.perf_counter()
    for _ in range(3):
        torch.nn.functional.linear(h1, lm).argmax(-1)
    t_head = (time.perf_counter() - t0) / 3
    del lm
    t_vit = t_vit_layer * full.vit_layers
    t_prefill = t_prefill_layer * full.layers + t_head
    t_step = t_decode_layer * full.layers + t_head
    t_decode = t_step * (new_tokens - 1)
    total = t_vit + t_proj + t_prefill + t_decode
    return {
        "cores": cores,
        "seconds": {"vit": t_vit, "projector": t_proj, "prefill": t_prefill, "decode": t_decode, "request": total},
        "tokens_per_sec_request": new_tokens / total,
        "decode_tokens_per_sec": 1.0 / t_step,
        "images_per_sec_vit_prefill": 1.0 / (t_vit + t_proj + t_prefill),
        "sample": (f"fp32 oracle on {cores} host threads: {vit_layers}/{full.vit_layers} ViT blocks on 1 crop ({S} tokens), "
                   f"projector, {llm_layers}/{full.layers} decoder layers prefill T={prefill_len} + {decode_steps} decode "
                   f"steps, lm_head row x3; scaled linearly to full depth and {new_tokens} tokens"),
    }
