#!/usr/bin/env python
"""A/B of this repository's tcgen05 flash attention (csrc/attention_sm100.cu) against the attention kernel the reference
itself calls - flash_attn_varlen_qkvpacked_func of flash-attn 2.8.3 (its sm_100 cubin is the FA2 mma.sync algorithm
recompiled; reference call site intern_vit_6b/flash_attention.py:43-55) - on the same box and the same shapes:
the ViT shape (non-causal, 1025 tokens, 25 heads, packed qkv) at 64 / 8 / 1 crops and the prefill shape (causal GQA
28q / 4kv; flash_attn_varlen_func) at 32 x 1024 and 1 x 1088 tokens. CUDA events over `reps` launches after warm-up,
both on the current stream; outputs are compared with each other. MEASUREMENT TOOL: flash-attn is never on the product path.
    python tools/bench_attention_ab.py [--reps 10] > profiles/rNN_attention_vs_flash_attn.txt"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from omchat_b200 import lib  # noqa: E402


def timeit(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    from flash_attn import __version__ as fa_version
    from flash_attn import flash_attn_varlen_func, flash_attn_varlen_qkvpacked_func
    torch.cuda.set_device(0)
    lib.load()
    print(f"# flash-attn {fa_version} vs omchat_b200 tcgen05 attention, {torch.cuda.get_device_name(0)}, CUDA events, {a.reps} launches")
    print(f"{'shape':44s} {'ours ms':>9s} {'ours TF/s':>10s} {'flash ms':>9s} {'flash TF/s':>11s} {'ours/flash':>10s} {'max |diff|':>10s}")
    scale = 128 ** -0.5
    g = torch.Generator(device="cuda").manual_seed(0)
    for n_seq, S, Hq, Hkv, causal in [(64, 1025, 25, 25, False), (8, 1025, 25, 25, False), (1, 1025, 25, 25, False),
                                      (32, 1024, 28, 4, True), (1, 1088, 28, 4, True)]:
        total = n_seq * S
        W = (Hq + 2 * Hkv) * 128
        qkv = (torch.randn(total, W, generator=g, device="cuda") * 0.5).to(torch.bfloat16)
        cu = (torch.arange(n_seq + 1, dtype=torch.int32) * S).cuda()
        out = torch.empty(total, Hq * 128, device="cuda", dtype=torch.bfloat16)
        q, k, v = qkv[:, :Hq * 128], qkv[:, Hq * 128:(Hq + Hkv) * 128], qkv[:, (Hq + Hkv) * 128:]

        def ours():
            lib.attention(q, k, v, out, cu, S, Hq, Hkv, causal, scale)

        if Hq == Hkv:
            packed = qkv.view(total, 3, Hq, 128)

            def flash():
                return flash_attn_varlen_qkvpacked_func(packed, cu, S, 0.0, softmax_scale=scale, causal=causal)
        else:
            q3, k3, v3 = q.reshape(total, Hq, 128), k.reshape(total, Hkv, 128), v.reshape(total, Hkv, 128)

            def flash():
                return flash_attn_varlen_func(q3, k3, v3, cu, cu, S, S, 0.0, softmax_scale=scale, causal=causal)
        t_o, t_f = timeit(ours, a.reps), timeit(flash, a.reps)
        ref = flash().reshape(total, Hq * 128)
        diff = (out.float() - ref.float()).abs().max().item()
        flops = 4.0 * n_seq * Hq * S * S * 128 * (0.5 if causal else 1.0)
        name = f"{n_seq} x {S} tok, {Hq}q/{Hkv}kv, {'causal' if causal else 'full'}"
        print(f"{name:44s} {t_o:9.3f} {flops / t_o / 1e9:10.1f} {t_f:9.3f} {flops / t_f / 1e9:11.1f} {t_f / t_o:10.2f} {diff:10.4f}")


if __name__ == "__main__":
    main()
