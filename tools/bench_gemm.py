"""Micro-benchmark of omc_gemm_bf16 tile configurations on the hot path's GEMM shapes (CUDA events, rotating buffers
larger than L2 so weights/activations stream from HBM as in the real layer loop). cuBLAS (torch.matmul) is timed
beside it as a yardstick only. Usage: python tools/bench_gemm.py [--quick]"""
import json
import sys

import torch

from omchat_b200 import lib

SHAPES = {
    # name: (M, N, K, epi)
    "vit_qkv_b8": (8 * 1025, 9600, 3200, "none"),
    "vit_proj_b8": (8 * 1025, 3200, 3200, "res"),
    "vit_fc1_b8": (8 * 1025, 12800, 3200, "gelu"),
    "vit_fc2_b8": (8 * 1025, 3200, 12800, "res"),
    "vit_fc1_b1": (1025, 12800, 3200, "gelu"),
    "vit_fc2_b1": (1025, 3200, 12800, "res"),
    "llm_gateup_t1088": (1088, 37888, 3584, "swiglu"),
    "llm_down_t1088": (1088, 3584, 18944, "res"),
    "llm_qkv_t1088": (1088, 4608, 3584, "none"),
    # the single-request shapes (c2): one crop, one 1088-token prefill; "--small" runs only these
    "vit_qkv_b1": (1025, 9600, 3200, "none"),
    "vit_proj_b1": (1025, 3200, 3200, "res"),
    "llm_o_t1088": (1088, 3584, 3584, "res"),
    "proj_fc1_b1": (1024, 3584, 3200, "gelu"),
}
SMALL = ("vit_qkv_b1", "vit_proj_b1", "vit_fc1_b1", "vit_fc2_b1", "llm_qkv_t1088", "llm_o_t1088", "llm_gateup_t1088",
         "llm_down_t1088")
CFGS = [(256, 1), (128, 1), (256, 2), (192, 2), (160, 2), (128, 2)]


def time_fn(fn, nbuf, iters=12, warm=3):
    for i in range(warm):
        fn(i % nbuf)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for i, (a, b) in enumerate(evs):
        a.record()
        fn(i % nbuf)
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


def main():
    quick = "--quick" in sys.argv
    small = "--small" in sys.argv
    lib.load()
    res = []
    for name, (M, N, K, epi) in SHAPES.items():
        if quick and not name.endswith("b8"):
            continue
        if small and name not in SMALL:
            continue
        bytes_per = (M * K + N * K + M * N) * 2
        nbuf = max(2, int(300e6 // bytes_per) + 1)
        xs = [torch.randn(M, K, device="cuda").bfloat16() for _ in range(nbuf)]
        ws = [(torch.randn(N, K, device="cuda") * 0.05).bfloat16() for _ in range(nbuf)]
        n_out = N // 2 if epi == "swiglu" else N
        outs = [torch.zeros(M, n_out, device="cuda", dtype=torch.bfloat16) for _ in range(nbuf)]
        bias = torch.zeros(N, device="cuda", dtype=torch.bfloat16)
        flops = 2.0 * M * N * K
        t = time_fn(lambda i: torch.matmul(xs[i], ws[i].t()), nbuf)
        row = {"shape": name, "M": M, "N": N, "K": K, "cublas_tflops": round(flops / t / 1e9, 1)}
        for bn, cg in [(0, 0)] + CFGS:  # (0, 0) = the library's own choice
            if epi == "swiglu" and bn not in (0, 256):
                continue
            cfg = bn | (cg << 16)
            kw = {}
            if epi == "gelu":
                kw = dict(bias=bias, epi=lib.EPI_GELU)
            elif epi == "res":
                kw = dict(bias=bias, epi=lib.EPI_RES)
            elif epi == "swiglu":
                kw = dict(epi=lib.EPI_SWIGLU)
            try:
                if epi == "res":
                    t = time_fn(lambda i: lib.gemm(xs[i], ws[i], out=outs[i], res=outs[i], tile_cfg=cfg, **kw), nbuf)
                else:
                    t = time_fn(lambda i: lib.gemm(xs[i], ws[i], out=outs[i], tile_cfg=cfg, **kw), nbuf)
                row["auto" if bn == 0 else f"bn{bn}_cg{cg}"] = round(flops / t / 1e9, 1)
            except Exception as e:  # noqa: BLE001
                row[f"bn{bn}_cg{cg}"] = f"ERR {e}"
        print(json.dumps(row), flush=True)
        res.append(row)
        del xs, ws, outs
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
