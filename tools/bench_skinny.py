#!/usr/bin/env python
"""Skinny GEMM microbenchmark (decode shapes of Qwen2-7B at batch M): GB/s of weight streaming per shape, split-K on/off.
Each timed launch uses a different copy of W (more copies than fit in L2) so weights always come from HBM."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes  # noqa: E402

import torch  # noqa: E402

from omchat_b200 import lib  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 32
L = lib.load()
shapes = [("tinyK", 3584, 64, 0), ("k512", 3584, 512, 0), ("k1024", 3584, 1024, 0), ("k2048", 3584, 2048, 0), ("qkv", 4608, 3584, 0), ("o", 3584, 3584, 2), ("down", 3584, 18944, 2)]
for name, N, K, epi in shapes:
    copies = max(2, int(400e6 // (N * K * 2)) + 1)
    ws_list = [torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * 0.02 for _ in range(copies)]
    x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    n_out = N // 2 if epi == 3 else N
    out = torch.zeros(M, n_out, device="cuda", dtype=torch.bfloat16)
    res = torch.zeros(M, n_out, device="cuda", dtype=torch.bfloat16)
    wsp = lib._skinny_workspace(x.device, N)
    for use_ws in (True, False):
        def call(w):
            rc = L.omc_gemm_skinny_bf16(x.data_ptr(), K, w.data_ptr(), K, out.data_ptr(), n_out, M, N, K, None, None,
                                        res.data_ptr() if epi == 2 else None, n_out, epi, 0,
                                        wsp.data_ptr() if use_ws else None, wsp.numel() if use_ws else 0,
                                        torch.cuda.current_stream().cuda_stream)
            assert rc == 0, L.omc_last_error()
        for w in ws_list:
            call(w)
        torch.cuda.synchronize()
        reps = 3
        g = torch.cuda.CUDAGraph()  # replayed graph: no host launch cost between the kernels (like the decode step)
        with torch.cuda.graph(g):
            for w in ws_list:
                call(w)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1000 / (reps * copies)
        print(f"M={M} {name:8s} N={N:6d} K={K:5d} split-K {'on ' if use_ws else 'off'}: {us:8.1f} us  {N * K * 2 / us / 1e3:8.1f} GB/s")
