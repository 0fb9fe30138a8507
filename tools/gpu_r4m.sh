#!/bin/bash
mkdir -p gpurun_out
# one --set full capture each: the grouped (MoE experts) tcgen05 GEMM at prefill size and the 8-tokens-per-CTA router
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 26 -c 2 -o gpurun_out/r4m_grouped -f python tools/bench_moe.py --layers 2 --batch 1 --steps 1 --prefill 8 > gpurun_out/r4m_a.out 2> gpurun_out/r4m_a.err; echo "rc=$?"
ls -la gpurun_out/r4m_*.ncu-rep
