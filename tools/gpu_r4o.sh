#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_moe_gpu.py tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_vit300m_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/bench_moe.py --batch 1 --batch 8 --batch 32 > gpurun_out/r4o_moe_bench.jsonl 2> gpurun_out/r4o_moe_bench.err; echo "moe rc=$?"; cat gpurun_out/r4o_moe_bench.jsonl | cut -c1-330; tail -3 gpurun_out/r4o_moe_bench.err
