#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_moe_gpu.py -x -q -m gpu -s 2>&1 | grep -v "^$" > gpurun_out/r4f_moe.log; grep -n "rows within\|cuda:\|passed\|failed\|Error\|seq \|clear rows" gpurun_out/r4f_moe.log | cut -c1-260 | tail -40
timeout 300 python tools/bench_moe.py --batch 1 --batch 8 --batch 32 > gpurun_out/r4f_moe_bench.jsonl 2> gpurun_out/r4f_moe_bench.err; echo "moe rc=$?"; cat gpurun_out/r4f_moe_bench.jsonl; tail -3 gpurun_out/r4f_moe_bench.err
