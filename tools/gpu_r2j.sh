#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_mega_gpu.py tests/test_serving_gpu.py -m gpu -q 2>&1 | tail -4
OMCHAT_FULL_PARITY=1 timeout 1500 python -m pytest tests/test_full_width_parity_gpu.py -m gpu -x -q -s -k full_depth > gpurun_out/r2j_full_depth.log 2>&1
echo "full depth rc=$?"; grep "FAIL\|equal\|passed\|failed\|prompt seed" gpurun_out/r2j_full_depth.log | head -20
timeout 600 python tools/bench_attention_ab.py > gpurun_out/r2j_attention_ab.txt 2>&1; cat gpurun_out/r2j_attention_ab.txt
