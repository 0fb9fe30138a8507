#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_moe_gpu.py -x -q -m gpu -s 2>&1 | grep -v "^$" > gpurun_out/r4b_moe.log; tail -40 gpurun_out/r4b_moe.log
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_decode_mega_gpu.py tests/test_kernels_gpu.py tests/test_hf_auto.py -x -q -m gpu 2>&1 | tail -5
