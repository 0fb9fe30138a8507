#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_full_width_parity_gpu.py tests/test_hf_auto.py -m gpu -x -q -s > gpurun_out/r2b_parity.log 2>&1
echo "parity rc=$?" | tee -a gpurun_out/r2b_parity.log
OMCHAT_FULL_PARITY=1 timeout 1500 python -m pytest tests/test_full_width_parity_gpu.py -m gpu -x -q -s -k full_depth > gpurun_out/r2b_full_depth.log 2>&1
echo "full depth rc=$?" | tee -a gpurun_out/r2b_full_depth.log
grep -v "decode step\|hidden state" gpurun_out/r2b_parity.log | tail -30
grep -v "decode step\|hidden state" gpurun_out/r2b_full_depth.log | tail -30
