#!/bin/bash
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 > gpurun_out/r4g_full_gpu.log; cat gpurun_out/r4g_full_gpu.log; echo "full suite in $(( $(date +%s) - t0 )) s"
bash tools/gpu_sanitize_r3.sh 2>&1 | tee gpurun_out/r4g_sanitize.log
