#!/bin/bash
export NCCL_DEBUG=WARN
timeout 500 python -m pytest tests/test_tp_multi_gpu.py -x -q -m gpu 2>&1 | tail -3
bash tools/gpu_r4j.sh 2
