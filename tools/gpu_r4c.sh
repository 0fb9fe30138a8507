#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_moe_gpu.py -x -q -m gpu -s 2>&1 | grep -v "^$" > gpurun_out/r4c_moe.log; grep -n "max-abs\|routing\|cuda:\|passed\|failed\|Error\|seq " gpurun_out/r4c_moe.log | cut -c1-420 | tail -60
