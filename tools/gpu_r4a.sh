#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vit300m_gpu.py tests/test_model_capi_gpu.py tests/test_model_gpu.py -x -q -m gpu -s 2>&1 | grep -v "^$" | tail -60 > gpurun_out/r4a_tests.log; tail -45 gpurun_out/r4a_tests.log
