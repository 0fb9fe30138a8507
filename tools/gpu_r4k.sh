#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_vit300m_gpu.py tests/test_model_capi_gpu.py tests/test_model_gpu.py tests/test_full_width_parity_gpu.py -x -q -m gpu -k "attention or vit300m or 300m or capi or vision or full_width or reduced" 2>&1 | tail -8
timeout 200 python tools/bench_vit300m.py 2>&1 | tail -1
timeout 300 python tools/bench_attention_ab.py 2>&1 | tail -6
