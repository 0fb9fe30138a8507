#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_model_capi_gpu.py tests/test_full_width_parity_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --workload c3 --steps 3 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('c3', round(d['value'],1), 'img/s', round(d['roofline']['frac'],3), d['clocks'])"
OMCHAT_B200_GEMM_AUTOTUNE=1 timeout 300 python bench.py --no-workloads --no-variants --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('c2', round(d['value'],1), d['phases'])"
