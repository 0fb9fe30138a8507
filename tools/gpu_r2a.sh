#!/bin/bash
# round 2, call A: new parity tests first (fail fast), then the whole GPU suite, then the default bench line
mkdir -p gpurun_out
nproc > gpurun_out/r2a_host.txt; free -g >> gpurun_out/r2a_host.txt; nvidia-smi -L >> gpurun_out/r2a_host.txt
timeout 900 python -m pytest tests/test_full_width_parity_gpu.py tests/test_hf_auto.py -m gpu -x -q -s > gpurun_out/r2a_parity.log 2>&1
echo "parity rc=$?" | tee -a gpurun_out/r2a_parity.log
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_full_width_parity_gpu.py > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 900 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"
tail -c 3000 gpurun_out/r2a_bench.json
tail -20 gpurun_out/r2a_parity.log
