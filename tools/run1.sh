set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r1_pytest.log; tail -15 gpurun_out/r1_pytest.log
python __graft_entry__.py --smoke 2>&1 | tail -5 | tee gpurun_out/r1_smoke.log
timeout 900 python bench.py --steps 2 --warmup 3 2>&1 | tail -20 | tee gpurun_out/r1_bench.log
