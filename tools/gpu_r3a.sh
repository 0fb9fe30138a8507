#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "folded_rmsnorm or gemm_plain or gemm_swiglu" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_model_capi_gpu.py tests/test_model_gpu.py tests/test_full_width_parity_gpu.py -m gpu -x -q 2>&1 | tail -8
for f in 1 0; do OMCHAT_B200_FOLD_NORMS=$f timeout 600 python bench.py --steps 3 --warmup 3 --no-workloads --no-cpu-baseline > gpurun_out/r3a_bench_fold$f.json 2>gpurun_out/r3a_err.log; python - <<PY
import json
d=json.loads(open("gpurun_out/r3a_bench_fold$f.json").read().strip().splitlines()[-1])
print("fold=$f", round(d["value"],1), {k: round(v,2) for k,v in d["phases"].items()}, d["gpu_launches"])
PY
done
for f in 1 0; do OMCHAT_B200_FOLD_NORMS=$f timeout 600 python bench.py --workload c3 --steps 3 --warmup 3 > gpurun_out/r3a_c3_fold$f.json 2>gpurun_out/r3a_err.log; python - <<PY
import json
d=json.loads(open("gpurun_out/r3a_c3_fold$f.json").read().strip().splitlines()[-1])
print("c3 fold=$f", round(d["value"],2), round(d["roofline"]["frac"],3))
PY
done
