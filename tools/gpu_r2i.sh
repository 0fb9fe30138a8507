#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2i_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r2i_pytest.log
OMCHAT_FULL_PARITY=1 timeout 1500 python -m pytest tests/test_full_width_parity_gpu.py -m gpu -x -q -s -k full_depth > gpurun_out/r2i_full_depth.log 2>&1
echo "full depth rc=$?"; grep "FAIL\|equal\|passed\|failed\|prompt seed" gpurun_out/r2i_full_depth.log | head -20
timeout 900 python bench.py > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2i_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2i_bench.json").read().strip().splitlines()[-1])
print("c2", d["value"], d["e2e"]["value"], d["phases"], d["roofline"]["frac"])
for k,v in d["workloads"].items(): print(k, v["value"], v.get("phases"), v["roofline"]["frac"])
print(d["cpu_baseline"])
PY
