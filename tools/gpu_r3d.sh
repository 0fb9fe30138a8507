#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r3d_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r3d_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r3d_bench.json 2> gpurun_out/r3d_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r3d_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r3d_bench.json").read().strip().splitlines()[-1])
print("c2", round(d["value"],1), round(d["e2e"]["value"],1), {k: round(v,2) for k,v in d["phases"].items()}, "frac", round(d["roofline"]["frac"],3), "launches", d["gpu_launches"])
print("tensor c2 phase", round(d["roofline_tensor_c2_phase"]["achieved"]), round(d["roofline_tensor_c2_phase"]["frac"],3))
for k,v in d["workloads"].items(): print(k, round(v["value"],1), v.get("phases") and {a: round(b,2) for a,b in v["phases"].items()}, round(v["roofline"]["frac"],3))
print(d["cpu_baseline"]["value"], d["clocks"])
PY
