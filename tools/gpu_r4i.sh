#!/bin/bash
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 > gpurun_out/r4i_full_gpu.log; cat gpurun_out/r4i_full_gpu.log; echo "full suite in $(( $(date +%s) - t0 )) s"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/r4i_bench.out 2> gpurun_out/r4i_bench.err; echo "bench rc=$? in $(( $(date +%s) - t0 )) s"; tail -3 gpurun_out/r4i_bench.err
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/r4i_bench.out") if l.startswith('{"metric')][-1]
print("c2", round(d["value"],1), round(d["e2e"]["value"],1), "roofline", round(d["roofline"]["frac"],3), "launches", d["gpu_launches"], d["clocks"])
for k,v in d["workloads"].items(): print(k, round(v["value"],1), round(v["roofline"]["frac"],3))
print(json.dumps(d.get("variants"))[:1500])
print("cpu_baseline", d.get("cpu_baseline"))
PY
