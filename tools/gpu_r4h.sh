#!/bin/bash
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 600 python -m pytest tests/test_tp_multi_gpu.py tests/test_hf_auto.py -x -q -m gpu 2>&1 | tail -6
t0=$(date +%s)
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r4h_bench_n2.out 2> gpurun_out/r4h_bench_n2.err
echo "bench n2 rc=$? in $(( $(date +%s) - t0 )) s"; grep -i "error\|Traceback" gpurun_out/r4h_bench_n2.err | head -5
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/r4h_bench_n2.out") if l.startswith('{"metric')][-1]
print("c2", round(d["value"],1), round(d["e2e"]["value"],1), {k: round(v,2) for k,v in d["phases"].items()})
for k,v in d["workloads"].items(): print(k, round(v["value"],1), v.get("phases") and {a: round(b,2) for a,b in v["phases"].items()}, round(v["roofline"]["frac"],3))
PY
