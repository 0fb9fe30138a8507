#!/bin/bash
# 8 GPUs: default bench line at N=8 (c2 TP8 + workloads c3 DP8 / c4 TP8), then c5 in both modes
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
t0=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2u_bench_n8.out 2> gpurun_out/r2u_bench_n8.err
echo "bench n8 rc=$? in $(( $(date +%s) - t0 )) s"; grep -i "error\|Traceback\|watchdog" gpurun_out/r2u_bench_n8.err | head -5
python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/r2u_bench_n8.out") if l.startswith('{"metric')][-1]
    print("c2", round(d["value"],1), round(d["e2e"]["value"],1), {k: round(v,2) for k,v in d["phases"].items()})
    for k,v in d["workloads"].items(): print(k, round(v["value"],1), v.get("phases") and {a: round(b,2) for a,b in v["phases"].items()}, round(v["roofline"]["frac"],3))
except Exception as e: print("n8 parse failed", e)
PY
for mode in tp dp; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29539 bench.py --gpus 8 --workload c5 --c5-mode $mode --steps 2 --warmup 3 > gpurun_out/r2u_c5_$mode.out 2> gpurun_out/r2u_c5_$mode.err
  echo "c5 $mode rc=$?"; grep -i "error\|Traceback" gpurun_out/r2u_c5_$mode.err | head -3
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/r2u_c5_$mode.out") if l.startswith('{"metric')][-1]
    print("c5 $mode", round(d["value"],1), d["config"]["parallelism"], {k: round(v,2) for k,v in d["phases"].items()}, "e2e", round(d["e2e"]["value"],1))
except Exception as e: print("c5 $mode parse failed", e)
PY
done
