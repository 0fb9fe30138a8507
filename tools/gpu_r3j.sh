#!/bin/bash
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
t0=$(date +%s)
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r3j_bench_n8.out 2> gpurun_out/r3j_bench_n8.err
echo "bench n8 rc=$? in $(( $(date +%s) - t0 )) s"; grep -i "error\|Traceback\|watchdog" gpurun_out/r3j_bench_n8.err | head -5
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/r3j_bench_n8.out") if l.startswith('{"metric')][-1]
print("c2", round(d["value"],1), round(d["e2e"]["value"],1), {k: round(v,2) for k,v in d["phases"].items()})
for k,v in d["workloads"].items(): print(k, round(v["value"],1), v.get("phases") and {a: round(b,2) for a,b in v["phases"].items()}, round(v["roofline"]["frac"],3))
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --workload c4 --steps 2 --warmup 3 > gpurun_out/r3j_c4_tp4.json 2> gpurun_out/r3j_c4_tp4.err; echo "c4 tp4 rc=$?"
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/r3j_c4_tp4.json") if l.startswith("{")][-1]
print("tp4", round(d["value"]), {k: round(v,2) for k,v in d["phases"].items()})
PY
