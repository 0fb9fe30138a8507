#!/bin/bash
# 2 GPUs: the default bench line at N=2 (c2 TP2 + workloads c3 DP2 / c4 TP2), NCCL_DEBUG as the driver sets it; clean exit?
mkdir -p gpurun_out
export NCCL_DEBUG=INFO
t0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2s_bench_n2.out 2> gpurun_out/r2s_bench_n2.err
echo "bench n2 rc=$? in $(( $(date +%s) - t0 )) s"
grep -c "NCCL INFO" gpurun_out/r2s_bench_n2.out gpurun_out/r2s_bench_n2.err
grep -i "destroy\|watchdog\|error\|Traceback" gpurun_out/r2s_bench_n2.err | head -5
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/r2s_bench_n2.out") if l.startswith('{"metric')][-1]
print("c2", round(d["value"],1), round(d["e2e"]["value"],1), {k: round(v,2) for k,v in d["phases"].items()})
for k,v in d["workloads"].items(): print(k, round(v["value"],1), v.get("phases") and {a: round(b,2) for a,b in v["phases"].items()}, round(v["roofline"]["frac"],3))
PY
python bench.py --impl reference --gpus 2 --steps 1 --warmup 1 | cut -c1-400
