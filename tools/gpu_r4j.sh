#!/bin/bash
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
N=${1:-2}
t0=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r4j_bench_n$N.out 2> gpurun_out/r4j_bench_n$N.err
echo "bench n$N rc=$? in $(( $(date +%s) - t0 )) s"; grep -i "error\|Traceback" gpurun_out/r4j_bench_n$N.err | head -5
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/r4j_bench_n$N.out") if l.startswith('{"metric')][-1]
print("c2", round(d["value"],1), round(d["e2e"]["value"],1), {k: round(v,2) for k,v in d["phases"].items()})
for k,v in d["workloads"].items():
    if "error" in v: print(k, v); continue
    print(k, round(v["value"],1), v.get("phases") and {a: round(b,2) for a,b in v["phases"].items()}, v.get("roofline") and round(v["roofline"]["frac"],3))
PY
