#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload c4 --steps 2 --warmup 3 > gpurun_out/r3i_c4_tp2.json 2> gpurun_out/r3i_c4_tp2.err; echo "c4 tp2 rc=$?"
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/r3i_c4_tp2.json") if l.startswith("{")][-1]
print("tp2", round(d["value"]), {k: round(v,2) for k,v in d["phases"].items()})
PY
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/prof_stream.py --batch 32 --ctx 1024 --layers 4 > gpurun_out/r3i_timeline_tp2.txt 2>&1
grep -v "^\*\|OMP_NUM\|^\[rank\|^$\|NCCL" gpurun_out/r3i_timeline_tp2.txt | sed -n 5,10p
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/tp_forward_check.py 2>&1 | grep '"world"'; echo "forward check rc=${PIPESTATUS[0]}"
