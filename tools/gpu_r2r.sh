#!/bin/bash
# 4 GPUs: c4 at TP4 and TP2 with the fused NVLink all-reduce (watchdog turns a hang into an error after 10 s)
mkdir -p gpurun_out
run() { n=$1; tag=$2; shift 2; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --workload c4 --steps 2 --warmup 3 > gpurun_out/r2r_c4_tp${n}_$tag.json 2> gpurun_out/r2r_c4_tp${n}_$tag.err; echo "tp$n $tag rc=$?"; grep -v "OMP_NUM\|^\*" gpurun_out/r2r_c4_tp${n}_$tag.err | tail -3; python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/r2r_c4_tp${n}_$tag.json") if l.startswith("{")][-1]
    print("tp$n $tag", round(d["value"]), {k: round(v, 2) for k, v in d["phases"].items()}, round(d["roofline"]["frac"], 3))
except Exception as e: print("tp$n $tag failed", e)
PY
}
run 4 fused OMCHAT_B200_TP_STREAM_FUSED=1
run 2 fused OMCHAT_B200_TP_STREAM_FUSED=1
run 4 nccl OMCHAT_B200_TP_STREAM_FUSED=0
