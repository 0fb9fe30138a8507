#!/bin/bash
# 2 GPUs: c4 (decoder TP2, 32 x 1024 prefill + batch-32 decode) with the fused NVLink all-reduce vs NCCL between kernels
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload c4 --steps 2 --warmup 3 > gpurun_out/r2m_c4_tp2_$tag.json 2> gpurun_out/r2m_c4_tp2_$tag.err; echo "$tag rc=$?"; tail -2 gpurun_out/r2m_c4_tp2_$tag.err; python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/r2m_c4_tp2_$tag.json") if l.startswith("{")][-1]
    print("$tag", d["value"], d["phases"], d["roofline"]["frac"])
except Exception as e: print("$tag failed", e)
PY
}
run fused OMCHAT_B200_TP_STREAM_FUSED=1
run nccl OMCHAT_B200_TP_STREAM_FUSED=0
