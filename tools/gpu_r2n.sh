#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/prof_stream.py --batch 32 --ctx 1024 --layers 4 > gpurun_out/r2o_timeline_tp2.txt 2>&1
grep -v "^\*\|OMP_NUM" gpurun_out/r2o_timeline_tp2.txt | head -24
