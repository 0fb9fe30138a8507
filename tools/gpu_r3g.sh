#!/bin/bash
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/tp_forward_check.py 2>&1 | grep -v "OMP_NUM\|^\*\*\*\|Traceback\|File \|    " | tail -14
