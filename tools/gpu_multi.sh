#!/bin/bash
# multi-GPU bench pass: usage  gpurun --gpus N --timeout 1200 -- 'bash tools/gpu_multi.sh <tag> N [workloads...]'
TAG=${1:-multi}; N=${2:-2}; shift; shift
WL=${@:-c2 c3 c4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
for w in $WL; do
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --workload $w --steps 3 --warmup 3 > $OUT/${w}_n$N.json 2> $OUT/${w}_n$N.err
  echo "== $w N=$N rc=$?"; tail -c 1500 $OUT/${w}_n$N.json; grep -v "^W\|^\[W\|^$" $OUT/${w}_n$N.err | tail -8
done
true
