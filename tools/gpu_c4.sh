#!/bin/bash
# batch-32 decode path (c4): skinny-GEMM parity, per-kernel launch list of one decode step, c4 bench on one GPU
TAG=${1:-c4}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "skinny or decode_attn or paged" 2>&1 | tail -5 | tee $OUT/pytest.log
timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum --kernel-name-base demangled -k regex:omc:: -s 40 -c 40 --csv --log-file $OUT/launches_decode_b32.csv python tools/prof_step.py decode --batch 32 --layers 4 --steps 3 --ctx 1024 > $OUT/prof.log 2>&1; tail -1 $OUT/prof.log
timeout 600 python bench.py --workload c4 --steps 2 --warmup 3 > $OUT/c4_n1.json 2> $OUT/c4_n1.err; tail -c 1800 $OUT/c4_n1.json; tail -3 $OUT/c4_n1.err
