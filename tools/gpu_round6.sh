#!/bin/bash
# session-6 round: GPU tests, smoke, bench, ncu launch list + full capture of the persistent decode kernel, in-kernel timeline
TAG=${1:-r01s6}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee $OUT/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; tail -c 3500 $OUT/bench.json
timeout 200 python tools/prof_mega.py 28 1 1200 > $OUT/mega_timeline.txt 2>&1; head -12 $OUT/mega_timeline.txt | cut -c1-160
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum --kernel-name-base demangled -k regex:omc:: -s 2 -c 4 --csv --log-file $OUT/launches_decode.csv python tools/prof_step.py decode --steps 6 > $OUT/prof_decode.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:decode_mega -s 2 -c 1 -o $OUT/mega_full -f python tools/prof_step.py decode --steps 4 > $OUT/ncu_mega.log 2>&1
ls -la $OUT
