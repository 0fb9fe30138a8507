#!/bin/bash
# One gpurun call: GPU tests, smoke, bench, ncu launch lists and full captures of the dominant kernels.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r01b'
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee $OUT/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; tail -c 3000 $OUT/bench.json
NCU="ncu --clock-control none"
# launch lists (cold-cache, serialised: compare shares)
timeout 600 $NCU --metrics gpu__time_duration.sum --kernel-name-base demangled -k regex:omc:: -s 2 -c 4 --csv --log-file $OUT/launches_decode.csv python tools/prof_step.py decode --steps 6 > $OUT/prof_decode.log 2>&1
timeout 600 $NCU --metrics gpu__time_duration.sum --kernel-name-base demangled -k regex:omc:: -s 50 -c 60 --csv --log-file $OUT/launches_vit.csv python tools/prof_step.py vit --crops 8 --layers 4 --steps 2 > $OUT/prof_vit.log 2>&1
timeout 600 $NCU --metrics gpu__time_duration.sum --kernel-name-base demangled -k regex:omc:: -s 36 -c 40 --csv --log-file $OUT/launches_prefill.csv python tools/prof_step.py prefill --layers 4 --steps 2 > $OUT/prof_prefill.log 2>&1
# full captures of the dominant kernels
timeout 900 $NCU --set full --import-source on -k regex:decode_mega -s 2 -c 1 -o $OUT/mega_full -f python tools/prof_step.py decode --steps 4 > $OUT/ncu_mega.log 2>&1
OMCHAT_B200_NO_MEGA=1 timeout 900 $NCU --set full --import-source on -k regex:gemv -s 12 -c 5 -o $OUT/gemv_full -f python tools/prof_step.py decode --layers 4 --steps 2 > $OUT/ncu_gemv.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:gemm_bf16 -s 17 -c 4 -o $OUT/gemm_full -f python tools/prof_step.py vit --crops 8 --layers 4 --steps 2 > $OUT/ncu_gemm.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:attention_fwd -s 4 -c 1 -o $OUT/attn_full -f python tools/prof_step.py vit --crops 8 --layers 4 --steps 2 > $OUT/ncu_attn.log 2>&1
ls -la $OUT
