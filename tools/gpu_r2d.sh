#!/bin/bash
mkdir -p gpurun_out
for mode in stream old; do
  export OMCHAT_B200_NO_STREAM=0
  [ $mode = old ] && export OMCHAT_B200_NO_STREAM=1
  timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum --kernel-name-base demangled -k regex:omc:: -s 60 -c 30 --csv --log-file gpurun_out/r2d_launches_b32_$mode.csv python tools/prof_step.py decode --batch 32 --layers 4 --steps 3 --ctx 1024 > gpurun_out/r2d_prof_$mode.log 2>&1
  tail -1 gpurun_out/r2d_prof_$mode.log
  python tools/ncu_summary.py launches gpurun_out/r2d_launches_b32_$mode.csv | tail -14
done
unset OMCHAT_B200_NO_STREAM
OMCHAT_FULL_PARITY=1 timeout 1500 python -m pytest tests/test_full_width_parity_gpu.py -m gpu -x -q -s -k full_depth > gpurun_out/r2d_full_depth.log 2>&1
echo "full depth rc=$?"; grep "FAIL\|equal\|passed\|failed" gpurun_out/r2d_full_depth.log | head -20
