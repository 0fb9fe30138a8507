#!/bin/bash
OUT=gpurun_out/s6e; mkdir -p $OUT
timeout 600 python -m pytest tests/test_decode_mega_gpu.py -x -q 2>&1 | tail -15 | tee $OUT/pytest_mega.log
for M in 0 1 2; do
  OMCHAT_B200_MEGA_PROFMODE=$M timeout 200 python tools/prof_mega.py 28 1 1200 > $OUT/prof_m$M.log 2>&1
  echo "== profmode=$M"; sed -n 1,16p $OUT/prof_m$M.log | cut -c1-200
done
for NS in 0 100 300 0 100 300; do
  echo "== poll_ns=$NS"; OMCHAT_B200_MEGA_PROF=0 OMCHAT_B200_MEGA_POLLNS=$NS timeout 200 python tools/prof_mega.py 28 1 1200 2>&1 | tail -1
done
