#!/bin/bash
OUT=gpurun_out/s6d; mkdir -p $OUT
for M in 1 2; do
  OMCHAT_B200_MEGA_PROFMODE=$M timeout 200 python tools/prof_mega.py 28 1 1200 > $OUT/prof_m$M.log 2>&1
  echo "== profmode=$M"; sed -n 1,1p $OUT/prof_m$M.log; sed -n 10,16p $OUT/prof_m$M.log | cut -c1-200
done
