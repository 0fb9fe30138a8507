#!/bin/bash
OUT=gpurun_out/s6c; mkdir -p $OUT
for SC in 0 1; do
  OMCHAT_B200_MEGA_SCALAR=$SC timeout 200 python tools/prof_mega.py 28 1 1200 > $OUT/prof_sc$SC.log 2>&1
  echo "== scalar=$SC"; head -20 $OUT/prof_sc$SC.log | cut -c1-200
  OMCHAT_B200_MEGA_PROF=0 OMCHAT_B200_MEGA_SCALAR=$SC timeout 200 python tools/prof_mega.py 28 1 1200 2>&1 | tail -1
done
