#!/bin/bash
# session-6 call B: parity of the mma.sync GEMV consumer + timeline A/B against the scalar consumer
OUT=gpurun_out/s6b; mkdir -p $OUT
timeout 600 python -m pytest tests/test_decode_mega_gpu.py -x -q 2>&1 | tail -15 | tee $OUT/pytest_mega.log
for SC in 0 1; do
  OMCHAT_B200_MEGA_SCALAR=$SC OMCHAT_B200_PROF_OUT=$OUT/percta_sc$SC.json timeout 200 python tools/prof_mega.py 28 1 1200 > $OUT/prof_sc$SC.log 2>&1
  echo "== scalar=$SC"; head -12 $OUT/prof_sc$SC.log | cut -c1-200; tail -4 $OUT/prof_sc$SC.log | cut -c1-250
done
OMCHAT_B200_PROF_OUT=$OUT/percta_b4.json timeout 200 python tools/prof_mega.py 28 4 1200 > $OUT/prof_b4.log 2>&1; echo "== batch 4 mma"; head -10 $OUT/prof_b4.log | cut -c1-200
OMCHAT_B200_MEGA_SCALAR=1 timeout 200 python tools/prof_mega.py 28 4 1200 > $OUT/prof_b4_sc.log 2>&1; echo "== batch 4 scalar"; head -10 $OUT/prof_b4_sc.log | cut -c1-200
