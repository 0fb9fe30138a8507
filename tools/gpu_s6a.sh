#!/bin/bash
# session-6 call A: parity of the new ring geometry + timeline A/B of ring slot sizes
OUT=gpurun_out/s6a; mkdir -p $OUT
timeout 600 python -m pytest tests/test_decode_mega_gpu.py -x -q 2>&1 | tail -5 | tee $OUT/pytest_mega.log
for SLOT in 14336 19456 28672; do
  OMCHAT_B200_MEGA_SLOT=$SLOT OMCHAT_B200_PROF_OUT=$OUT/percta_$SLOT.json timeout 200 python tools/prof_mega.py 28 1 1200 > $OUT/prof_$SLOT.log 2>&1
  echo "== slot $SLOT"; head -12 $OUT/prof_$SLOT.log | cut -c1-200; tail -4 $OUT/prof_$SLOT.log | cut -c1-250
done
