#!/bin/bash
# A/B on the same box: HEAD (in _base/) against the working tree
OUT=gpurun_out/s6f; mkdir -p $OUT
timeout 600 python -m pytest tests/test_decode_mega_gpu.py -x -q 2>&1 | tail -5 | tee $OUT/pytest_mega.log
for i in 1 2; do
  echo "== base";  (cd _base && OMCHAT_B200_MEGA_PROF=0 timeout 200 python tools/prof_mega.py 28 1 1200 2>&1 | tail -1)
  echo "== new";   OMCHAT_B200_MEGA_PROF=0 timeout 200 python tools/prof_mega.py 28 1 1200 2>&1 | tail -1
  echo "== new scalar";   OMCHAT_B200_MEGA_SCALAR=1 OMCHAT_B200_MEGA_PROF=0 timeout 200 python tools/prof_mega.py 28 1 1200 2>&1 | tail -1
done
timeout 200 python tools/prof_mega.py 28 1 1200 > $OUT/prof_new.log 2>&1; sed -n 1,10p $OUT/prof_new.log | cut -c1-200
(cd _base && timeout 200 python tools/prof_mega.py 28 1 1200 > ../$OUT/prof_base.log 2>&1); sed -n 1,10p $OUT/prof_base.log | cut -c1-200
