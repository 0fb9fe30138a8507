#!/bin/bash
TAG=${1:-ab}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_decode_mega_gpu.py -x -q 2>&1 | tail -15 | tee $OUT/pytest.log
for i in 1 2; do
  echo "== base"; (cd _base && OMCHAT_B200_MEGA_PROF=0 timeout 200 python tools/prof_mega.py 28 1 1200 2>&1 | grep "step time")
  echo "== new";  OMCHAT_B200_MEGA_PROF=0 timeout 200 python tools/prof_mega.py 28 1 1200 2>&1 | grep "step time"
  echo "== new, one-row split-K stages";  OMCHAT_B200_MEGA_SCALAR=2 OMCHAT_B200_MEGA_PROF=0 timeout 200 python tools/prof_mega.py 28 1 1200 2>&1 | grep "step time"
done
timeout 200 python tools/prof_mega.py 28 1 1200 > $OUT/prof_new.log 2>&1; sed -n 1,9p $OUT/prof_new.log | cut -c1-200
