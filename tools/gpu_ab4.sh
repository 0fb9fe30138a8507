#!/bin/bash
# parity then A/B: _base (HEAD of the session start), new default, new with switches. usage: gpu_ab4.sh <tag>
TAG=${1:-ab}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_decode_mega_gpu.py tests/test_model_gpu.py -x -q 2>&1 | tail -15 | tee $OUT/pytest.log
t() { OMCHAT_B200_MEGA_PROF=0 timeout 200 python tools/prof_mega.py "$@" 2>&1 | grep "step time"; }
for i in 1 2; do
  echo "== base"; (cd _base && OMCHAT_B200_MEGA_PROF=0 timeout 200 python tools/prof_mega.py 28 1 1200 2>&1 | grep "step time")
  echo "== new (sub-ops)";  t 28 1 1200
  echo "== new, down split only"; OMCHAT_B200_MEGA_SCALAR=8 t 28 1 1200
  echo "== new, no sub-ops"; OMCHAT_B200_MEGA_SCALAR=4 t 28 1 1200
done
timeout 200 python tools/prof_mega.py 28 1 1200 > $OUT/prof_new.log 2>&1; sed -n 1,12p $OUT/prof_new.log | cut -c1-200
echo "== ctx 8000 / batch 2 / batch 4"
t 28 1 8000; t 28 2 1200; t 28 4 1200
