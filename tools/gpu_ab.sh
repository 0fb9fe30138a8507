#!/bin/bash
# A/B of the persistent decode kernel on ONE box (box-to-box variance is ~10 %, so never compare across gpurun calls):
# parity first, then step times of (a) a baseline checkout in _base/ (git archive <rev> | tar -x -C _base; build it there),
# (b) the working tree, (c) the working tree under each omc_decode_desc.tune value given after the tag.
#   gpurun --timeout 1500 -- 'bash tools/gpu_ab.sh <tag> [tune ...]'      e.g. tune 1 = FFMA dots, 4 = no MLP sub-ops, 176 = 11 ring slots
TAG=${1:-ab}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_decode_mega_gpu.py tests/test_model_gpu.py -x -q 2>&1 | tail -15 | tee $OUT/pytest.log
t() { OMCHAT_B200_MEGA_PROF=0 timeout 200 python tools/prof_mega.py "$@" 2>&1 | grep "step time"; }
for i in 1 2; do
  if [ -d _base ]; then echo "== base"; (cd _base && OMCHAT_B200_MEGA_PROF=0 timeout 200 python tools/prof_mega.py 28 1 1200 2>&1 | grep "step time"); fi
  echo "== new"; t 28 1 1200
  for T in "$@"; do echo "== new, tune $T"; OMCHAT_B200_MEGA_TUNE=$T t 28 1 1200; done
done
timeout 200 python tools/prof_mega.py 28 1 1200 > $OUT/prof_new.log 2>&1; sed -n 1,12p $OUT/prof_new.log | cut -c1-200
echo "== ctx 8000 / batch 2 / batch 4"
t 28 1 8000; t 28 2 1200; t 28 4 1200
