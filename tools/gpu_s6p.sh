#!/bin/bash
t() { OMCHAT_B200_MEGA_PROF=0 timeout 200 python tools/prof_mega.py "$@" 2>&1 | grep "step time"; }
for B in 2 3 4; do
  echo "== batch $B sub-ops"; t 28 $B 1200
  echo "== batch $B no sub-ops"; OMCHAT_B200_MEGA_SCALAR=4 t 28 $B 1200
done
echo "== batch 1 ctx 4096: sub / none"; t 28 1 4096; OMCHAT_B200_MEGA_SCALAR=4 t 28 1 4096
