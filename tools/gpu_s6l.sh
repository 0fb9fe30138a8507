#!/bin/bash
for i in 1 2; do
for NS in 0 11 10 8; do
  echo "== nslots cap $NS"; OMCHAT_B200_MEGA_SCALAR=$((NS*16)) OMCHAT_B200_MEGA_PROF=0 timeout 200 python tools/prof_mega.py 28 1 1200 2>&1 | grep "step time"
done
done
OMCHAT_B200_MEGA_SCALAR=$((10*16)) timeout 200 python tools/prof_mega.py 28 1 1200 2>&1 | sed -n 1,9p
