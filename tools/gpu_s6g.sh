#!/bin/bash
OUT=gpurun_out/s6g; mkdir -p $OUT
OMCHAT_B200_LIB=$PWD/omchat_b200/_lib/libomchat_b200_detail.so timeout 200 python tools/prof_mega.py 28 1 1200 > $OUT/prof_detail.log 2>&1; sed -n 1,18p $OUT/prof_detail.log | cut -c1-200
OMCHAT_B200_MEGA_PROF=0 timeout 200 python tools/prof_mega.py 28 1 1200 2>&1 | tail -1
