#!/bin/bash
mkdir -p gpurun_out
OMCHAT_B200_STREAM_MIN_B=1 timeout 300 python tools/prof_stream.py --batch 1 --ctx 1100 --layers 4 > gpurun_out/r3c_timeline_b1.txt 2>&1
sed -n 1,12p gpurun_out/r3c_timeline_b1.txt; tail -2 gpurun_out/r3c_timeline_b1.txt
OMCHAT_B200_STREAM_MIN_B=1 timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum --kernel-name-base demangled -k regex:omc:: -s 50 -c 25 --csv --log-file gpurun_out/r3c_launches_b1.csv python tools/prof_step.py decode --batch 1 --layers 4 --steps 3 --ctx 1100 > /dev/null 2>&1
python tools/ncu_summary.py launches gpurun_out/r3c_launches_b1.csv | tail -9
