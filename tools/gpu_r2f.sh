#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_stream.py --batch 32 --ctx 1024 --layers 4 > gpurun_out/r2f_timeline.txt 2>&1
cat gpurun_out/r2f_timeline.txt
