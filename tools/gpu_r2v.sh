#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:paged_decode_attn -s 8 -c 1 -o gpurun_out/r2v_attn python tools/prof_step.py decode --batch 32 --layers 4 --steps 3 --ctx 1024 > gpurun_out/r2v_ncu.log 2>&1
tail -2 gpurun_out/r2v_ncu.log
ncu -i gpurun_out/r2v_attn.ncu-rep --page raw --csv > gpurun_out/r2v_attn_raw.csv 2>/dev/null
python tools/ncu_summary.py full gpurun_out/r2v_attn.ncu-rep | head -40
