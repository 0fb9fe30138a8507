#!/bin/bash
mkdir -p gpurun_out
# launch list of ONE request of the default bench (after 3 warm-up requests), as in round 1
timeout 900 ncu --clock-control none --metrics gpu__time_duration.sum --kernel-name-base demangled -k regex:omc:: -s 2253 -c 751 --csv --log-file gpurun_out/r3b_launches_bench_c2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/r3b_bench_under_ncu.log 2>&1
python tools/ncu_summary.py launches gpurun_out/r3b_launches_bench_c2.csv > gpurun_out/r3b_launches_bench_c2.txt; cat gpurun_out/r3b_launches_bench_c2.txt
# full capture: the persistent decode kernel (traffic), the stream GEMM on gate|up and down at batch 32
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:decode_mega -s 20 -c 1 -o gpurun_out/r3b_mega python bench.py --steps 1 --warmup 1 --new-tokens 40 --no-cpu-baseline --no-workloads > /dev/null 2>&1
python tools/ncu_summary.py full gpurun_out/r3b_mega.ncu-rep > gpurun_out/r3b_ncu_mega.txt; head -12 gpurun_out/r3b_ncu_mega.txt
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:gemm_stream -s 14 -c 4 -o gpurun_out/r3b_stream python tools/prof_step.py decode --batch 32 --layers 4 --steps 3 --ctx 1024 > /dev/null 2>&1
python tools/ncu_summary.py full gpurun_out/r3b_stream.ncu-rep > gpurun_out/r3b_ncu_stream.txt; grep -A6 "^launch" gpurun_out/r3b_ncu_stream.txt | head -40
