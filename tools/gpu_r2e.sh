#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_decode_mega_gpu.py -m gpu -x -q -k "gemm_stream or batched_decode_stream" 2>&1 | tail -3
: > gpurun_out/r2e_decode_batch.jsonl
run() { env "$@" timeout 300 python tools/bench_decode_batch.py --tag "$*" >> gpurun_out/r2e_decode_batch.jsonl 2>gpurun_out/r2e_err.log || tail -5 gpurun_out/r2e_err.log; }
run OMCHAT_B200_STREAM_CTAS_PER_SM=1 OMCHAT_B200_PDL=1
run OMCHAT_B200_STREAM_CTAS_PER_SM=2 OMCHAT_B200_PDL=1
run OMCHAT_B200_STREAM_CTAS_PER_SM=1 OMCHAT_B200_PDL=0
run OMCHAT_B200_NO_STREAM=1
cat gpurun_out/r2e_decode_batch.jsonl
timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum --kernel-name-base demangled -k regex:omc:: -s 60 -c 30 --csv --log-file gpurun_out/r2e_launches_b32_stream.csv python tools/prof_step.py decode --batch 32 --layers 4 --steps 3 --ctx 1024 > gpurun_out/r2e_prof.log 2>&1
python tools/ncu_summary.py launches gpurun_out/r2e_launches_b32_stream.csv | tail -8
grep gemm_stream gpurun_out/r2e_launches_b32_stream.csv | awk -F'","' '{print $NF}' | tr -d '"' | head -12 | tr '\n' ' '
