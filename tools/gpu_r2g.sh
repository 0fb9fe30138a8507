#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_decode_mega_gpu.py -m gpu -x -q -k "gemm_stream or batched_decode_stream" 2>&1 | tail -8
timeout 300 python tools/prof_stream.py --batch 32 --ctx 1024 --layers 4 > gpurun_out/r2h_timeline.txt 2>&1
sed -n 1,10p gpurun_out/r2h_timeline.txt; tail -2 gpurun_out/r2h_timeline.txt
: > gpurun_out/r2h_decode_batch.jsonl
run() { env "$@" timeout 300 python tools/bench_decode_batch.py --tag "$*" >> gpurun_out/r2h_decode_batch.jsonl 2>gpurun_out/r2h_err.log || tail -5 gpurun_out/r2h_err.log; }
run OMCHAT_B200_PDL=1
run OMCHAT_B200_PDL=0
cat gpurun_out/r2h_decode_batch.jsonl
