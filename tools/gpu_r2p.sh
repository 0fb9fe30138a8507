#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_decode_mega_gpu.py -m gpu -x -q -k "gemm_stream or batched_decode_stream or fused_allreduce" 2>&1 | tail -3
timeout 300 python tools/prof_stream.py --batch 32 --ctx 1024 --layers 4 > gpurun_out/r2w_timeline.txt 2>&1
sed -n 1,10p gpurun_out/r2w_timeline.txt
timeout 300 python tools/bench_decode_batch.py --tag dsmem_pipelined > gpurun_out/r2w_decode_batch.jsonl; cat gpurun_out/r2w_decode_batch.jsonl
