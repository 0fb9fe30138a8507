#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_decode_mega_gpu.py tests/test_serving_gpu.py tests/test_model_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/prof_stream.py --batch 32 --ctx 1024 --layers 4 > gpurun_out/r2x_timeline.txt 2>&1
sed -n 5,10p gpurun_out/r2x_timeline.txt
timeout 300 python tools/bench_decode_batch.py --tag attn_prefetch > gpurun_out/r2x_decode_batch.jsonl; cat gpurun_out/r2x_decode_batch.jsonl
timeout 300 python tools/bench_decode_batch.py --batch 8 --tag b8 >> gpurun_out/r2x_decode_batch.jsonl; timeout 300 python tools/bench_decode_batch.py --batch 64 --tag b64 >> gpurun_out/r2x_decode_batch.jsonl; tail -2 gpurun_out/r2x_decode_batch.jsonl
