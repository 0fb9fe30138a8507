#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2y_small_batch.jsonl
for b in 1 2 4; do
  OMCHAT_B200_STREAM_MIN_B=1 timeout 300 python tools/bench_decode_batch.py --batch $b --tag stream_b$b >> gpurun_out/r2y_small_batch.jsonl
  timeout 300 python tools/bench_decode_batch.py --batch $b --tag mega_b$b >> gpurun_out/r2y_small_batch.jsonl
done
cat gpurun_out/r2y_small_batch.jsonl
