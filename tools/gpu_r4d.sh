#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_moe.py --batch 1 --batch 8 --batch 32 > gpurun_out/r4d_moe_bench.jsonl 2> gpurun_out/r4d_moe_bench.err; echo "moe rc=$?"; cat gpurun_out/r4d_moe_bench.jsonl; tail -3 gpurun_out/r4d_moe_bench.err
timeout 200 python tools/bench_vit300m.py > gpurun_out/r4d_vit300m.json 2> gpurun_out/r4d_vit300m.err; echo "vit rc=$?"; cat gpurun_out/r4d_vit300m.json; tail -3 gpurun_out/r4d_vit300m.err
