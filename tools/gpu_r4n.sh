#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_r4i.sh
OUT=gpurun_out/sanitize_r3; mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
run() { name=$1; tool=$2; shift 2; timeout 700 $CS --tool $tool --error-exitcode 9 python -m pytest "$@" > $OUT/${tool}_$name.log 2>&1; echo "## $tool $name: pytest $* -> rc=$?"; grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Invalid|out of bounds|misaligned" $OUT/${tool}_$name.log | head -6; }
run moe2 memcheck tests/test_moe_gpu.py -x -q -k "kernels_vs_torch or teacher_forced or real_width or serving or bookkeeping"
run attn64 memcheck tests/test_kernels_gpu.py tests/test_vit300m_gpu.py -x -q -k "head_dim_64 or zero_padded or c_entry"
