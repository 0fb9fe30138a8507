#!/bin/bash
# compute-sanitizer passes over the kernels added for the InternViT-300M and Qwen2-MoE variants: layernorm, router / plan /
# scatter / combine, the grouped mode of the tcgen05 GEMM, the flag-in-data exchange of gemm_stream (emulated ranks)
OUT=gpurun_out/sanitize_r3; mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
run() { name=$1; tool=$2; shift 2; timeout 700 $CS --tool $tool --error-exitcode 9 python -m pytest "$@" > $OUT/${tool}_$name.log 2>&1; echo "## $tool $name: pytest $* -> rc=$?"; grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Invalid|out of bounds|misaligned" $OUT/${tool}_$name.log | head -8; }
run moe memcheck tests/test_moe_gpu.py -x -q -k "kernels_vs_torch or grouped_gemm or model_vs_reference"
run vit300m memcheck tests/test_vit300m_gpu.py -x -q -k "layernorm or golden or c_entry"
run moe racecheck tests/test_moe_gpu.py -x -q -k "kernels_vs_torch and (37-256 or 700-512)"
run moe initcheck tests/test_moe_gpu.py -x -q -k "kernels_vs_torch and (37-256 or 700-512)"
