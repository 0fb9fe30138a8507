#!/bin/bash
# compute-sanitizer passes over the round-2 kernels: gemm_stream (stream-K, cluster split-K, folded norm), the folded-norm
# epilogue of gemm_sm100, the batched decode step, the model-level C entry points
OUT=gpurun_out/sanitize_r2; mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
run() { name=$1; tool=$2; shift 2; timeout 900 $CS --tool $tool --error-exitcode 9 python -m pytest "$@" > $OUT/${tool}_$name.log 2>&1; echo "## $tool $name: pytest $* -> rc=$?"; grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Invalid|out of bounds|misaligned" $OUT/${tool}_$name.log | head -8; }
run stream memcheck tests/test_kernels_gpu.py -x -q -k "gemm_stream and (M9 or 17 or 64) and (256-64 or 1000-4104 or 4608)"
run folded memcheck tests/test_kernels_gpu.py -x -q -k "folded_rmsnorm and (77-256 or 300-1920)"
run decode memcheck tests/test_decode_mega_gpu.py -x -q -k "batched_decode_stream and (lens0 or lens2)"
run capi memcheck tests/test_model_capi_gpu.py tests/test_serving_gpu.py -x -q
run stream racecheck tests/test_kernels_gpu.py -x -q -k "gemm_stream and M17 and (256-64 or 1000-4104)"
