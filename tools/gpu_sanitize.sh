#!/bin/bash
# compute-sanitizer passes over the persistent decode kernel and the skinny GEMM (tiny + one full-width case)
OUT=gpurun_out/${1:-sanitize}; mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_decode_mega_gpu.py -x -q -k "oracle_tiny or (full_width and lens1)" > $OUT/memcheck_mega.log 2>&1; echo "memcheck mega rc=$?"; tail -4 $OUT/memcheck_mega.log
timeout 600 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -x -q -k "skinny and 3584 and (M32 or 17)" > $OUT/memcheck_skinny.log 2>&1; echo "memcheck skinny rc=$?"; tail -4 $OUT/memcheck_skinny.log
timeout 900 $CS --tool racecheck --racecheck-report analysis --error-exitcode 9 python -m pytest tests/test_decode_mega_gpu.py -x -q -k "oracle_tiny" > $OUT/racecheck_mega.log 2>&1; echo "racecheck mega rc=$?"; grep -c "RACECHECK SUMMARY\|Race reported\|hazard" $OUT/racecheck_mega.log; tail -6 $OUT/racecheck_mega.log
