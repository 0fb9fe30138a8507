#!/bin/bash
# quick GPU check: usage  gpurun --timeout 900 -- 'bash tools/gpu_quick.sh <tag> [pytest args]'
TAG=${1:-quick}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest "$@" -x -q 2>&1 | tail -40 > $OUT/pytest.log; tail -40 $OUT/pytest.log
