#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm_stream" > gpurun_out/r2c_stream_kernel.log 2>&1
echo "stream kernel rc=$?"; tail -15 gpurun_out/r2c_stream_kernel.log
timeout 600 python -m pytest tests/test_decode_mega_gpu.py -m gpu -x -q -k "batched_decode_stream" > gpurun_out/r2c_stream_decode.log 2>&1
echo "stream decode rc=$?"; tail -15 gpurun_out/r2c_stream_decode.log
for mode in "stream" "nopdl" "old"; do
  export OMCHAT_B200_NO_STREAM=0 OMCHAT_B200_PDL=1
  [ $mode = nopdl ] && export OMCHAT_B200_PDL=0
  [ $mode = old ] && export OMCHAT_B200_NO_STREAM=1
  timeout 600 python bench.py --workload c4 --steps 2 --warmup 3 > gpurun_out/r2c_c4_$mode.json 2> gpurun_out/r2c_c4_$mode.err
  echo "c4 $mode rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2c_c4_$mode.json").read().strip().splitlines()[-1])
    print("$mode", d["value"], d["phases"], d["roofline"]["frac"])
except Exception as e: print("$mode failed", e)
PY
done
OMCHAT_FULL_PARITY=1 timeout 1500 python -m pytest tests/test_full_width_parity_gpu.py -m gpu -x -q -s -k full_depth > gpurun_out/r2c_full_depth.log 2>&1
echo "full depth rc=$?"; grep -v "hidden state [0-9 ][1-9]\|decode step [0-9 ][1-9]" gpurun_out/r2c_full_depth.log | tail -40
