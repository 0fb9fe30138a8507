#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do timeout 80 python tests/tp_stream_emulation.py 2 > gpurun_out/r3h_tp2.log 2>&1; echo "tp2 rc=$?"; tail -2 gpurun_out/r3h_tp2.log; done
timeout 80 python tests/tp_stream_emulation.py 4 > gpurun_out/r3h_tp4.log 2>&1; echo "tp4 rc=$?"; tail -2 gpurun_out/r3h_tp4.log
