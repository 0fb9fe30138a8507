#!/bin/bash
mkdir -p gpurun_out
for tp in 2 4; do timeout 90 python tests/tp_stream_emulation.py $tp > gpurun_out/r2l_tp${tp}.log 2>&1; echo "tp$tp rc=$?"; tail -12 gpurun_out/r2l_tp${tp}.log; done
