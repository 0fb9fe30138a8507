#!/bin/bash
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r4e_moe_launches.csv python tools/bench_moe.py --layers 2 --batch 1 --steps 2 --prefill 1 > gpurun_out/r4e.out 2> gpurun_out/r4e.err; echo "rc=$?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r4e_moe_launches.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
seq=[(r[ki], float(r[vi].replace(",",""))/ (1000.0 if r[ui]=="ns" else 1.0)) for r in rows[1:]]
print(len(seq),"launches")
# last decode step = last ~45 launches
for n,t in seq[-34:]: print(f"{t:8.2f} us  {n[:90]}")
PY
