#!/bin/bash
mkdir -p gpurun_out
OMCHAT_B200_GEMM_AUTOTUNE=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r4p_c3_launches.csv python bench.py --workload c3 --steps 1 --warmup 3 > gpurun_out/r4p.out 2> gpurun_out/r4p.err; echo "rc=$?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r4p_c3_launches.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
seq=[(r[ki], float(r[vi].replace(",",""))/(1000.0 if r[ui]=="ns" else 1.0)) for r in rows[1:]]
print(len(seq),"launches")
# one pass = from the last im2col launch to the end
idx=[i for i,(n,t) in enumerate(seq) if "im2col" in n]
last=seq[idx[-1]:]
agg=collections.OrderedDict()
for n,t in last:
    k=n.split("(")[0][:60]
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=t
tot=sum(t for _,t in last)
print(f"last pass: {len(last)} launches, {tot/1000:.1f} ms")
for k,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f"{t/1000:9.2f} ms {100*t/tot:5.1f} %  x{c:4d}  {k}")
PY
