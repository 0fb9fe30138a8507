#!/bin/bash
mkdir -p gpurun_out
for f in 1 0 1 0; do OMCHAT_B200_FOLD_NORMS=$f timeout 600 python bench.py --workload c4 --steps 2 --warmup 3 > gpurun_out/r3e_c4_fold$f.json 2>gpurun_out/r3e_err.log; python - <<PY
import json
d=json.loads(open("gpurun_out/r3e_c4_fold$f.json").read().strip().splitlines()[-1])
print("c4 fold=$f", round(d["value"]), {k: round(v,2) for k,v in d["phases"].items()}, d["clocks"])
PY
done
