#!/bin/bash
# per-kernel timings of the HBM-side kernels through bench.py's hbm_kernels extra (no headline loop changes)
O=gpurun_out/$1; mkdir -p $O
timeout 600 python bench.py --steps 5 --warmup 3 --cpu-iters 1 --sustain-s 0.01 > $O/b.json 2> $O/b.err
python - <<PY
import json
d=json.loads(open("$O/b.json").read())
for k,v in d["extras"]["hbm_kernels"]["kernels"].items(): print(f"{v['us']:7.1f} us  {v['frac_of_hbm_peak']:.2f}  {k}")
PY
