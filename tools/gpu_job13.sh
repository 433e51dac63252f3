#!/bin/bash
mkdir -p gpurun_out/j13
O=gpurun_out/j13
timeout 900 python -m pytest tests/test_gpu_forward.py -q -x > $O/pytest_fwd.log 2>&1; echo "rc=$?" >> $O/pytest_fwd.log
timeout 300 python bench.py --steps 30 --warmup 5 --cpu-iters 2 > $O/bench_two.json 2> $O/bench_two.err
HESIC_ONE_STREAM=1 timeout 300 python bench.py --steps 30 --warmup 5 --cpu-iters 2 > $O/bench_one.json 2> $O/bench_one.err
timeout 300 python bench.py --steps 30 --warmup 5 --cpu-iters 2 --model hesic_plus > $O/bench_plus_two.json 2> $O/bench_plus_two.err
HESIC_ONE_STREAM=1 timeout 300 python bench.py --steps 30 --warmup 5 --cpu-iters 2 --model hesic_plus > $O/bench_plus_one.json 2> $O/bench_plus_one.err
tail -3 $O/pytest_fwd.log
for f in two one plus_two plus_one; do python -c "
import json,sys; d=json.loads(open('$O/bench_$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['clocks'], d['config']['parity_metrics'])"; done
