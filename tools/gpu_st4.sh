#!/bin/bash
TAG=${1:-st4}; O=gpurun_out/$TAG; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_forward.py tests/test_gpu_dsic.py -m gpu -q -x > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
LAYERS=("128 128 5 2 0 256 256 16 1 0" "128 128 5 2 1 128 128 16 2 0" "128 960 5 1 0 32 32 16 0 1" "128 128 5 2 0 128 128 16 1 0" "320 128 5 1 0 32 32 16 0 1" "192 128 5 2 1 32 32 16 2 0" "128 192 5 2 0 64 64 16 0 0")
for L in "${LAYERS[@]}"; do
  echo "== $L" >> $O/t.txt
  timeout 120 python tools/time_layer.py $L 2>&1 | tail -1 >> $O/t.txt
  HESIC_TC_PAIR_STAGES=3 timeout 120 python tools/time_layer.py $L 2>&1 | tail -1 >> $O/t.txt
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --cpu-iters 3 > $O/bench.json 2> $O/bench.err
HESIC_TC_PAIR_STAGES=3 timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --cpu-iters 3 > $O/bench_st3.json 2>> $O/bench.err
tail -4 $O/pytest.log; cat $O/t.txt; cut -c1-200 $O/bench.json; cut -c1-200 $O/bench_st3.json
