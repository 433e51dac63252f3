#!/bin/bash
TAG=${1:-r3g}; O=gpurun_out/$TAG; mkdir -p $O
for L in "128 128 5 2 0 32 32 16 0 0" "128 128 5 2 0 16 16 16 0 0" "128 128 5 2 1 8 8 16 0 0" "128 128 5 2 1 16 16 16 0 0" "128 128 5 1 0 32 32 16 0 1" "320 128 5 1 0 32 32 16 0 1" "192 128 5 1 0 32 32 16 0 0" "128 960 5 2 1 16 16 16 0 0" "192 128 5 2 1 32 32 16 2 0" "128 128 5 2 1 64 64 16 2 0"; do
  timeout 120 python tools/time_layer.py $L 10 >> $O/t.txt 2>&1
done
timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_fullsize.py -q -x -k "perspective or max_pool or dsic_at_512" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
cat $O/t.txt
