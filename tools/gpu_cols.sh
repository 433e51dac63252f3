#!/bin/bash
# A/B of the pair kernel's activation column reuse:  tools/gpu_cols.sh <tag>
TAG=${1:-cols}; O=gpurun_out/$TAG; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_forward.py -m gpu -q -x > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
LAYERS=("128 128 5 2 0 256 256 16 1 0" "128 128 5 2 1 128 128 16 2 0" "128 960 5 1 0 32 32 16 0 1" "128 128 5 2 0 128 128 16 1 0" "320 128 5 1 0 32 32 16 0 1" "192 128 5 2 1 32 32 16 2 0" "128 192 5 2 0 64 64 16 0 0")
for L in "${LAYERS[@]}"; do
  echo "== $L" >> $O/time_layer.txt
  timeout 120 python tools/time_layer.py $L >> $O/time_layer.txt 2>&1
  HESIC_TC_PAIR_NO_COLS=1 timeout 120 python tools/time_layer.py $L >> $O/time_layer.txt 2>&1
done
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:"conv_tc_pair_kernel" -s 2 -c 1 -o $O/conv2_cols python tools/run_layer.py 128 128 5 2 0 256 256 16 1 0 3 > $O/ncu1.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --cpu-iters 3 > $O/bench.json 2> $O/bench.err
tail -4 $O/pytest.log; cat $O/time_layer.txt; cut -c1-300 $O/bench.json
