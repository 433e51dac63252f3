#!/bin/bash
TAG=${1:-cols2}; O=gpurun_out/$TAG; mkdir -p $O
LAYERS=("128 128 5 2 0 256 256 16 1 0" "128 128 5 2 1 128 128 16 2 0" "128 960 5 1 0 32 32 16 0 1")
for L in "${LAYERS[@]}"; do
  echo "== $L" >> $O/t.txt
  for cfg in "HESIC_TC_PAIR_NO_COLS=1" "HESIC_TC_NA=2" "HESIC_TC_NA=3" "HESIC_TC_NA=2 HESIC_TC_PF=1" "HESIC_TC_NA=3 HESIC_TC_PF=1"; do
    echo "-- $cfg" >> $O/t.txt
    env $cfg timeout 120 python tools/time_layer.py $L >> $O/t.txt 2>&1
  done
done
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "warp" > $O/pytest_warp.log 2>&1; tail -2 $O/pytest_warp.log
timeout 120 python tools/run_warp.py > $O/warp.txt 2>&1; tail -5 $O/warp.txt
cat $O/t.txt
