#!/bin/bash
TAG=${1:-dbg}; O=gpurun_out/$TAG; mkdir -p $O
for cfg in "A=1" "HESIC_TC_EPI8=1" "HESIC_ONE_STREAM=1" "HESIC_ONE_STREAM=1 HESIC_TC_EPI8=1"; do
  echo "== $cfg" >> $O/graph.txt
  env $cfg timeout 300 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -k "cuda_graph" 2>&1 | tail -3 >> $O/graph.txt
done
timeout 900 python -m pytest tests/test_gpu_dsic.py -m gpu -q -x > $O/pytest_dsic.log 2>&1; tail -3 $O/pytest_dsic.log
timeout 300 python tools/dsic_time.py 8 512 512 3 > $O/dsic_time.txt 2>&1; tail -1 $O/dsic_time.txt
cat $O/graph.txt
