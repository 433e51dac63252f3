#!/bin/bash
TAG=${1:-dsicab}; O=gpurun_out/$TAG; mkdir -p $O
for i in 1 2; do
  timeout 300 python tools/dsic_time.py 8 512 512 5 >> $O/dsic_time.txt 2>&1
  HESIC_GN_UNFUSED=1 timeout 300 python tools/dsic_time.py 8 512 512 5 >> $O/dsic_time.txt 2>&1
done
timeout 900 python -m pytest tests/test_gpu_dsic.py tests/test_gpu_fullsize.py -m gpu -q -x > $O/pytest.log 2>&1; tail -3 $O/pytest.log
cat $O/dsic_time.txt
